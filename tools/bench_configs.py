#!/usr/bin/env python
"""Kernel-only throughput of every BASELINE.json config on one GPU (not the driver's bench line: that is bench.py / C2).
Writes a JSON report (default profiles/rNN_configs.json).  Usage: python tools/bench_configs.py [--out FILE] [--only C1,C3]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ray_tracing_octrees_b200 as rto

DT_GRID = os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz")


def timed(scene, cams, mode, flags, bias, H, W, reps=5):
    F = len(cams)
    rgba = torch.empty((F, H, W, 4), dtype=torch.float32, device="cuda")
    hid = torch.empty((F, H, W), dtype=torch.int32, device="cuda")
    tt = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
    ms = []
    for r in range(reps + 2):
        scene.render_device(cams, mode, flags, bias, 0, H, rgba.data_ptr(), hid.data_ptr(), tt.data_ptr())
        ms.append(scene.last_kernel_ms())
    hits = int((hid >= 0).sum().item())
    rays = F * W * H + (hits if (flags & rto.FLAG_SHADOWS) else 0)
    best = float(np.median(ms[2:]))
    return dict(ms=best, rays=rays, hit_fraction=hits / (F * W * H), Mrays_s=rays / best / 1e3)


def cams_orbit(theta, radius, W, H, n, phi0=40.0, target=(0, 0, 0)):
    return [rto.Camera.from_degrees(theta, phi0 + 360.0 * k / max(n, 1), radius, target).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0] for k in range(n)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_configs.json"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    assert rto.lib().rto_init(0) == 0
    rep = {}

    def want(c):
        return not only or c in only

    if want("C1"):
        g = rto.generate_test_volume(128)
        nodes = rto.create_octree_from_voxel_grid(g)
        tris = rto.marching_cubes_mesh(g, nodes)
        sc = rto.Scene.bvh(tris)
        cams = cams_orbit(30, 1.2, 1024, 768, 8)
        rep["C1 sphere128 MC mesh BVH primary 1024x768 (8 frames)"] = dict(tris=len(tris), **timed(sc, cams, rto.MODE_BVH, 0, 0.0, 768, 1024))
        oc = rto.Scene.octree(nodes, g.min, g.voxel_size)
        for mode, name in ((rto.MODE_OCTREE_SKIP, "A octreeRaySkip"), (rto.MODE_OCTREE_GLSL, "B GLSL")):
            rep["C1 sphere128 octree mode %s 1024x768 (8 frames)" % name] = dict(nodes=len(nodes), **timed(oc, cams, mode, 0, 0.0, 768, 1024))
    if want("C2") or want("C5") or want("C2DC"):
        g = rto.VoxelGrid.load(DT_GRID)
        nodes = rto.create_octree_from_voxel_grid(g)
        tris = rto.marching_cubes_mesh(g, nodes)
        sc = rto.Scene.bvh(tris)
        bias = 1e-3 * g.voxel_size
        if want("C2"):
            for name, (th, ph, r) in {"far": (35, 40, 0.6 * 4250), "near": (60, 10, 0.35 * 4250)}.items():
                cams = [rto.Camera.from_degrees(th, ph, r).consts(45.0, float(np.float32(1920) / np.float32(1080)), 1920, 1080)[0]] * 8
                rep["C2 DT mesh BVH primary+shadow 1080p camera %s (8 frames)" % name] = dict(tris=len(tris), **timed(sc, cams, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 1080, 1920))
                rep["C2 DT mesh BVH primary only 1080p camera %s (8 frames)" % name] = timed(sc, cams, rto.MODE_BVH, 0, bias, 1080, 1920)
            oc = rto.Scene.octree(nodes, g.min, g.voxel_size)
            cams = cams_orbit(35, 0.6 * 4250, 1920, 1080, 8)
            for mode, name in ((rto.MODE_OCTREE_SKIP, "A octreeRaySkip"), (rto.MODE_OCTREE_GLSL, "B GLSL")):
                rep["DT octree mode %s 1080p (8 frames)" % name] = dict(nodes=len(nodes), **timed(oc, cams, mode, 0, 0.0, 1080, 1920))
        if want("C2DC"):
            dct = rto.dual_contouring_mesh(g, nodes)
            sdc = rto.Scene.bvh(dct)
            for name, (th, ph, r) in {"far": (35, 40, 0.6 * 4250), "near": (60, 10, 0.35 * 4250)}.items():
                cams = [rto.Camera.from_degrees(th, ph, r).consts(45.0, float(np.float32(1920) / np.float32(1080)), 1920, 1080)[0]] * 8
                rep["C2 DT Dual-Contouring mesh BVH primary+shadow 1080p camera %s (8 frames)" % name] = dict(tris=len(dct), **timed(sdc, cams, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 1080, 1920))
            del sdc
        if want("C5"):
            cams = cams_orbit(35, 0.6 * 4250, 3840, 2160, 16, phi0=0.0)
            rep["C5 DT mesh 4K orbit primary+shadow (16 of 64 frames per launch)"] = timed(sc, cams, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 2160, 3840, reps=3)
    if want("C3"):
        t0 = time.time()
        g = rto.city_block_grid(512, 1234, 32)
        nodes = rto.create_octree_from_voxel_grid(g)
        build = time.time() - t0
        oc = rto.Scene.octree(nodes, g.min, g.voxel_size)
        cams = cams_orbit(35, 0.9 * 512, 1920, 1080, 8)
        for mode, name in ((rto.MODE_OCTREE_SKIP, "A octreeRaySkip"), (rto.MODE_OCTREE_GLSL, "B GLSL")):
            r = timed(oc, cams, mode, 0, 0.0, 1080, 1920)
            r["nodes_visited_per_ray"] = float(oc.stats(cams[0], mode)[0]) / (1920 * 1080)
            rep["C3 city 512^3 octree mode %s 1080p (8 frames)" % name] = dict(nodes=len(nodes), grid_and_octree_build_s=build, **r)
    if want("C4"):
        t0 = time.time()
        g = rto.city_block_grid(1024, 4321, 64)
        nodes = rto.create_octree_from_voxel_grid(g)
        tris = rto.marching_cubes_mesh(g, nodes)
        t1 = time.time()
        sc = rto.Scene.bvh(tris)
        t2 = time.time()
        cams = cams_orbit(35, 0.9 * 1024, 3840, 2160, 4)
        rep["C4 city 1024^3 MC mesh (DC mesher out of scope) BVH 4K primary+shadow (4 frames)"] = dict(
            tris=len(tris), octree_nodes=len(nodes), grid_octree_mc_s=t1 - t0, bvh_build_upload_s=t2 - t1, device_MB=sc.info()["device_bytes"] / 1e6,
            **timed(sc, cams, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3, 2160, 3840, reps=3))
    if want("C4DC"):
        # C4 as BASELINE.json names it: the Adaptive Dual Contouring mesh (rto_host_dc_mesh == the reference's createTriangles per leaf)
        t0 = time.time()
        g = rto.city_block_grid(1024, 4321, 64)
        nodes = rto.create_octree_on_device(g)
        t1 = time.time()
        tris = rto.dual_contouring_mesh(g, nodes)
        t2 = time.time()
        tdev = rto.dual_contouring_mesh(g, nodes, algo="device")
        t2d = time.time()
        rep["C4 Dual-Contouring mesh extraction, host (all threads) vs device (incl. grid + node upload and triangle read-back)"] = dict(
            ms=0.0, Mrays_s=0.0, hit_fraction=0.0, host_s=t2 - t1, device_s=t2d - t2, identical=bool(tris.shape == tdev.shape and np.array_equal(tris.view(np.uint32), tdev.view(np.uint32))))
        del tdev
        cams = cams_orbit(35, 0.9 * 1024, 3840, 2160, 4)
        t5 = time.time()
        sg = rto.Scene.bvh_from_grid_dc(g)
        t6 = time.time()
        rep["C4 city 1024^3: grid -> octree -> Dual-Contouring mesh -> linear BVH entirely on the device, 4K primary+shadow (4 frames)"] = dict(
            tris=sg.info()["prims"], scene_from_grid_s=t6 - t5, device_MB=sg.info()["device_bytes"] / 1e6,
            **timed(sg, cams, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3, 2160, 3840, reps=3))
        del sg
        if os.environ.get("RTO_C4DC_HOST_TREE", "1") == "1":
            sc = rto.Scene.bvh(tris)
            t3 = time.time()
            rep["C4 city 1024^3 Dual-Contouring mesh BVH 4K primary+shadow (4 frames), reference-shaped tree + SAH"] = dict(
                tris=len(tris), octree_nodes=len(nodes), grid_octree_s=t1 - t0, dc_mesh_s=t2 - t1, bvh_build_upload_s=t3 - t2, device_MB=sc.info()["device_bytes"] / 1e6,
                **timed(sc, cams, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3, 2160, 3840, reps=3))
            del sc
        t3 = time.time()
        sd = rto.Scene.bvh_device(tris)
        t4 = time.time()
        rep["C4 city 1024^3 Dual-Contouring mesh BVH 4K primary+shadow (4 frames), device-built LBVH"] = dict(
            tris=len(tris), dc_mesh_s=t2 - t1, bvh_build_upload_s=t4 - t3, device_MB=sd.info()["device_bytes"] / 1e6,
            **timed(sd, cams, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3, 2160, 3840, reps=3))
    for k, v in rep.items():
        print("%-90s %8.0f Mrays/s  %7.3f ms  hit %.2f" % (k, v["Mrays_s"], v["ms"], v["hit_fraction"]))
    json.dump(rep, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
