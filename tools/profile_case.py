#!/usr/bin/env python
"""One named case rendered a few times -- the target of `ncu -k regex:k_render -s 2 -c 1` captures and of quick A/B timing.
    python tools/profile_case.py CASE [--reps N]      CASE in dt-bvh, dt-bvh-primary, dt-dcbvh, dt-octA, dt-octB, c3-octA, c3-octB, c3-dcbvh, c4-dcbvh, c1-bvh, c3-dcbuild, c4-dcbuild, dt-bvh-resolve"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ray_tracing_octrees_b200 as rto

DT_GRID = os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--phi-step", type=float, default=360.0 / 64, help="orbit step between frames in degrees (tools/bench_configs.py uses 45)")
    a = ap.parse_args()
    assert rto.lib().rto_init(0) == 0
    W, H, F = 1920, 1080, a.frames
    flags, bias = 0, 0.0
    if a.case.startswith("dt") or a.case.startswith("c5"):
        g = rto.VoxelGrid.load(DT_GRID)
        radius, theta = 0.6 * 4250, 35
    elif a.case.startswith("c3"):
        g = rto.city_block_grid(512, 1234, 32)
        radius, theta = 0.9 * 512, 35
    elif a.case.startswith("c4"):
        g = rto.city_block_grid(1024, 4321, 64)
        radius, theta, W, H = 0.9 * 1024, 35, 3840, 2160
    else:
        g = rto.generate_test_volume(128)
        radius, theta, W, H = 1.2, 30, 1024, 768
    if a.case.endswith("-dcbuild"):
        # scene construction only: grid -> octree -> Dual-Contouring mesh -> linear BVH on the device (launch-list target)
        if a.case.startswith("c4"):
            g = rto.city_block_grid(1024, 4321, 64)
        import time
        for r in range(2):
            t0 = time.time(); sc = rto.Scene.bvh_from_grid_dc(g); dt = time.time() - t0
            print("%s: scene from grid in %.3f s, %d triangles" % (a.case, dt, sc.info()["prims"]))
            del sc
        return
    nodes = rto.create_octree_on_device(g) if a.case.startswith("c4") else rto.create_octree_from_voxel_grid(g)
    if "devbvh" in a.case:          # the device-built tree over the Marching-Cubes soup 
        import time
        tris = rto.marching_cubes_mesh_on_device(g) if a.case.startswith("c4") else rto.marching_cubes_mesh(g, nodes)
        t0 = time.time(); sc = rto.Scene.bvh_device(tris); print("device BVH over %d triangles built in %.3f s" % (len(tris), time.time() - t0))
        mode, flags, bias = rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3 * g.voxel_size
    elif "dcbvh" in a.case:
        sc = rto.Scene.bvh(rto.dual_contouring_mesh(g, nodes, algo="device" if a.case.startswith("c4") else "default"))
        mode, flags, bias = rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3 * g.voxel_size
    elif "bvh" in a.case:
        sc = rto.Scene.bvh(rto.marching_cubes_mesh(g, nodes))
        mode = rto.MODE_BVH
        if "primary" not in a.case and not a.case.startswith("c1"):
            flags, bias = rto.FLAG_SHADOWS, 1e-3 * g.voxel_size
    else:
        sc = rto.Scene.octree(nodes, g.min, g.voxel_size)
        mode = rto.MODE_OCTREE_SKIP if a.case.endswith("A") else rto.MODE_OCTREE_GLSL
    cams = [rto.Camera.from_degrees(theta, 40.0 + a.phi_step * k, radius).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0] for k in range(F)]
    rgba = torch.empty((F, H, W, 4), dtype=torch.float32, device="cuda")
    hid = torch.empty((F, H, W), dtype=torch.int32, device="cuda")
    tt = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
    ms = []
    if a.case.endswith("-resolve"):
        # the expansion of hit codes into planes (k_resolve_bvh): what the gathering GPU pays per frame it did not trace itself
        buf = rto.ExchangeBuffer(rto.codes_frame_words(W, H) * 4 * F)
        arr = (rto.RtoCamera * F)(*cams)
        sc.render_codes(arr, flags, bias, buf.ptr)
        cms = sc.last_kernel_ms()
        for r in range(a.reps):
            sc.resolve_codes(arr, buf.ptr, rgba_ptr=rgba.data_ptr(), id_ptr=hid.data_ptr(), t_ptr=tt.data_ptr())
            ms.append(sc.last_kernel_ms())
        print("%s: trace into codes %.3f ms for %d frames; expansion %.3f ms = %.1f us per frame" % (a.case, cms, F, float(np.median(ms[2:])), 1e3 * float(np.median(ms[2:])) / F))
    for r in range(0 if a.case.endswith("-resolve") else a.reps):
        sc.render_device(cams, mode, flags, bias, 0, H, rgba.data_ptr(), hid.data_ptr(), tt.data_ptr())
        ms.append(sc.last_kernel_ms())
    hits = int((hid >= 0).sum().item())
    rays = F * W * H + (hits if flags else 0)
    best = float(np.median(ms[2:])) if len(ms) > 2 else ms[-1]
    print("%s: %.3f ms  %.0f Mrays/s  hit %.3f  checksum id %d t %.6e" % (a.case, best, rays / best / 1e3, hits / (F * W * H),
          int(hid.to(torch.int64).sum().item()), float(tt[hid >= 0].double().sum().item())))


if __name__ == "__main__":
    main()
