#!/usr/bin/env python
"""Explicit ray lists through rto_trace_rays, in the caller's order and with RTO_FLAG_SORT_RAYS (ray coherence sorting on the device):
the camera rays of a 1080p frame in pixel order, the same rays shuffled, and incoherent rays (random origins on a sphere around the
scene, aimed at random points inside it).  Times include key generation and the radix sort.  Writes a JSON report."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ray_tracing_octrees_b200 as rto

DT_GRID = os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz")


def camera_rays(cam, view_inv, W, H):
    """generateRay (RayTracerBVH.cpp:338-355) in numpy float32; only used to make a ray list, not a parity path."""
    px, py = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    nx = ((px + 0.5) / W * 2 - 1) * cam.aspect * cam.tanHalfFov
    ny = (1 - (py + 0.5) / H * 2) * cam.tanHalfFov
    d = np.stack([nx, ny, -np.ones_like(nx)], -1).reshape(-1, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    m = np.array(list(cam.invView), np.float32).reshape(4, 4).T[:3, :3]
    d = (d @ m.T).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.tile(np.array(list(cam.camPos), np.float32), (len(d), 1))
    return o, d.astype(np.float32)


def timed(scene, o, d, mode, flags, reps=6):
    n = len(o)
    do, dd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    t = torch.empty(n, dtype=torch.float32, device="cuda"); ids = torch.empty(n, dtype=torch.int32, device="cuda")
    ms = []
    for _ in range(reps):
        scene.trace_rays_device(do.data_ptr(), dd.data_ptr(), n, mode, flags, t.data_ptr(), ids.data_ptr())
        ms.append(scene.last_kernel_ms())
    best = float(np.median(ms[2:]))
    return dict(ms=best, Mrays_s=n / best / 1e3, hits=int((ids >= 0).sum().item())), t.cpu().numpy(), ids.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_ray_lists.json"))
    a = ap.parse_args()
    assert rto.lib().rto_init(0) == 0
    g = rto.VoxelGrid.load(DT_GRID)
    nodes = rto.create_octree_from_voxel_grid(g)
    mesh = rto.Scene.bvh(rto.marching_cubes_mesh(g, nodes))
    octree = rto.Scene.octree(nodes, g.min, g.voxel_size)
    W, H = 1920, 1080
    cam, _ = rto.Camera.from_degrees(35, 40, 0.6 * 4250).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)
    o, d = camera_rays(cam, None, W, H)
    rng = np.random.default_rng(1)
    perm = rng.permutation(len(o))
    n = 4 * 1024 * 1024
    ext = np.array(g.dims, np.float32) * g.voxel_size
    centre = np.array(g.min, np.float32) + 0.5 * ext
    v = rng.normal(size=(n, 3)).astype(np.float32); v /= np.linalg.norm(v, axis=1, keepdims=True)
    ro = (centre + v * np.float32(0.8 * np.linalg.norm(ext))).astype(np.float32)
    aim = (np.array(g.min, np.float32) + rng.random((n, 3)).astype(np.float32) * ext).astype(np.float32)
    rd = aim - ro; rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    rep = {}
    for sname, scene, mode in (("DT mesh BVH closest hit", mesh, rto.MODE_BVH), ("DT octree mode A", octree, rto.MODE_OCTREE_SKIP), ("DT octree mode B", octree, rto.MODE_OCTREE_GLSL)):
        for lname, (lo, ld) in {"camera rays in pixel order (row-major)": (o, d), "camera rays shuffled": (o[perm], d[perm]), "incoherent rays (random origins around the scene)": (ro, rd.astype(np.float32))}.items():
            plain, t0, i0 = timed(scene, lo, ld, mode, 0)
            srt, t1, i1 = timed(scene, lo, ld, mode, rto.FLAG_SORT_RAYS)
            same = bool(np.array_equal(i0, i1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32)))
            rep["%s | %s" % (sname, lname)] = dict(rays=len(lo), as_given=plain, sorted=srt, identical_results=same)
            print("%-28s %-52s %9.0f -> %9.0f Mrays/s  (%.3f -> %.3f ms)  identical %s" % (sname, lname, plain["Mrays_s"], srt["Mrays_s"], plain["ms"], srt["ms"], same))
    json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
