#!/usr/bin/env python
"""A/B timing of the BVH kernel on the DT mesh, the 512^3 city mesh and the C1 sphere mesh (primary + shadow rays) in ONE process per
library variant:
    RTO_LIB_VARIANT=<name> python tools/experiments/ab_bvh.py [--frames 16] [--reps 7] [--scenes dt,c3,c1]
prints the median kernel time of a launch of `frames` 1080p orbit frames and checksums of ids, t and colours (variants must reproduce
the baseline's checksums digit for digit)."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import ray_tracing_octrees_b200 as rto

DT_GRID = os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--scenes", default="dt,c3,c1")
    a = ap.parse_args()
    assert rto.lib().rto_init(0) == 0
    tag = os.environ.get("RTO_LIB_VARIANT", "baseline")
    F = a.frames
    for name in a.scenes.split(","):
        W, H = 1920, 1080
        if name == "dt":
            g = rto.VoxelGrid.load(DT_GRID); radius, theta = 0.6 * 4250, 35
        elif name == "c3":
            g = rto.city_block_grid(512, 1234, 32); radius, theta = 0.9 * 512, 35
        else:
            g = rto.generate_test_volume(128); radius, theta = 1.2, 30
        nodes = rto.create_octree_from_voxel_grid(g)
        sc = rto.Scene.bvh(rto.marching_cubes_mesh(g, nodes))
        bias = 1e-3 * g.voxel_size
        cams = [rto.Camera.from_degrees(theta, 40.0 + 360.0 / 64 * k, radius).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0] for k in range(F)]
        hid = torch.empty((F, H, W), dtype=torch.int32, device="cuda")
        tt = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
        rgba = torch.empty((F, H, W, 4), dtype=torch.float32, device="cuda")
        for flags, label in ((rto.FLAG_SHADOWS, "primary+shadow"), (0, "primary")):
            ms = []
            for r in range(a.reps):
                sc.render_device(cams, rto.MODE_BVH, flags, bias, 0, H, rgba.data_ptr(), hid.data_ptr(), tt.data_ptr())
                ms.append(sc.last_kernel_ms())
            best = float(np.median(ms[2:]))
            hits = int((hid >= 0).sum().item())
            rays = F * W * H + (hits if flags else 0)
            print("[%s] %s-bvh %s: %.3f ms per %d frames  %.0f Mrays/s  checksum id %d t %.6e rgba %.6e" % (tag, name, label, best, F, rays / best / 1e3,
                  int(hid.to(torch.int64).sum().item()), float(tt[hid >= 0].double().sum().item()), float(rgba.double().sum().item())), flush=True)
        del sc


if __name__ == "__main__":
    main()
