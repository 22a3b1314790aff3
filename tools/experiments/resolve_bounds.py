#!/usr/bin/env python
"""What bounds k_resolve_bvh: the expansion of an all-miss code buffer (pure streaming: 4 bytes in, 24 bytes out per pixel) against the
DT orbit frames (hit rate 0.32) and against a close-up (hit rate ~1), us per 1080p frame."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import ray_tracing_octrees_b200 as rto

assert rto.lib().rto_init(0) == 0
g = rto.VoxelGrid.load(os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz"))
nodes = rto.create_octree_from_voxel_grid(g)
sc = rto.Scene.bvh(rto.marching_cubes_mesh(g, nodes))
W, H, F = 1920, 1080, 16
bias = 1e-3 * g.voxel_size
rgba = torch.empty((F, H, W, 4), dtype=torch.float32, device="cuda")
hid = torch.empty((F, H, W), dtype=torch.int32, device="cuda")
tt = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
words = rto.codes_frame_words(W, H)
buf = rto.ExchangeBuffer(words * 4 * F)
for name, theta, radius in (("orbit", 35.0, 0.6 * 4250), ("close-up", 80.0, 0.12 * 4250), ("all-miss", None, None)):
    cams = [rto.Camera.from_degrees(theta or 35.0, 40.0 + 360.0 / 64 * k, radius or 0.6 * 4250).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0] for k in range(F)]
    arr = (rto.RtoCamera * F)(*cams)
    if theta is None:
        torch.cuda.synchronize()
        import ctypes
        from cuda import cudart
        cudart.cudaMemset(buf.ptr, 0, words * 4 * F)
        cudart.cudaDeviceSynchronize()
    else:
        sc.render_codes(arr, rto.FLAG_SHADOWS, bias, buf.ptr)
        sc.sync()
    ms = []
    for r in range(8):
        sc.resolve_codes(arr, buf.ptr, rgba_ptr=rgba.data_ptr(), id_ptr=hid.data_ptr(), t_ptr=tt.data_ptr())
        ms.append(sc.last_kernel_ms())
    hit = float((hid >= 0).float().mean().item())
    print("%-9s hit rate %.3f  expansion %.1f us per 1080p frame (%.2f TB/s of codes + planes)" % (name, hit, 1e3 * float(np.median(ms[2:])) / F, 28.0 * W * H / (1e-3 * float(np.median(ms[2:])) / F) / 1e12))
