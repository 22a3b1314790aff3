#!/usr/bin/env python
"""One frame per call into pinned host memory -- the call RayTracerBVH::renderSceneCompute maps to -- on the DT mesh at 1080p and 4K:
    RTO_HOST_BANDS=<n> python tools/experiments/e2e_single_frame.py
wall-clock ms per rto_render(RTO_MEM_HOST) call (all three planes, rgba only), median over the orbit; RTO_HOST_BANDS=1 is one launch and
one set of copies per frame, the default cuts the frame into 4 row bands whose copies overlap the next band's trace."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import ray_tracing_octrees_b200 as rto

assert rto.lib().rto_init(0) == 0
g = rto.VoxelGrid.load(os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz"))
sc = rto.Scene.bvh(rto.marching_cubes_mesh(g, rto.create_octree_from_voxel_grid(g)))
bias = 1e-3 * g.voxel_size
for W, H in ((1920, 1080), (3840, 2160)):
    pageable = os.environ.get("E2E_PAGEABLE") == "1"       # ordinary memory, as a std::vector would be: the driver stages the copy
    h_rgba = torch.zeros((H, W, 4), dtype=torch.float32); h_id = torch.zeros((H, W), dtype=torch.int32); h_t = torch.zeros((H, W), dtype=torch.float32)
    if not pageable:
        h_rgba, h_id, h_t = h_rgba.pin_memory(), h_id.pin_memory(), h_t.pin_memory()
    cams = [(rto.RtoCamera * 1)(rto.Camera.from_degrees(35, 40.0 + 360.0 / 64 * k, 0.6 * 4250).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0]) for k in range(64)]
    for label, ptrs in (("rgba + id + t", (h_rgba.data_ptr(), h_id.data_ptr(), h_t.data_ptr())), ("rgba only", (h_rgba.data_ptr(), None, None))):
        ms = []
        for k in range(64 + 8):
            t0 = time.perf_counter()
            sc.render_host_ptrs(cams[k % 64], rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 0, H, *ptrs)
            ms.append((time.perf_counter() - t0) * 1e3)
        ms = ms[8:]
        print("[%s, bands %s] %dx%d %s: median %.3f ms per call, min %.3f; checksum id %d" % ("pageable" if pageable else "pinned", os.environ.get("RTO_HOST_BANDS", "default"), W, H, label,
              float(np.median(ms)), min(ms), int(h_id.to(torch.int64).sum().item())), flush=True)
