"""One process per GPU (torch.distributed over NCCL, as bench.py runs under torchrun): frames traced by all ranks and delivered on
rank 0 (sharding.GatheredRenderer) equal a local render bit for bit, with hit codes written through CUDA-IPC peer mappings and with
the NCCL send/receive fallback.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box, where tests/test_gpu_codes.py covers the same
kernels, the IPC mapping between two processes and the device group."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, transport, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=dev)
        import ray_tracing_octrees_b200 as rto
        from ray_tracing_octrees_b200 import sharding
        assert rto.lib().rto_init(rank) == 0
        grid = rto.VoxelGrid.load(os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz"))
        tris = rto.marching_cubes_mesh(grid, rto.create_octree_from_voxel_grid(grid))
        scene = rto.Scene.bvh(tris)
        bias = float(np.float32(1e-3) * np.float32(grid.voxel_size))
        W, H, n = 640, 360, 6
        cams = (rto.RtoCamera * n)(*[rto.Camera.from_degrees(35.0, 40.0 + 60.0 * k, 0.6 * 4250).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0] for k in range(n)])
        gr = sharding.GatheredRenderer(rto, scene, W, H, n, rto.FLAG_SHADOWS, bias, dev, transport=transport)
        planes = [dict(rgba=torch.zeros((n, H, W, 4), dtype=torch.float32, device=dev), id=torch.zeros((n, H, W), dtype=torch.int32, device=dev),
                       t=torch.zeros((n, H, W), dtype=torch.float32, device=dev)) for _ in range(2)]
        ok = True
        msg = ""
        rounds = [None, [0.2] + [1.0] * (world - 1), [3.0] + [0.4] * (world - 1)]
        for rnd, wts in enumerate(rounds * 2):         # several batches back to back: both code buffers, re-use while the other one is expanded
            if wts is not None:
                gr.weights = list(wts)
            p = planes[rnd & 1]
            gr.render(cams, p["rgba"].data_ptr(), p["id"].data_ptr(), p["t"].data_ptr())
        hist = gr.calibrate(lambda: gr.render(cams, planes[0]["rgba"].data_ptr(), planes[0]["id"].data_ptr(), planes[0]["t"].data_ptr()), rounds=2)
        gr.render(cams, planes[1]["rgba"].data_ptr(), planes[1]["id"].data_ptr(), planes[1]["t"].data_ptr())
        gr.finish()
        torch.cuda.synchronize()
        if rank == 0:
            ref = dict(rgba=torch.empty_like(planes[0]["rgba"]), id=torch.empty_like(planes[0]["id"]), t=torch.empty_like(planes[0]["t"]))
            scene.render_device(cams, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 0, H, ref["rgba"].data_ptr(), ref["id"].data_ptr(), ref["t"].data_ptr())
            torch.cuda.synchronize()
            for pi, p in enumerate(planes):
                for key in ("id", "t", "rgba"):
                    same = bool(torch.equal(p[key].view(torch.int32), ref[key].view(torch.int32)))
                    ok = ok and same
                    if not same:
                        msg += "plane set %d %s differs; " % (pi, key)
            ok = ok and float((ref["id"] >= 0).float().mean()) > 0.2
        want = transport if transport != "auto" else gr.transport
        ok = ok and gr.transport == want and len(hist) == 2 and len(hist[0]) == world
        q.put((rank, ok, gr.transport, msg))
        gr.close()
        dist.destroy_process_group()
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, False, "?", "worker %d failed: %r\n%s" % (rank, e, traceback.format_exc())))


@pytest.mark.parametrize("transport", ["auto", "nccl"])
@pytest.mark.parametrize("world", [2, 4])
def test_gathered_frames_equal_a_local_render(rto, world, transport):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, transport, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, ok, used, msg in sorted(res):
        assert ok, "rank %d (%s): %s" % (rank, used, msg)
    if transport == "auto":
        print("transport used:", res[0][2])
