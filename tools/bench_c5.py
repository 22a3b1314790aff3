#!/usr/bin/env python
"""BASELINE.json configs[4] ("C5"): 64-frame camera orbit of the DT mesh at 3840x2160, primary + shadow rays, frames sharded over
the GPUs of one box (scene replicated), framebuffers gathered to rank 0 over NCCL while the next batch is traced.

    python tools/bench_c5.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_c5.py

Prints one JSON line on rank 0: Mrays/s of the trace alone and of trace + gather (device time, max over ranks)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

W, H, FRAMES = 3840, 2160, 64
DT_GRID = os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4, help="frames per launch and per gather")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ray_tracing_octrees_b200 import build, sharding
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    import ray_tracing_octrees_b200 as rto
    assert rto.lib().rto_init(local) == 0
    g = rto.VoxelGrid.load(DT_GRID)
    nodes = rto.create_octree_from_voxel_grid(g)
    tris = rto.marching_cubes_mesh(g, nodes)
    scene = rto.Scene.bvh(tris)
    bias = float(np.float32(1e-3) * np.float32(g.voxel_size))
    aspect = float(np.float32(W) / np.float32(H))
    cam_of = lambda k: rto.Camera.from_degrees(35.0, 360.0 * k / FRAMES, 0.6 * 4250.0).consts(45.0, aspect, W, H)[0]
    mine = sharding.shard_frames(FRAMES, world, rank)
    B = a.batch
    assert len(mine) % B == 0, "frames per rank must be a multiple of --batch"
    nb = len(mine) // B
    trace_stream = torch.cuda.ExternalStream(scene.stream, device=torch.device("cuda", local))
    comm_stream = torch.cuda.Stream()
    # two plane sets so that batch b+1 is traced while batch b is gathered
    planes = [dict(rgba=torch.empty((B, H, W, 4), dtype=torch.float32, device="cuda"), id=torch.empty((B, H, W), dtype=torch.int32, device="cuda"),
                   t=torch.empty((B, H, W), dtype=torch.float32, device="cuda")) for _ in range(2)]
    full = None
    if rank == 0 and world > 1:
        full = dict(rgba=torch.empty((FRAMES, H, W, 4), dtype=torch.float32, device="cuda"), id=torch.empty((FRAMES, H, W), dtype=torch.int32, device="cuda"),
                    t=torch.empty((FRAMES, H, W), dtype=torch.float32, device="cuda"))

    def run(with_gather):
        done = [torch.cuda.Event(), torch.cuda.Event()]      # plane set free again (gather finished)
        traced = [torch.cuda.Event(), torch.cuda.Event()]
        hits = 0
        for b in range(nb):
            p = planes[b & 1]
            if b >= 2 and with_gather and world > 1:
                trace_stream.wait_event(done[b & 1])
            cams = [cam_of(k) for k in mine[b * B:(b + 1) * B]]
            scene.render_device(cams, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 0, H, p["rgba"].data_ptr(), p["id"].data_ptr(), p["t"].data_ptr())
            traced[b & 1].record(trace_stream)
            if with_gather and world > 1:
                with torch.cuda.stream(comm_stream):
                    comm_stream.wait_event(traced[b & 1])
                    for name in ("rgba", "id", "t"):
                        outs = sharding.gather_planes(p[name], dst=0)
                        if rank == 0:
                            for r, o in enumerate(outs):          # rank r's batch b holds global frames shard_frames(...)[b*B:(b+1)*B]
                                ks = sharding.shard_frames(FRAMES, world, r)[b * B:(b + 1) * B]
                                for i, k in enumerate(ks):
                                    full[name][k].copy_(o[i], non_blocking=True)
                    done[b & 1].record(comm_stream)
        torch.cuda.current_stream().wait_stream(comm_stream)

    def timed(with_gather):
        best = 1e30
        for _ in range(a.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(with_gather)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tv = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            best = min(best, float(tv.item()))
        return best

    run(False); torch.cuda.synchronize()                      # warm-up
    t_trace = timed(False)
    t_both = timed(True) if world > 1 else t_trace
    # rays: primaries + one shadow ray per primary hit, counted on every rank's last two batches and extrapolated per frame
    hits = torch.zeros(1, dtype=torch.float64, device="cuda")
    for b in range(nb):
        p = planes[b & 1]
        cams = [cam_of(k) for k in mine[b * B:(b + 1) * B]]
        scene.render_device(cams, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 0, H, p["rgba"].data_ptr(), p["id"].data_ptr(), p["t"].data_ptr())
        torch.cuda.synchronize()
        hits += (p["id"] >= 0).sum()
    if world > 1:
        dist.all_reduce(hits)
    rays = FRAMES * W * H + float(hits.item())
    ok = None
    if rank == 0 and world > 1:
        # frame 1 belongs to rank 1: render it here and compare with what the gather delivered
        p = planes[0]
        scene.render_device([cam_of(1)] * B, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 0, H, p["rgba"].data_ptr(), p["id"].data_ptr(), p["t"].data_ptr())
        torch.cuda.synchronize()
        ok = bool(torch.equal(p["id"][0], full["id"][1]) and torch.equal(p["t"][0], full["t"][1]) and torch.equal(p["rgba"][0], full["rgba"][1]))
    if rank == 0:
        print(json.dumps({"config": "C5: 64-frame DT orbit, 3840x2160, primary+shadow", "n_gpus": world, "frames_per_launch": B, "rays": rays,
                          "trace_s": t_trace, "trace_Mrays_s": rays / t_trace / 1e6, "trace_and_gather_s": t_both, "trace_and_gather_Mrays_s": rays / t_both / 1e6,
                          "gather_bytes_into_rank0": (world - 1) * (FRAMES // world) * W * H * 24 if world > 1 else 0,
                          "gathered_frame_equals_local_render": ok, "timing": "wall clock around synchronised device work, max over ranks, best of %d" % a.reps}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
