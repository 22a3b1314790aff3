#!/usr/bin/env python
"""bench.py -- headline benchmark of the ray-casting hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): the DT Calgary scene -- the reference's voxelised sceneCache.bin (the DT
CSVs are absent from the reference checkout, so its own voxel grid -> octree -> Marching-Cubes mesh, 487 832
triangles, is the stand-in) -- ray cast through the reference-shaped BVH at 1920x1080, primary rays
plus one shadow ray per primary hit.  One "step" = one batch of F orbit frames per GPU (camera theta 35 deg,
radius 0.6*4250, phi advancing 360/64 deg per frame; ranks take disjoint phi ranges: frames shard, scene replicated).

metric  Mrays/s = (primary + shadow rays traced by all ranks) / (max over ranks of the device time of the K steps).
value   inputs (scene, cameras) resident in HBM, outputs written to HBM.
e2e     the same metric through the public host API: cameras go host->device and the three frame planes
        (rgba32f, hit id, t: 24 B/pixel) come back device->host into pinned buffers inside the timed region.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
THETA_DEG, RADIUS, FOV = 35.0, 0.6 * 4250.0, 45.0
PHI0_DEG, PHI_STEP_DEG = 40.0, 360.0 / 64.0
DT_GRID = os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz")
METRIC = "Mrays/s primary+shadow (DT mesh, 1080p)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    """All host threads this process may use.  (torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently run the
    CPU reference on one thread; the count is therefore taken from the affinity mask and passed to the oracle explicitly.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def orbit_cameras(rto, first_frame, count):
    cams = []
    aspect = float(np.float32(W) / np.float32(H))
    for k in range(first_frame, first_frame + count):
        cam, _ = rto.Camera.from_degrees(THETA_DEG, PHI0_DEG + PHI_STEP_DEG * (k % 64), RADIUS).consts(FOV, aspect, W, H)
        cams.append(cam)
    return cams


# =====================================================================================================
# reference arm: the reference's own CPU implementation (compiled in place -> oracle/_ref), else the port
# =====================================================================================================
def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import bind
    chk = bind.best()
    oc = chk.octree(*bind.load_scene_cache(DT_GRID))
    oc.build()
    mesh = oc.mesh()
    mesh.build()
    cores = host_cores()
    bias = 1e-3 * oc.voxel
    cam, _ = chk.camera(THETA_DEG, PHI0_DEG, RADIUS, width=W, height=H)
    # per-step sample: one whole 1080p orbit frame (a fraction of a second on a many-core host)
    rows, y0 = H, 0

    def step(k):
        c, _ = chk.camera(THETA_DEG, PHI0_DEG + PHI_STEP_DEG * (k % 64), RADIUS, width=W, height=H)
        out = mesh.render(c, 1, bias, y0, y0 + rows, threads=cores)
        return rows * W + int((out["id"] >= 0).sum()), out["sec"]
    for k in range(args.warmup):
        step(k)
    rays = 0
    secs = 0.0
    for k in range(args.steps):
        r, s = step(args.warmup + k)
        rays += r
        secs += s
    v = rays / secs / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 DT-voxel MC mesh 487832 tris, BVH ray cast 1920x1080 primary+shadow", "sample": "one full 1080p orbit frame per step (ours: %d frames per step)" % args.frames},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": chk.kind,
                             "sample": "%d full 1080p orbit frames, BVH::query + Moller-Trumbore + shadow, OpenMP over scanlines" % args.steps},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# =====================================================================================================
# our arm
# =====================================================================================================
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ray_tracing_octrees_b200 import build
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    import ray_tracing_octrees_b200 as rto
    rc = rto.lib().rto_init(local_rank)
    if rc != 0:
        raise SystemExit("rto_init failed: " + rto.lib().rto_last_error().decode())

    t0 = time.time()
    grid = rto.VoxelGrid.load(DT_GRID)
    nodes = rto.create_octree_from_voxel_grid(grid)
    tris = rto.marching_cubes_mesh(grid, nodes)
    host_bvh = rto.HostBVH(tris)
    scene = rto.Scene.bvh(None, prebuilt=host_bvh)
    build_s = time.time() - t0
    bias = float(np.float32(1e-3) * np.float32(grid.voxel_size))
    flags = rto.FLAG_SHADOWS
    F = args.frames
    stream = torch.cuda.ExternalStream(scene.stream, device=torch.device("cuda", local_rank))
    rgba = torch.empty((F, H, W, 4), dtype=torch.float32, device="cuda")
    hid = torch.empty((F, H, W), dtype=torch.int32, device="cuda")
    tt = torch.empty((F, H, W), dtype=torch.float32, device="cuda")

    def step_device(k):
        cams = orbit_cameras(rto, (rank * args.steps_total + k) * F, F)
        scene.render_device(cams, rto.MODE_BVH, flags, bias, 0, H, rgba.data_ptr(), hid.data_ptr(), tt.data_ptr())
        return cams

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident ----------------------------------------------------------------------
    for k in range(args.warmup):
        step_device(k)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = scene.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays = 0
    kern_ms = []
    sync_all()
    ev0.record(stream)
    for k in range(args.steps):
        step_device(args.warmup + k)
        if args.per_launch_timing:
            kern_ms.append(scene.last_kernel_ms())
    ev1.record(stream)
    sync_all()
    dev_ms = ev0.elapsed_time(ev1)
    launches = scene.launch_count - launches0
    clocks = sampler.stop()
    # rays actually traced: primaries + one shadow ray per primary hit (counted on the device, outside the timed region)
    rays_per_step = []
    for k in range(args.steps):
        step_device(args.warmup + k)
        torch.cuda.synchronize()
        rays_per_step.append(F * W * H + int((hid >= 0).sum().item()))
    rays = sum(rays_per_step)

    # ---- kernel duration for the roofline (CUDA events on the launching stream around each launch) ------------
    if not kern_ms:
        for k in range(min(args.steps, 5)):
            step_device(args.warmup + k)
            kern_ms.append(scene.last_kernel_ms())
    kern_ms_avg = float(np.mean(kern_ms))

    # ---- e2e: host API, pinned host buffers, H2D cameras + D2H frame planes inside the timed region ------------
    h_rgba = torch.empty((F, H, W, 4), dtype=torch.float32).pin_memory()
    h_id = torch.empty((F, H, W), dtype=torch.int32).pin_memory()
    h_t = torch.empty((F, H, W), dtype=torch.float32).pin_memory()

    def step_host(k):
        cams = orbit_cameras(rto, (rank * args.steps_total + k) * F, F)
        scene.render_host_ptrs(cams, rto.MODE_BVH, flags, bias, 0, H, h_rgba.data_ptr(), h_id.data_ptr(), h_t.data_ptr())

    e2e_steps = max(1, min(args.steps, 10))
    for k in range(2):
        step_host(k)
    sync_all()
    te = time.perf_counter()
    for k in range(e2e_steps):
        step_host(args.warmup + k)                      # synchronous: returns when the planes are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - te
    e2e_rays = sum(rays_per_step[:e2e_steps])            # same frames as the device-timed steps
    assert int((h_id >= 0).sum().item()) + F * W * H == rays_per_step[e2e_steps - 1], "host planes differ from device planes"
    if world > 1:
        dist.barrier()

    # ---- reduce over ranks: max time, sum rays ---------------------------------------------------------------
    tvec = torch.tensor([dev_ms, e2e_s, kern_ms_avg], dtype=torch.float64, device="cuda")
    rvec = torch.tensor([rays, e2e_rays, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(rvec, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max, kern_ms_max = [float(x) for x in tvec.tolist()]
    rays_all, e2e_rays_all, launches_all = [float(x) for x in rvec.tolist()]

    # ---- optional: NCCL framebuffer gather to rank 0 (the one collective of the path), timed separately --------
    gather = None
    if world > 1:
        from ray_tracing_octrees_b200 import sharding
        sharding.gather_planes(hid, dst=0)                      # warm up the communicator
        torch.cuda.synchronize(); dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for plane in (rgba, hid, tt):
            sharding.gather_planes(plane, dst=0)
        g1.record(); torch.cuda.synchronize()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        nbytes = rgba.numel() * 4 + hid.numel() * 4 + tt.numel() * 4
        gather = {"what": "all three planes of one step (F frames) from every rank to rank 0 over NCCL", "bytes_per_rank": nbytes,
                  "ms": float(gms.item()), "GBps_into_rank0": nbytes * (world - 1) / (float(gms.item()) * 1e-3) / 1e9}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_render_bvh<shadows, pruned>) --------------------------------------
    # Algorithmic bytes/flops per ray are defined on the REFERENCE's structures: B box tests, C candidates,
    # counted on the GPU by replaying the reference's visit pattern (rto_render_stats, checked against the oracle in tests):
    #   primary ray: 24 B + 36 C + 24 bytes, 18 B + 51 C + 60 flops;  shadow ray: 24 Bs + 36 Cs bytes, 18 Bs + 51 Cs flops.
    cams = orbit_cameras(rto, args.warmup * F, F)
    st = np.zeros(5, np.float64)
    for c in cams:
        st += scene.stats(c, rto.MODE_BVH, flags, bias).astype(np.float64)
    prim = F * W * H
    alg_bytes = 24 * st[0] + 36 * st[1] + 24 * prim + 24 * st[2] + 36 * st[3]
    alg_flops = 18 * st[0] + 51 * st[1] + 60 * prim + 18 * st[2] + 51 * st[3]
    peak, peak_src, sm_max = load_peaks()
    achieved = alg_bytes / (kern_ms_avg * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_render_bvh_bytes_per_launch")
        except Exception:
            traffic = None
    sms = ctypes.c_int()
    rto.lib().rto_device_info(ctypes.byref(sms), None, None, None, None)
    fp32_peak = sms.value * 128 * 2 * sm_max * 1e6 / 1e12
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src + " MEASURED_PEAKS.json hbm_gbs (burst copy)", "kernel": "k_render_bvh<shadows,pruned>",
                "kernel_ms": kern_ms_avg, "algorithmic_bytes_per_launch": alg_bytes,
                "per_primary_ray": {"box_tests": st[0] / prim, "candidates": st[1] / prim}, "shadow_rays_per_launch": st[4],
                "fp32": {"achieved_tflops": alg_flops / (kern_ms_avg * 1e-3) / 1e12, "peak_tflops": fp32_peak, "frac": alg_flops / (kern_ms_avg * 1e-3) / 1e12 / fp32_peak},
                "note": "scene (~64 MB) is L2-resident by design, so algorithmic bytes are served mostly from L2/L1; frac>1 of HBM is expected"}

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample on the host cores ---------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import bind
        chk = bind.best()
        oc = chk.octree(*bind.load_scene_cache(DT_GRID))
        oc.build()
        mesh = oc.mesh()
        mesh.build()
        cores = host_cores()
        crays, csec, cframes = 0, 0.0, 0
        while csec < 10.0 and cframes < 64:       # whole 1080p orbit frames until ~10 s of wall time on all host cores (or the whole orbit)
            ccam, _ = chk.camera(THETA_DEG, PHI0_DEG + PHI_STEP_DEG * cframes, RADIUS, width=W, height=H)
            out = mesh.render(ccam, 1, bias, 0, H, threads=cores)
            crays += W * H + int((out["id"] >= 0).sum())
            csec += out["sec"]
            cframes += 1
        ccam, _ = chk.camera(THETA_DEG, PHI0_DEG, RADIUS, width=W, height=H)
        one = mesh.render(ccam, 1, bias, H // 2 - 32, H // 2 + 32, threads=1)
        cpu = {"value": crays / csec / 1e6, "unit": "Mrays/s", "cores": cores, "kind": chk.kind,
               "sample": "%d full 1080p orbit frames (%.2f s wall), BVH::query + Moller-Trumbore + shadow rays, OpenMP over scanlines" % (cframes, csec),
               "one_thread_value": (64 * W + int((one["id"] >= 0).sum())) / one["sec"] / 1e6}

    value = rays_all / (dev_ms_max * 1e-3) / 1e6
    e2e_value = e2e_rays_all / e2e_s_max / 1e6
    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "C2 DT-voxel MC mesh 487832 tris (sceneCache.bin stand-in for the absent DT CSVs), BVH ray cast 1920x1080 primary+shadow",
                       "frames_per_step_per_gpu": F, "triangles": int(len(tris)), "bvh_nodes": int(host_bvh.num_nodes),
                       "l2": "outputs %.0f MB/step written through L2 (126 MB) between repeats; scene resident by design" % (F * W * H * 24 / 1e6),
                       "parallelism": "frames sharded over %d GPU(s), scene replicated" % world, "scene_build_s": build_s},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": F * ctypes.sizeof(rto.RtoCamera), "d2h_bytes_per_step": F * W * H * 24,
                    "steps": e2e_steps, "api": "rto_render_batch(RTO_MEM_HOST) into pinned host planes; per-frame D2H overlaps the next frame's kernel"},
            "gpu_launches": int(launches_all),
            "roofline": roofline, "cpu_baseline": cpu}
    if gather:
        line["nccl_gather"] = gather
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=8, help="orbit frames per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-launch-timing", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.steps_total = args.steps + args.warmup
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
