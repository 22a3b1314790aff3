#!/usr/bin/env python
"""bench.py -- headline benchmark of the ray-casting hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): the DT Calgary scene -- the reference's voxelised sceneCache.bin (the DT CSVs are absent
from the reference checkout, so its own voxel grid -> octree -> Marching-Cubes mesh, 487 832 triangles, is the stand-in, SURVEY.md
8d) -- ray cast through the reference-shaped BVH at 1920x1080, primary rays plus one shadow ray per primary hit, cameras on the
64-position orbit (theta 35 deg, radius 0.6 * 4250, phi_k = 40 + 360 k / 64 deg).

One "step" = F orbit frames PER GPU (default 128 = two orbits), traced in launches of BATCH (32) frames per GPU.
  N = 1   the frames are rendered straight into the frame planes (rgba32f + hit id + t, 24 B/pixel) in HBM.
  N > 1   scene replicated, the rows of every 16 x N-frame batch are dealt to the ranks, and EVERY frame is delivered as full planes
          in rank 0's HBM inside the timed region (BASELINE north_star: "only the framebuffer gather uses ... NVLink"): ranks != 0
          write 4-byte hit codes straight into rank 0's memory from the trace kernel (CUDA-IPC peer stores over NVLink), rank 0
          expands them into planes on a second stream (ray_tracing_octrees_b200/sharding.py GatheredRenderer; csrc/rto_group.cu is
          the same thing for one process).  The comm-free number (every rank keeps its frames) is reported beside it.

metric  Mrays/s = (primary + shadow rays of all frames of the K steps) / (max over ranks of the device time of the K steps).
value   inputs (scene, cameras) resident, outputs in HBM (N > 1: all in rank 0's HBM).
e2e     the same metric through the public host API: cameras go host->device and the three frame planes come back device->host
        into pinned buffers inside the timed region (rto_render_batch(RTO_MEM_HOST), 32 frames per call).
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
PHI0_DEG, PHI_STEP_DEG, ORBIT = 40.0, 360.0 / 64.0, 64
DT_GRID = os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz")
METRIC = "Mrays/s primary+shadow (DT mesh, 1080p)"
CONFIG = {"workload": "C2: DT Calgary mesh (Marching-Cubes soup of the reference's sceneCache.bin, 487832 triangles; the DT CSVs are absent "
                      "from the checkout), BVH ray cast at 1920x1080, primary + shadow rays, 64-camera orbit",
          "image": "1920x1080", "rays": "primary + one shadow ray per primary hit", "cameras": "orbit theta 35 deg, r 0.6*4250, phi 40 + 360k/64 deg"}
DATA = "the reference's own voxelised DT Calgary grid (sceneCache.bin, committed gzip-compressed as tests/golden/dt_sceneCache.bin.gz) -> octree -> Marching-Cubes mesh; cameras synthetic"
# frames per GPU per launch.  Measured on one 8-GPU box (third session of round 2; Grays/s at N = 8 with every frame on rank 0 / at N = 1):
# 16 frames 95.1 / 13.45, 32 frames 99.0 / 13.59, 64 frames 98.1 / 13.60 -- a batch is what the ranks meet for (one 4-byte all-reduce per
# batch), so fewer, larger batches cost less waiting; beyond 32 the larger plane sets on rank 0 take it back.
BATCH = int(os.environ.get("RTO_BENCH_BATCH", "32"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
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

    def step(k):       # per-step sample: one whole 1080p orbit frame (a fraction of a second on a many-core host)
        c, _ = chk.camera(THETA_DEG, PHI0_DEG + PHI_STEP_DEG * (k % ORBIT), RADIUS, width=W, height=H)
        out = mesh.render(c, 1, bias, 0, H, threads=cores)
        return H * W + int((out["id"] >= 0).sum()), out["sec"]
    for k in range(args.warmup):
        step(k)
    rays, secs = 0, 0.0
    for k in range(args.steps):
        r, s = step(args.warmup + k)
        rays += r
        secs += s
    v = rays / secs / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": DATA,
            "config": CONFIG, "step": "one full 1080p orbit frame (bounded sample of the workload; the CUDA arm traces %d frames per GPU per step)" % args.frames,
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
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ray_tracing_octrees_b200 import build
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    import ray_tracing_octrees_b200 as rto
    from ray_tracing_octrees_b200 import sharding
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
    F = max(BATCH, (args.frames // BATCH) * BATCH)
    chunks = F // BATCH
    stream = torch.cuda.ExternalStream(scene.stream, device=dev)
    aspect = float(np.float32(W) / np.float32(H))
    orbit = [rto.Camera.from_degrees(THETA_DEG, PHI0_DEG + PHI_STEP_DEG * k, RADIUS).consts(FOV, aspect, W, H)[0] for k in range(ORBIT)]
    cam_cache = {}

    def cam_array(start, n):
        """ctypes array of the n orbit cameras from global frame index `start` (built once per distinct phase: nothing per step).
        The frames of a batch are listed in a strided order (j -> start + j * s mod n, s odd ~ 0.38 n, a permutation because n is a
        multiple of 16 and a power of two here): the cost of a frame varies smoothly along the orbit, and with this order every
        contiguous share of a batch samples the whole arc, so the shares of the ranks cost the same in EVERY batch."""
        key = (start % ORBIT, n)
        if key not in cam_cache:
            s = max(1, int(0.382 * n)) | 1
            while np.gcd(s, n) != 1:
                s += 2
            cam_cache[key] = (rto.RtoCamera * n)(*[orbit[(start + (j * s) % n) % ORBIT] for j in range(n)])
        return cam_cache[key]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # primary hits of every orbit camera (rays per frame = W*H + hits), counted once outside every timed region
    tmp_id = torch.empty((BATCH, H, W), dtype=torch.int32, device=dev)
    hits_of = []
    for k0 in range(0, ORBIT, BATCH):
        scene.render_device((rto.RtoCamera * BATCH)(*orbit[k0:k0 + BATCH]), rto.MODE_BVH, flags, bias, 0, H, None, tmp_id.data_ptr(), None)
        torch.cuda.synchronize()
        hits_of += [int(x) for x in (tmp_id >= 0).flatten(1).sum(1).tolist()]
    del tmp_id
    rays_of = [W * H + h for h in hits_of]

    def rays_in(start, n):
        return sum(rays_of[(start + j) % ORBIT] for j in range(n))

    # ---- planes: a ring of two batches (a batch is BATCH frames per GPU; N > 1: all of them on rank 0) --------------------------
    gathered = world > 1 and not args.comm_free_only
    nb = BATCH * world if (gathered and rank == 0) else BATCH
    ring = [dict(rgba=torch.empty((nb, H, W, 4), dtype=torch.float32, device=dev), id=torch.empty((nb, H, W), dtype=torch.int32, device=dev),
                 t=torch.empty((nb, H, W), dtype=torch.float32, device=dev)) for _ in range(2)]
    gr = None
    if gathered:
        gr = sharding.GatheredRenderer(rto, scene, W, H, BATCH * world, flags, bias, dev, transport=os.environ.get("RTO_GATHER_TRANSPORT", "auto"))

    def step_local(k, rank_offset=True):
        """comm-free: this rank's F frames of step k straight into its own planes."""
        base = (k * world + (rank if rank_offset else 0)) * F
        for c in range(chunks):
            p = ring[c & 1]
            scene.render_device(cam_array(base + c * BATCH, BATCH), rto.MODE_BVH, flags, bias, 0, H, p["rgba"].data_ptr(), p["id"].data_ptr(), p["t"].data_ptr())

    def step_gathered(k):
        """all ranks trace, every frame of the step ends up as planes on rank 0."""
        base = k * world * F
        for c in range(chunks):
            p = ring[c & 1]
            gr.render(cam_array(base + c * BATCH * world, BATCH * world), p["rgba"].data_ptr(), p["id"].data_ptr(), p["t"].data_ptr())

    def timed(step_fn, steps, finish=None):
        sync_all()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        end = torch.cuda.Event(enable_timing=True)
        for k in range(steps):
            step_fn(args.warmup + k)
            ev[k + 1].record(stream)              # (per-step marks; on rank 0 the expansion of a step's last batch may still run beside the next step)
        if finish:
            finish()                              # rank 0: everything expanded
        end.record(stream)
        sync_all()
        per = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        return ev[0].elapsed_time(end), per

    # ---- calibration of the shares (N > 1) + warm-up ----------------------------------------------------------------------------
    cal = None
    if gathered:
        p = ring[0]

        def four_batches():       # back to back, so that rank 0's trace is measured with the expansion of the previous batch beside it
            for c in range(4):
                q = ring[c & 1]
                gr.render(cam_array(c * BATCH * world, BATCH * world), q["rgba"].data_ptr(), q["id"].data_ptr(), q["t"].data_ptr())
        cal = gr.calibrate(four_batches, rounds=8)
    main_step = step_gathered if gathered else step_local
    main_finish = gr.finish if gathered else None
    for k in range(args.warmup):
        main_step(k)
    if main_finish:
        main_finish()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.2)
    launches0 = scene.launch_count
    dev_ms, per_step = timed(main_step, args.steps, main_finish)
    launches = scene.launch_count - launches0
    clocks = sampler.stop()
    rays_main = sum(rays_in((args.warmup + k) * world * F, world * F) for k in range(args.steps)) if gathered else \
        sum(rays_in(((args.warmup + k) * world + rank) * F, F) for k in range(args.steps))

    # ---- N > 1: the gathered planes equal a local render (last batch of the last step), and the comm-free number beside it ------
    gather_check, comm_free = None, None
    if gathered:
        if rank == 0:
            last = (args.warmup + args.steps - 1) * world * F + (chunks - 1) * BATCH * world
            got = ring[(chunks - 1) & 1]
            ref = dict(rgba=torch.empty_like(got["rgba"][:BATCH]), id=torch.empty_like(got["id"][:BATCH]), t=torch.empty_like(got["t"][:BATCH]))
            same = True
            batch_cams = cam_array(last, BATCH * world)
            for j in range(0, BATCH * world, BATCH):
                sub = (rto.RtoCamera * BATCH).from_buffer(batch_cams, j * ctypes.sizeof(rto.RtoCamera))
                scene.render_device(sub, rto.MODE_BVH, flags, bias, 0, H, ref["rgba"].data_ptr(), ref["id"].data_ptr(), ref["t"].data_ptr())
                torch.cuda.synchronize()
                for key in ("rgba", "id", "t"):
                    same = same and bool(torch.equal(got[key][j:j + BATCH].view(torch.int32), ref[key].view(torch.int32)))
            gather_check = "bit-equal to a local render (%d frames)" % (BATCH * world) if same else "MISMATCH"
            del ref
        cf_ms, _ = timed(step_local, min(args.steps, 10))
        comm_free = (cf_ms, sum(rays_in(((args.warmup + k) * world + rank) * F, F) for k in range(min(args.steps, 10))))

    # ---- kernel duration for the roofline (CUDA events on the launching stream around one launch of BATCH frames) ------------------
    kern_ms = []
    p = ring[0]
    for k in range(7):
        scene.render_device(cam_array(k * BATCH, BATCH), rto.MODE_BVH, flags, bias, 0, H, p["rgba"].data_ptr(), p["id"].data_ptr(), p["t"].data_ptr())
        kern_ms.append(scene.last_kernel_ms())
    kern_ms_avg = float(np.mean(kern_ms[2:6]))          # launches 2..5: the whole orbit (twice at 32 frames per launch)

    # ---- e2e: host API, pinned host buffers, H2D cameras + D2H frame planes inside the timed region ---------------------------
    EB = min(args.e2e_batch, BATCH) if args.e2e_batch > 0 and F % min(args.e2e_batch, BATCH) == 0 else 8      # frames per host call: the first kernel of a call has no copy to hide behind
    h_rgba = torch.empty((EB, H, W, 4), dtype=torch.float32).pin_memory()
    h_id = torch.empty((EB, H, W), dtype=torch.int32).pin_memory()
    h_t = torch.empty((EB, H, W), dtype=torch.float32).pin_memory()

    def step_host(k):
        base = (k * world + rank) * F
        for c in range(F // EB):
            scene.render_host_ptrs(cam_array(base + c * EB, EB), rto.MODE_BVH, flags, bias, 0, H, h_rgba.data_ptr(), h_id.data_ptr(), h_t.data_ptr())

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    scene.render_host_ptrs(cam_array(0, EB), rto.MODE_BVH, flags, bias, 0, H, h_rgba.data_ptr(), h_id.data_ptr(), h_t.data_ptr())
    sync_all()
    te = time.perf_counter()
    for k in range(e2e_steps):
        step_host(args.warmup + k)                      # synchronous: returns when the planes are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - te
    e2e_rays = sum(rays_in(((args.warmup + k) * world + rank) * F, F) for k in range(e2e_steps))
    lastb = ((args.warmup + e2e_steps - 1) * world + rank) * F + F - EB
    assert int((h_id >= 0).sum().item()) == sum(hits_of[(lastb + j) % ORBIT] for j in range(EB)), "host planes differ from device planes"
    # the same call asking for hit id + t only (8 B/pixel): what a caller that shades on its own side, or only needs visibility, moves
    def step_host_idt(k):
        base = (k * world + rank) * F
        for c in range(F // EB):
            scene.render_host_ptrs(cam_array(base + c * EB, EB), rto.MODE_BVH, flags, bias, 0, H, None, h_id.data_ptr(), h_t.data_ptr())
    sync_all()
    te = time.perf_counter()
    step_host_idt(args.warmup)
    torch.cuda.synchronize()
    e2e_idt_s = time.perf_counter() - te
    e2e_idt_rays = rays_in((args.warmup * world + rank) * F, F)
    # ... and asking for the framebuffer only: the RGBA32F image is all RayTracerBVH::renderSceneCompute itself produces (16 B/pixel)
    def step_host_rgba(k):
        base = (k * world + rank) * F
        for c in range(F // EB):
            scene.render_host_ptrs(cam_array(base + c * EB, EB), rto.MODE_BVH, flags, bias, 0, H, h_rgba.data_ptr(), None, None)
    sync_all()
    te = time.perf_counter()
    step_host_rgba(args.warmup)
    torch.cuda.synchronize()
    e2e_rgba_s = time.perf_counter() - te
    # the ceiling of the link: all ranks copy device -> pinned host at the same time, nothing else running
    d_src = ring[0]["rgba"].view(-1)[:h_rgba.numel()]
    sync_all()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(6):
        h_rgba.view(-1).copy_(d_src, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    d2h_gbps = 6 * h_rgba.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    if world > 1:
        dist.barrier()

    # ---- reduce over ranks: max time, sum rays ---------------------------------------------------------------
    tvec = torch.tensor([dev_ms, e2e_s, kern_ms_avg, comm_free[0] if comm_free else 0.0, e2e_idt_s, e2e_rgba_s], dtype=torch.float64, device=dev)
    rvec = torch.tensor([0 if gathered else rays_main, e2e_rays, launches, comm_free[1] if comm_free else 0.0, e2e_idt_rays, d2h_gbps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(rvec, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max, kern_ms_max, cf_ms_max, e2e_idt_s_max, e2e_rgba_s_max = [float(x) for x in tvec.tolist()]
    rays_all, e2e_rays_all, launches_all, cf_rays_all, e2e_idt_rays_all, d2h_gbps_all = [float(x) for x in rvec.tolist()]
    if gathered:
        rays_all = float(rays_main)                     # every rank counted the same global frames

    if rank != 0:
        if gr:
            gr.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_render_bvh<shadows, pruned>), one launch = BATCH frames ------------------------------
    # Algorithmic bytes/flops per ray are defined on the REFERENCE's structures: B box tests, C candidates, counted on
    # the GPU by replaying the reference's visit pattern (rto_render_stats, checked against the oracle's counters in the tests):
    #   primary ray: 24 B + 36 C + 24 bytes, 18 B + 51 C + 60 flops;  shadow ray: 24 Bs + 36 Cs bytes, 18 Bs + 51 Cs flops.
    # The orbit's 64 cameras are sampled every 4th and scaled to the BATCH cameras of an average launch.
    st = np.zeros(5, np.float64)
    sampled = list(range(0, ORBIT, 4))
    for k in sampled:
        st += scene.stats(orbit[k], rto.MODE_BVH, flags, bias).astype(np.float64)
    st *= BATCH / float(len(sampled))              # counters of the BATCH cameras of an average launch
    prim = BATCH * W * H
    alg_bytes = 24 * st[0] + 36 * st[1] + 24 * prim + 24 * st[2] + 36 * st[3]
    alg_flops = 18 * st[0] + 51 * st[1] + 60 * prim + 18 * st[2] + 51 * st[3]
    hbm_peak, peak_src, sm_max = load_peaks()
    sms = ctypes.c_int()
    rto.lib().rto_device_info(ctypes.byref(sms), None, None, None, None)
    fp32_peak = sms.value * 128 * 2 * sm_max * 1e6 / 1e12
    ksec = kern_ms_avg * 1e-3
    counters = {}
    cp = os.path.join(ROOT, "profiles", "r02_bvh_counters.json")
    if os.path.exists(cp):
        try:
            counters = json.load(open(cp))
        except Exception:
            counters = {}
    traffic = counters.get("dram_bytes_per_launch")
    if traffic and ("%d)" % BATCH) not in str(counters.get("grid", "")).replace(" ", ""):
        traffic = None                              # the capture was taken with another number of frames per launch: not this launch's traffic
    roofline = {
        # neither HBM nor the tensor cores bound this kernel (scene L1/L2-resident, no contraction): the limit it is nearest to is FP32
        # instruction issue, so that is the one `frac` is quoted against (<= 1 by construction); the other limits are listed beside it.
        "bound": "fp32-issue", "achieved": alg_flops / ksec / 1e12, "peak": fp32_peak, "unit": "TFLOP/s", "frac": alg_flops / ksec / 1e12 / fp32_peak,
        "traffic": traffic, "kernel": "k_render_bvh<shadows,pruned>", "kernel_ms": kern_ms_avg, "frames_per_launch": BATCH,
        "peak_source": "SMs x 128 lanes x 2 x sm_max_mhz (%d SMs, %.0f MHz)" % (sms.value, sm_max),
        "definition": "algorithmic flops of the REFERENCE's structures (SURVEY.md 8d) per launch / kernel time; the production tree does fewer box and triangle tests than counted",
        "hbm": {"achieved_GBps": (traffic / ksec / 1e9) if traffic else None, "peak_GBps": hbm_peak, "frac": (traffic / ksec / 1e9 / hbm_peak) if traffic else None,
                "peak_source": peak_src, "compulsory_bytes_per_launch": prim * 24, "source": "profiles/r02_bvh_counters.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None},
        # utilisations ncu measured for this kernel (percent of each unit's own peak): the two highest -- L1/TEX throughput and issue slots --
        # are what the kernel is bound by; L2 and DRAM are far from theirs
        "ncu": {k: counters.get(k) for k in ("l1_throughput_pct", "issue_active_pct", "alu_pipe_pct", "fma_pipe_pct", "lsu_pipe_pct", "lanes_per_instruction",
                                             "long_scoreboard_warps_per_issue", "l1_hit_pct", "l2_hit_pct", "l2_throughput_pct", "dram_throughput_pct",
                                             "warp_instructions", "source")} if counters else None,
        "algorithmic_intensity": {"bytes_per_launch": alg_bytes, "GBps": alg_bytes / ksec / 1e9, "flops_per_launch": alg_flops,
                                  "note": "reference-defined bytes, served from L1/L2; not a fraction of any limit"},
        "per_primary_ray": {"box_tests": st[0] / prim, "candidates": st[1] / prim}, "shadow_rays_per_launch": st[4]}

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
    run = ({"frames_per_step_per_gpu": F, "frames_per_launch_per_gpu": BATCH, "triangles": int(len(tris)), "bvh_nodes": int(host_bvh.num_nodes),
                "l2": "every launch writes %.0f MB of planes per GPU through L2 (126 MB), two plane sets alternate; scene resident by design" % (BATCH * W * H * 24 / 1e6),
                "parallelism": ("rows of every %d-frame batch dealt to %d GPUs, scene replicated, all frames gathered on rank 0 (%s)" % (BATCH * world, world, gr.transport)) if gathered
                else "frames sharded over %d GPU(s), scene replicated, no exchange" % world, "scene_build_s": build_s})
    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "ms_per_step_median": float(np.median(per_step)), "timed_region_s": dev_ms_max * 1e-3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": DATA, "config": CONFIG, "run": run, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": F * ctypes.sizeof(rto.RtoCamera), "d2h_bytes_per_step": F * W * H * 24,
                    "steps": e2e_steps, "api": "rto_render_batch(RTO_MEM_HOST), %d frames per call into pinned host planes; per-frame D2H overlaps the next frame's kernel" % EB,
                    "d2h_GBps": world * e2e_steps * F * W * H * 24 / e2e_s_max / 1e9,
                    "d2h_ceiling_GBps": d2h_gbps_all, "ceiling": "all %d rank(s) copying 530 MB device -> pinned host six times concurrently, nothing else running (sum over ranks)" % world,
                    "id_t_only": {"value": e2e_idt_rays_all / e2e_idt_s_max / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": F * W * H * 8, "steps": 1,
                                  "what": "the same call with rgba = NULL: hit id + t planes only (8 B/pixel)"},
                    "rgba_only": {"value": e2e_idt_rays_all / e2e_rgba_s_max / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": F * W * H * 16, "steps": 1,
                                  "what": "the same call with hitId = t = NULL: the RGBA32F framebuffer alone, which is all the reference's renderSceneCompute produces (16 B/pixel)"}},
            "gpu_launches": int(launches_all), "roofline": roofline, "cpu_baseline": cpu}
    if gathered:
        line["gather"] = {"what": "every frame delivered as rgba32f + id + t planes in rank 0's HBM inside the timed region", "transport": gr.transport,
                          "bytes_per_pixel_on_the_wire": 4, "check": gather_check, "shares": [round(w / sum(gr.weights), 4) for w in gr.weights],
                          "calibration_ms": [[round(x, 3) for x in r] for r in (cal or [])][-2:],
                          "into_rank0_GBps": (1.0 - gr.weights[0] / sum(gr.weights)) * world * F * args.steps * W * H * 4 / (dev_ms_max * 1e-3) / 1e9}
        line["comm_free"] = {"value": cf_rays_all / (cf_ms_max * 1e-3) / 1e6, "unit": "Mrays/s", "steps": min(args.steps, 10),
                             "what": "the same frames, every rank keeps its own planes (no exchange)"}
    print(json.dumps(line), flush=True)
    if gr:
        gr.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=128, help="orbit frames per step per GPU (multiple of 16)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-batch", type=int, default=32, help="frames per rto_render_batch(RTO_MEM_HOST) call of the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--comm-free-only", action="store_true", help="N > 1: skip the gather (round-1 behaviour)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
