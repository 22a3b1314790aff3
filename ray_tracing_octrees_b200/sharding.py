"""Multi-GPU host logic of the path: the scene is replicated on every rank, work shards with no data-path collective
(pixels and frames are independent), and the only exchange is the framebuffer gather to rank 0.

Works with any torch.distributed backend: NCCL on the B200 box (device tensors over NVLink), gloo in the CPU tests.
"""
import numpy as np


def shard_frames(num_frames, world, rank):
    """Frames of a camera batch dealt round-robin (frame k -> rank k % world): neighbouring orbit angles cost about the same,
    so round-robin balances better than contiguous blocks.  Returns the list of global frame indices of `rank`."""
    return list(range(rank, num_frames, world))


def shard_rows(height, world, rank, band=8):
    """Interleaved scanline bands of ONE frame: band b (rows [b*band, (b+1)*band)) -> rank b % world.  Sky rows are cheap and
    skyline rows expensive, so interleaving keeps ranks balanced.  Returns [(y0, y1), ...] for `rank` (possibly empty)."""
    out = []
    for b, y0 in enumerate(range(0, height, band)):
        if b % world == rank:
            out.append((y0, min(height, y0 + band)))
    return out


def assemble_rows(height, width, world, parts, band=8, channels=None):
    """Inverse of shard_rows: parts[r] is rank r's planes concatenated in band order -> the full (height, width[, C]) plane."""
    first = next(p for p in parts if p is not None and len(p))
    shape = (height, width) if channels is None else (height, width, channels)
    full = np.empty(shape, dtype=np.asarray(first).dtype)
    for r in range(world):
        off = 0
        for (y0, y1) in shard_rows(height, world, r, band):
            n = y1 - y0
            full[y0:y1] = np.asarray(parts[r])[off:off + n]
            off += n
    return full


def gather_planes(plane, dst=0, group=None):
    """Framebuffer gather: every rank contributes a tensor of identical shape, rank `dst` gets the list (others None).
    `dst` is a rank OF `group` (the default group's ranks are the global ones); torch's gather wants the global rank.
    One grouped NCCL/gloo gather; the caller may run it on a side stream to overlap the next batch."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return [plane]
    outs = [torch.empty_like(plane) for _ in range(world)] if rank == dst else None
    dist.gather(plane, outs, dst=dst if group is None else dist.get_global_rank(group, dst), group=group)
    return outs


def interleave_frames(per_rank, num_frames, world):
    """Inverse of shard_frames: per_rank[r][i] is rank r's i-th frame -> frames in global order."""
    out = [None] * num_frames
    for r in range(world):
        for i, k in enumerate(shard_frames(num_frames, world, r)):
            out[k] = per_rank[r][i]
    return out


# ---------------------------------------------------------------------------------------------------------------------------
# Weighted row ranges + the gather of hit codes: the multi-process mirror of csrc/rto_group.cu (one process per GPU)
# ---------------------------------------------------------------------------------------------------------------------------
def deal_tiles(tiles, weights):
    """Cut [0, tiles) into len(weights) contiguous ranges proportional to the weights -> list of len(weights) + 1 cut points."""
    total = float(sum(weights))
    cuts, acc = [0], 0.0
    for i, w in enumerate(weights):
        acc += w
        c = tiles if i + 1 == len(weights) else int(round(acc / total * tiles))
        cuts.append(min(tiles, max(cuts[-1], c)))
    return cuts


def split_rows(g0, g1, height):
    """Rows [g0, g1) of a batch's row space (frame f, row y <-> f * height + y) as at most three (frame0, frames, y0, y1) launches:
    a partial first frame, whole frames, a partial last frame."""
    if g1 <= g0:
        return []
    f0, f1 = g0 // height, (g1 - 1) // height
    ya, yb = g0 - f0 * height, g1 - f1 * height
    if f0 == f1:
        return [(f0, 1, ya, yb)]
    out = []
    if ya != 0:
        out.append((f0, 1, ya, height))
        f0 += 1
    full_end = f1 + 1 if yb == height else f1
    if full_end > f0:
        out.append((f0, full_end - f0, 0, height))
    if yb != height:
        out.append((f1, 1, 0, yb))
    return out


def tile_row_to_row(tile, height):
    """First image row (in the batch's row space) of 8-row tile row `tile`; tile == frames * tiles_per_frame gives frames * height."""
    tpf = (height + 7) // 8
    return (tile // tpf) * height + (tile % tpf) * 8


def rebalance(weights, times, gain=0.7, floor=0.02):
    """New shares from the measured time of every rank: a rank that took longer than the mean gets less."""
    mean = float(sum(times)) / len(times)
    w = [wi * (mean / max(ti, 1e-9)) ** gain for wi, ti in zip(weights, times)]
    s = sum(w)
    return [max(floor, wi * len(w) / s) for wi in w]


class _CudaBuffer:
    """Lets torch see a raw device allocation (an ExchangeBuffer) as a tensor."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr, False), "version": 2}


class GatheredRenderer:
    """Frames of a batch traced by all ranks, delivered as full planes on rank 0.

    Every rank holds a replica of the BVH scene.  The rows of a batch are dealt to the ranks as contiguous ranges (weights); ranks
    other than 0 write 4-byte hit codes into rank 0's exchange buffer -- `transport` "ipc": straight from the trace kernel through a
    CUDA-IPC mapping of that buffer (NVLink peer stores), the arrival barrier being a 4-byte all-reduce on the trace stream;
    "nccl": into a local buffer that is sent to rank 0 (fallback where IPC mappings are not available) -- and rank 0 expands the
    codes into planes on a second stream while it traces its own, smaller share.  Nothing here synchronises the host.
    """

    def __init__(self, rto, scene, width, height, max_frames, flags, bias, device, transport="auto"):
        import torch
        import torch.distributed as dist
        self.rto, self.scene, self.W, self.H, self.flags, self.bias = rto, scene, width, height, flags, bias
        self.torch, self.dist = torch, dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.max_frames = max_frames
        self.words = rto.codes_frame_words(width, height)
        self.tiles_per_frame = (height + 7) // 8
        self.words_per_tile_row = self.words // self.tiles_per_frame
        self.device = device
        self.main = torch.cuda.ExternalStream(scene.stream, device=device)
        self.side = torch.cuda.Stream(device=device, priority=-1)       # the expansion gets SM slots as they free up, ahead of the next trace
        self.weights = [0.55 if (r == 0 and self.world > 1) else 1.0 for r in range(self.world)]
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        nbytes = self.words * 4 * max_frames
        self.local = [None, None]
        self.remote = [None, None]
        self.resolved = [torch.cuda.Event(), torch.cuda.Event()]
        self.trace_ev = None
        # exchange buffers live on rank 0; try the IPC mapping, agree on the transport
        handles = [None, None]
        if self.rank == 0:
            self.local = [rto.ExchangeBuffer(nbytes), rto.ExchangeBuffer(nbytes)]
            handles = [b.handle for b in self.local]
        box = [handles]
        dist.broadcast_object_list(box, src=0)
        ok = 1
        if self.rank != 0 and transport in ("auto", "ipc"):
            try:
                self.remote = [rto.ExchangeBuffer.open(h) for h in box[0]]
            except Exception:
                ok = 0
        if transport == "nccl":
            ok = 0
        agree = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        self.transport = "ipc" if int(agree.item()) == 1 else "nccl"
        if self.transport == "nccl":
            for b in self.remote:
                if b is not None:
                    b.close()
            self.remote = [None, None]
            if self.rank != 0:
                self.local = [rto.ExchangeBuffer(nbytes), rto.ExchangeBuffer(nbytes)]
            self.views = [torch.as_tensor(_CudaBuffer(b.ptr, nbytes), device=device) for b in self.local]
        self.batches = 0
        self.timing = False
        self.t_begin, self.t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def calibrate(self, render_once, rounds=6):
        """Adapt the shares to the measured trace time of every rank (rank 0's with the expansion running beside it): `render_once()`
        enqueues a representative batch.  Synchronises; call outside timed regions.  Returns the list of per-round times."""
        torch, dist = self.torch, self.dist
        hist = []
        self.timing = True
        for _ in range(rounds):
            render_once()
            self.finish()
            torch.cuda.synchronize()
            mine = torch.tensor([self.t_begin.elapsed_time(self.t_end)], dtype=torch.float64, device=self.device)
            allt = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(allt, mine)
            times = [float(x.item()) for x in allt]
            hist.append(times)
            if self.world > 1:
                self.weights = rebalance(self.weights, times)
        self.timing = False
        return hist

    def ranges(self, num_frames):
        cuts = deal_tiles(self.tiles_per_frame * num_frames, self.weights)
        return [(tile_row_to_row(cuts[r], self.H), tile_row_to_row(cuts[r + 1], self.H), cuts[r], cuts[r + 1]) for r in range(self.world)]

    def render(self, cams, rgba_ptr, id_ptr, t_ptr):
        """Enqueue one batch: `cams` (ctypes array of all frames of the batch, the same on every rank); on rank 0 the planes of all
        frames are complete when both streams have run (wait(), or order after self.main).  Returns this rank's (row0, row1)."""
        rto, torch, dist = self.rto, self.torch, self.dist
        n = len(cams)
        assert n <= self.max_frames
        par = self.batches & 1
        self.batches += 1
        rng = self.ranges(n)
        g0, g1, t0, t1 = rng[self.rank]
        W, H = self.W, self.H
        if self.timing:
            self.t_begin.record(self.main)
        if self.rank == 0:
            self.main.wait_event(self.resolved[par])                   # the plane set this batch goes into was last written two batches ago
            for (f, k, y0, y1) in split_rows(g0, g1, H):
                off = (f * H + y0) * W
                sub = (rto.RtoCamera * k).from_buffer(cams, f * C_sizeof_cam(rto))
                self.scene.render_device(sub, rto.MODE_BVH, self.flags, self.bias, y0, y1, rgba_ptr + off * 16, id_ptr + off * 4, t_ptr + off * 4)
        else:
            if self.transport == "ipc":
                # rank 0 may still be expanding the batch before last out of this buffer: its all-reduce below comes after that
                dst = self.remote[par].ptr
            else:
                dst = self.local[par].ptr
            for (f, k, y0, y1) in split_rows(g0, g1, H):
                sub = (rto.RtoCamera * k).from_buffer(cams, f * C_sizeof_cam(rto))
                self.scene.render_codes(sub, self.flags, self.bias, dst, first_frame=f, y0=y0, y1=y1)
        if self.timing:
            self.t_end.record(self.main)
        with torch.cuda.stream(self.main):
            if self.transport == "ipc":
                if self.rank == 0:
                    self.main.wait_event(self.resolved[par ^ 1])       # peers start the NEXT batch (other buffer) after this barrier
                dist.all_reduce(self.flag)                             # "every rank's codes of this batch are in rank 0's memory"
            else:
                if self.rank == 0:
                    self.main.wait_event(self.resolved[par])           # the buffer the receives land in is free again
                    reqs = []
                    for r in range(1, self.world):
                        a, b = rng[r][2] * self.words_per_tile_row, rng[r][3] * self.words_per_tile_row
                        if b > a:
                            reqs.append(dist.P2POp(dist.irecv, self.views[par][a:b], r))
                    if reqs:
                        for w in dist.batch_isend_irecv(reqs):
                            w.wait()
                else:
                    a, b = t0 * self.words_per_tile_row, t1 * self.words_per_tile_row
                    if b > a:
                        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, self.views[par][a:b], 0)]):
                            w.wait()
        if self.rank == 0:
            arrived = torch.cuda.Event()
            arrived.record(self.main)
            self.side.wait_event(arrived)
            src = self.local[par].ptr
            for r in range(1, self.world):
                for (f, k, y0, y1) in split_rows(rng[r][0], rng[r][1], H):
                    off = (f * H + y0) * W
                    sub = (rto.RtoCamera * k).from_buffer(cams, f * C_sizeof_cam(rto))
                    self.scene.resolve_codes(sub, src, first_frame=f, y0=y0, y1=y1, rgba_ptr=rgba_ptr + off * 16, id_ptr=id_ptr + off * 4,
                                             t_ptr=t_ptr + off * 4, stream=self.side.cuda_stream)
            self.resolved[par].record(self.side)
        return g0, g1

    def finish(self):
        """Order the main stream after everything rank 0 has expanded so far (call before timing events / reading planes)."""
        if self.rank == 0:
            self.main.wait_event(self.resolved[0])
            self.main.wait_event(self.resolved[1])

    def close(self):
        self.torch.cuda.synchronize()
        self.dist.barrier()                                            # nobody unmaps a buffer a peer may still be writing into
        self.torch.cuda.synchronize()
        for b in list(self.remote) + list(self.local):
            if b is not None:
                b.close()


def C_sizeof_cam(rto):
    import ctypes
    return ctypes.sizeof(rto.RtoCamera)
