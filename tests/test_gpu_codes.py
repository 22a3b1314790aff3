"""Compact frames, exchange memory and device groups (include/rto_c.h): a frame that travels as 4-byte hit codes and is rebuilt
where it is wanted equals a direct render bit for bit -- on one GPU, between two processes (CUDA IPC) and across the GPUs of a box."""
import ctypes as C
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cams(rto, n, W, H, radius=0.6 * 4250, theta=35.0):
    return [rto.Camera.from_degrees(theta, 40.0 + 360.0 * k / max(n, 1), radius).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0] for k in range(n)]


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_codes_then_resolve_equals_direct_render(gpu, dt_scene):
    """Full frames, odd sizes (partial tiles on both axes) and several cameras per call, with and without shadows."""
    rto = gpu
    sc = dt_scene["bvh"]
    bias = 1e-3 * dt_scene["grid"].voxel_size
    for (W, H, n) in ((640, 360, 1), (333, 203, 3), (16, 8, 2), (1, 1, 1)):
        cams = _cams(rto, n, W, H)
        for flags in (0, rto.FLAG_SHADOWS):
            words = rto.codes_frame_words(W, H)
            assert words == ((W + 15) // 16) * ((H + 7) // 8) * 128
            buf = rto.ExchangeBuffer(words * 4 * n)
            sc.render_codes(cams, flags, bias, buf.ptr)
            out = dict(rgba=np.empty((n, W * H, 4), np.float32), id=np.empty((n, W * H), np.int32), t=np.empty((n, W * H), np.float32))
            sc.resolve_codes(cams, buf.ptr, rgba_ptr=out["rgba"].ctypes.data, id_ptr=out["id"].ctypes.data, t_ptr=out["t"].ctypes.data, memory=rto.MEM_HOST)
            buf.close()
            for k, cam in enumerate(cams):
                want = sc.render(cam, rto.MODE_BVH, flags, bias)
                for key in ("id", "t", "rgba"):
                    assert_bit_equal(out[key][k], want[key], "%dx%d camera %d flags %d %s" % (W, H, k, flags, key))
            if W > 100:
                assert (out["id"] >= 0).mean() > 0.2


def test_codes_in_row_bands_and_frame_offsets(gpu, dt_scene):
    """The way a sharded batch is assembled: different calls fill different row bands and frame slots of one code buffer."""
    rto = gpu
    sc = dt_scene["bvh"]
    bias = 1e-3 * dt_scene["grid"].voxel_size
    W, H, n = 480, 270, 3
    cams = _cams(rto, n, W, H)
    words = rto.codes_frame_words(W, H)
    buf = rto.ExchangeBuffer(words * 4 * n)
    # frame 0: three bands; frame 1: whole; frame 2: two bands issued in reverse order
    for (f, y0, y1) in ((0, 0, 64), (0, 64, 200), (0, 200, H), (1, 0, H), (2, 136, H), (2, 0, 136)):
        sc.render_codes(cams[f], rto.FLAG_SHADOWS, bias, buf.ptr, first_frame=f, y0=y0, y1=y1)
    out = dict(rgba=np.empty((n, W * H, 4), np.float32), id=np.empty((n, W * H), np.int32), t=np.empty((n, W * H), np.float32))
    sc.resolve_codes(cams, buf.ptr, rgba_ptr=out["rgba"].ctypes.data, id_ptr=out["id"].ctypes.data, t_ptr=out["t"].ctypes.data, memory=rto.MEM_HOST)
    for k, cam in enumerate(cams):
        want = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
        for key in ("id", "t", "rgba"):
            assert_bit_equal(out[key][k], want[key], "frame %d %s" % (k, key))
    # a band resolved on its own lands at the start of the planes it is given
    band = dict(rgba=np.empty(((H - 136) * W, 4), np.float32), id=np.empty((H - 136) * W, np.int32), t=np.empty((H - 136) * W, np.float32))
    sc.resolve_codes(cams[2], buf.ptr, first_frame=2, y0=136, y1=H, rgba_ptr=band["rgba"].ctypes.data, id_ptr=band["id"].ctypes.data, t_ptr=band["t"].ctypes.data, memory=rto.MEM_HOST)
    assert_bit_equal(band["id"], out["id"][2][136 * W:], "band ids")
    assert_bit_equal(band["rgba"], out["rgba"][2][136 * W:], "band rgba")
    buf.close()
    with pytest.raises(rto.RtoError):
        sc.render_codes(cams[0], 0, bias, 4096, y0=4, y1=64)            # bands must start on a tile row
    with pytest.raises(rto.RtoError):
        dt_scene["oct"].render_codes(cams[0], 0, bias, 4096)             # hit codes exist for BVH scenes only


def _ipc_worker(handle, W, H, n, bias, q):
    """Second process: maps the first process's exchange buffer and traces frames 1.. into it."""
    try:
        sys.path.insert(0, ROOT)
        import ray_tracing_octrees_b200 as rto
        assert rto.lib().rto_init(0) == 0
        grid = rto.VoxelGrid.load(os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz"))
        tris = rto.marching_cubes_mesh(grid, rto.create_octree_from_voxel_grid(grid))
        sc = rto.Scene.bvh(tris)
        buf = rto.ExchangeBuffer.open(handle)
        cams = _cams(rto, n, W, H)
        sc.render_codes(cams[1:], rto.FLAG_SHADOWS, bias, buf.ptr, first_frame=1)
        sc.sync()
        buf.close()
        q.put("ok")
    except Exception as e:      # pragma: no cover
        q.put("worker failed: %r" % (e,))


def test_codes_from_another_process_through_cuda_ipc(gpu, dt_scene):
    """One process per GPU is how the bench (torchrun) and most multi-GPU hosts run: the gathering process exports its exchange
    buffer, the tracing process maps it and writes hit codes into it from its trace kernel, the gathering process expands them.
    (Both processes share cuda:0 here; across GPUs the same mapping goes over NVLink.)"""
    rto = gpu
    sc = dt_scene["bvh"]
    bias = 1e-3 * dt_scene["grid"].voxel_size
    W, H, n = 320, 200, 3
    cams = _cams(rto, n, W, H)
    words = rto.codes_frame_words(W, H)
    buf = rto.ExchangeBuffer(words * 4 * n)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_ipc_worker, args=(buf.handle, W, H, n, bias, q))
    p.start()
    sc.render_codes(cams[0], rto.FLAG_SHADOWS, bias, buf.ptr, first_frame=0)      # frame 0 is traced here, 1.. over there
    msg = q.get(timeout=300)
    p.join(timeout=60)
    assert msg == "ok", msg
    out = dict(rgba=np.empty((n, W * H, 4), np.float32), id=np.empty((n, W * H), np.int32), t=np.empty((n, W * H), np.float32))
    sc.resolve_codes(cams, buf.ptr, rgba_ptr=out["rgba"].ctypes.data, id_ptr=out["id"].ctypes.data, t_ptr=out["t"].ctypes.data, memory=rto.MEM_HOST)
    buf.close()
    for k, cam in enumerate(cams):
        want = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
        for key in ("id", "t", "rgba"):
            assert_bit_equal(out[key][k], want[key], "frame %d %s (codes written by another process)" % (k, key))


@pytest.mark.parametrize("ndev", [1, 2, 4, 8])
def test_device_group_equals_single_device(gpu, dt_scene, ndev):
    """rto_group_*: the planes gathered on the first device equal a single-device render bit for bit, whatever the shares."""
    rto = gpu
    if _device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    bias = 1e-3 * dt_scene["grid"].voxel_size
    W, H, n = 640, 360, 5
    cams = _cams(rto, n, W, H)
    grp = rto.Group(list(range(ndev)))
    grp.set_mesh(dt_scene["tris"])
    want = [dt_scene["bvh"].render(c, rto.MODE_BVH, rto.FLAG_SHADOWS, bias) for c in cams]
    for rnd, (weights, chunks) in enumerate(((None, 0), (None, 0), ([0.3] + [1.0] * (ndev - 1), 1), ([2.0] + [0.5] * (ndev - 1), 7))):
        if weights is not None:
            grp.set_balancing(False, weights, chunks)
        got = grp.render(cams, rto.FLAG_SHADOWS, bias)
        for k in range(n):
            for key in ("id", "t", "rgba"):
                assert_bit_equal(got[key][k], want[k][key], "round %d frame %d %s on %d devices" % (rnd, k, key, ndev))
    assert (grp.last_ms()[:ndev] > 0).all()
    # one camera, odd size, no shadows
    cam = _cams(rto, 1, 333, 203)[0]
    got = grp.render(cam, 0, bias)
    w1 = dt_scene["bvh"].render(cam, rto.MODE_BVH, 0, bias)
    for key in ("id", "t", "rgba"):
        assert_bit_equal(got[key][0], w1[key], "odd-size frame " + key)
    grp.close()


def test_group_refuses_bad_devices(gpu):
    rto = gpu
    with pytest.raises(rto.RtoError):
        rto.Group([0, 0])
    with pytest.raises(rto.RtoError):
        rto.Group([4096])
    g = rto.Group([0])
    with pytest.raises(rto.RtoError):
        g.render(_cams(rto, 1, 32, 32), 0, 0.0)                          # no scene yet
    g.close()
