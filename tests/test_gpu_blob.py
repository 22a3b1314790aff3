"""Scene files (rto_scene_save / rto_scene_load, include/rto_c.h): a scene read back from its file renders the same bits as the one
that was saved -- BVH scenes of both construction routes (incl. the exact replay and the hit codes), octree scenes in the compact and
in the general layout -- and files that are truncated, damaged or not scene files are refused."""
import os

import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def _cam(rto, W, H, phi=40.0, radius=0.6 * 4250, theta=35.0):
    return rto.Camera.from_degrees(theta, phi, radius).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0]


def _same_frames(a, b, what):
    for key in ("id", "t", "rgba"):
        assert_bit_equal(a[key], b[key], what + " " + key)


def test_bvh_scene_round_trip(gpu, dt_scene, tmp_path):
    rto = gpu
    sc = dt_scene["bvh"]
    bias = 1e-3 * dt_scene["grid"].voxel_size
    path = str(tmp_path / "dt_bvh.rtoscene")
    sc.save(path)
    back = rto.Scene.load(path)
    assert back.info()["prims"] == sc.info()["prims"] and back.info()["nodes"] == sc.info()["nodes"] and back.info()["kind"] == rto.MODE_BVH
    cam = _cam(rto, 640, 360)
    for flags in (0, rto.FLAG_SHADOWS, rto.FLAG_SHADOWS | rto.FLAG_NO_PRUNE):
        _same_frames(back.render(cam, rto.MODE_BVH, flags, bias), sc.render(cam, rto.MODE_BVH, flags, bias), "flags %d" % flags)
    # the reference-shaped tree travels too: BVH::query replay and the work counters
    o = np.array([[0.0, 3000.0, 0.0], [1500.0, 900.0, -1200.0]], np.float32)
    d = np.array([[0.01, -1.0, 0.02], [-0.6, -0.5, 0.6]], np.float32)
    off_a, ids_a = sc.query(o, d)
    off_b, ids_b = back.query(o, d)
    assert np.array_equal(off_a, off_b) and np.array_equal(ids_a, ids_b) and len(ids_a) > 0
    # hit codes out of the loaded scene, expanded by the original one
    words = rto.codes_frame_words(640, 360)
    buf = rto.ExchangeBuffer(words * 4)
    back.render_codes([cam], rto.FLAG_SHADOWS, bias, buf.ptr)
    back.sync()
    out = dict(rgba=np.empty((640 * 360, 4), np.float32), id=np.empty(640 * 360, np.int32), t=np.empty(640 * 360, np.float32))
    sc.resolve_codes([cam], buf.ptr, rgba_ptr=out["rgba"].ctypes.data, id_ptr=out["id"].ctypes.data, t_ptr=out["t"].ctypes.data, memory=rto.MEM_HOST)
    buf.close()
    _same_frames(out, sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias), "codes")
    back.close()


def test_device_built_bvh_round_trip(gpu, tmp_path):
    rto = gpu
    g = rto.generate_test_volume(64)
    sc = rto.Scene.bvh_from_grid(g)
    path = str(tmp_path / "sphere_dev.rtoscene")
    sc.save(path)
    back = rto.Scene.load(path)
    cam = rto.Camera.from_degrees(30, 40, 1.2).consts(45.0, float(np.float32(320) / np.float32(240)), 320, 240)[0]
    bias = 1e-3 * g.voxel_size
    a, b = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias), back.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
    _same_frames(b, a, "device-built tree")
    assert (a["id"] >= 0).mean() > 0.2
    with pytest.raises(rto.RtoError):            # still a device-built tree: no reference-shaped replay
        back.query(np.zeros((1, 3), np.float32), np.ones((1, 3), np.float32))


def test_octree_scene_round_trip(gpu, dt_scene, tmp_path):
    rto = gpu
    grid, nodes = dt_scene["grid"], dt_scene["nodes"]
    cam = _cam(rto, 480, 270)
    # compact layout (the reference builder's shape) and the general one (a frustum-culled array)
    aspect = float(np.float32(480) / np.float32(270))
    view = rto.Camera.from_degrees(35.0, 40.0, 0.6 * 4250).consts(45.0, aspect, 480, 270)[1]
    culled = rto.frustum_cull(nodes, grid, rto.view_proj(view, 45.0, aspect), 150.0)[0]
    for name, sc in (("compact", dt_scene["oct"]), ("general", rto.Scene.octree(culled, grid.min, grid.voxel_size))):
        path = str(tmp_path / ("oct_%s.rtoscene" % name))
        sc.save(path)
        back = rto.Scene.load(path)
        assert back.info()["compact"] == sc.info()["compact"] == (1 if name == "compact" else 0)
        modes = (rto.MODE_OCTREE_SKIP, rto.MODE_OCTREE_GLSL) if name == "compact" else (rto.MODE_OCTREE_GLSL,)
        for mode in modes:
            _same_frames(back.render(cam, mode), sc.render(cam, mode), "%s mode %d" % (name, mode))
        back.close()


def test_bad_files_are_refused(gpu, dt_scene, tmp_path):
    rto = gpu
    path = str(tmp_path / "oct.rtoscene")
    dt_scene["oct"].save(path)
    blob = open(path, "rb").read()
    cases = {"truncated": blob[: len(blob) // 2], "no checksum": blob[:-8], "not a scene": b"RTOSCN0\0" + blob[8:],
             "other version": blob[:8] + (99).to_bytes(4, "little") + blob[12:], "empty": b""}
    flipped = bytearray(blob)
    flipped[len(blob) // 2] ^= 0x10
    cases["bit flip"] = bytes(flipped)
    for what, data in cases.items():
        bad = str(tmp_path / "bad.rtoscene")
        open(bad, "wb").write(data)
        with pytest.raises(rto.RtoError) as e:
            rto.Scene.load(bad)
        assert e.value.code in (5, 6), what
        if what == "other version":
            assert e.value.code == 6
    with pytest.raises(rto.RtoError) as e:
        rto.Scene.load(str(tmp_path / "missing.rtoscene"))
    assert e.value.code == 5
    with pytest.raises(rto.RtoError) as e:
        dt_scene["oct"].save(str(tmp_path / "no_such_dir" / "x.rtoscene"))
    assert e.value.code == 5
