"""CPU only: pin the oracle.  (1) the stand-alone port reproduces every golden vector generated from the compiled
reference; (2) where the compiled reference is present (this container, or its prebuilt .so on the GPU box) the
port is compared with it directly on more inputs."""
import hashlib
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_bit_equal, cam_from_dict
from oracle import bind


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_dt_fixture_is_the_reference_file(golden_meta, dt_grid_path):
    raw = gzip.open(dt_grid_path, "rb").read()
    assert hashlib.sha256(raw).hexdigest() == golden_meta["dt_sceneCache_sha256"]
    assert len(raw) == 2995011


def test_port_sphere32_structures(port, golden_sphere32):
    g = golden_sphere32
    data = np.unpackbits(g["voxels"])[:32 ** 3]
    dims, gmin, vox, data2 = bind.sphere_grid(32)
    assert np.array_equal(data, data2)                     # closed-form generateTestVolume
    oc = port.octree(dims, gmin, vox, data)
    assert oc.build() == len(g["flat"])
    assert np.array_equal(oc.flat(), g["flat"])
    m = oc.mesh()
    assert_bit_equal(m.tris(), g["tris"], "MC triangles")
    m.build()
    boxes, meta = m.export()
    assert_bit_equal(boxes, g["bvh_boxes"], "BVH boxes")
    assert np.array_equal(meta, g["bvh_meta"])
    off, ids = m.query(g["query_o"], g["query_d"])
    assert np.array_equal(off, g["query_off"]) and np.array_equal(ids, g["query_ids"])
    t, ids = oc.rayskip(g["edge_o"], g["edge_d"])
    assert_bit_equal(t, g["edge_t"], "octreeRaySkip on axis-parallel rays")
    assert np.array_equal(ids, g["edge_id"])


@pytest.mark.parametrize("name", ["a", "b", "inside"])
def test_port_sphere32_renders(port, golden_sphere32, golden_meta, name):
    g = golden_sphere32
    dims, gmin, vox, data = bind.sphere_grid(32)
    oc = port.octree(dims, gmin, vox, data)
    oc.build()
    m = oc.mesh()
    m.build()
    cam = cam_from_dict(bind.CamConsts, golden_meta["sphere32_cams"][name])
    for mode in (0, 1):
        o = oc.render(cam, mode, stats=True)
        for k in ("rgba", "id", "t"):
            assert_bit_equal(o[k], g["oct%d_%s_%s" % (mode, name, k)], "octree mode %d %s" % (mode, k))
        assert o["stats"][0] == g["oct%d_%s_stats" % (mode, name)][0]
    for flags in (0, 1):
        o = m.render(cam, flags, 1e-3 * vox, stats=True)
        for k in ("rgba", "id", "t"):
            assert_bit_equal(o[k], g["bvh%d_%s_%s" % (flags, name, k)], "bvh flags %d %s" % (flags, k))
        assert np.array_equal(o["stats"], g["bvh%d_%s_stats" % (flags, name)])


def test_port_camera_matches_golden(port, golden_meta):
    for name, (th, ph, r, w, h) in {"a": (30, 40, 1.2, 96, 64), "b": (-20, 200, 0.9, 64, 96), "inside": (10, 75, 0.25, 48, 48)}.items():
        cam, _ = port.camera(th, ph, r, width=w, height=h)
        want = cam_from_dict(bind.CamConsts, golden_meta["sphere32_cams"][name])
        assert bytes(cam) == bytes(want)


def test_port_dt_structures_and_rows(port, golden_meta, golden_dt, dt_grid_path, tmp_path):
    raw = gzip.open(dt_grid_path, "rb").read()
    p = tmp_path / "sceneCache.bin"
    p.write_bytes(raw)
    oc = port.octree(path=str(p))
    d = golden_meta["dt"]
    assert list(oc.dims) == d["dims"] and oc.voxel == d["voxel"]
    assert oc.build() == d["nodes"]
    assert sha(oc.flat()) == d["flat_sha"]
    m = oc.mesh()
    assert m.n == d["tris"] and sha(m.tris()) == d["tris_sha"]
    m.build()
    boxes, meta = m.export()
    assert sha(boxes) == d["bvh_boxes_sha"] and sha(meta) == d["bvh_meta_sha"]
    for name in ("far", "near"):
        cam = cam_from_dict(bind.CamConsts, golden_meta["dt_cams"][name])
        for bi, (y0, y1) in enumerate(golden_dt["bands"]):
            for mode in (0, 1):
                o = oc.render(cam, mode, int(y0), int(y1), stats=True)
                for k in ("id", "t", "rgba"):
                    assert_bit_equal(o[k], golden_dt["oct%d_%s_%d_%s" % (mode, name, bi, k)], "dt octree %d %s %s" % (mode, name, k))
                assert o["stats"][0] == golden_dt["oct%d_%s_%d_stats" % (mode, name, bi)][0]
            o = m.render(cam, 1, 1e-3 * oc.voxel, int(y0), int(y1), stats=True)
            for k in ("id", "t", "rgba"):
                assert_bit_equal(o[k], golden_dt["bvh1_%s_%d_%s" % (name, bi, k)], "dt bvh %s %s" % (name, k))
            assert np.array_equal(o["stats"], golden_dt["bvh1_%s_%d_stats" % (name, bi)])


def test_port_sphere128_checksums(port, golden_meta):
    dims, gmin, vox, data = bind.sphere_grid(128)
    oc = port.octree(dims, gmin, vox, data)
    d = golden_meta["sphere128"]
    assert oc.build() == d["nodes"] and sha(oc.flat()) == d["flat_sha"]
    m = oc.mesh()
    assert m.n == d["tris"] and sha(m.tris()) == d["tris_sha"]
    m.build()
    boxes, meta = m.export()
    assert sha(boxes) == d["bvh_boxes_sha"] and sha(meta) == d["bvh_meta_sha"]


# ---- direct port-vs-compiled-reference comparisons (skipped where libref.so is absent) ---------------------
@pytest.mark.parametrize("dims", [(20, 13, 7), (1, 1, 1), (5, 1, 3), (16, 16, 16)])
def test_port_vs_ref_random_grids(port, ref, dims):
    rng = np.random.default_rng(sum(dims))
    data = (rng.random(dims[0] * dims[1] * dims[2]) < 0.35).astype(np.uint8)
    gmin, vox = (-1.5, 0.25, 3.0), 0.37
    a, b = ref.octree(dims, gmin, vox, data), port.octree(dims, gmin, vox, data)
    assert a.build() == b.build()
    assert np.array_equal(a.flat(), b.flat())
    ma, mb = a.mesh(), b.mesh()
    assert_bit_equal(ma.tris(), mb.tris(), "MC triangles")
    if ma.n:
        ma.build(); mb.build()
        ba, mea = ma.export(); bb, meb = mb.export()
        assert_bit_equal(ba, bb, "BVH boxes"); assert np.array_equal(mea, meb)
    ext = max(dims) * vox
    cam_r, _ = ref.camera(25, 130, 2.2 * ext, target=(gmin[0] + dims[0] * vox / 2, gmin[1] + dims[1] * vox / 2, gmin[2] + dims[2] * vox / 2), width=80, height=60)
    cam_p, _ = port.camera(25, 130, 2.2 * ext, target=(gmin[0] + dims[0] * vox / 2, gmin[1] + dims[1] * vox / 2, gmin[2] + dims[2] * vox / 2), width=80, height=60)
    assert bytes(cam_r) == bytes(cam_p)
    for mode in (0, 1):
        oa, ob = a.render(cam_r, mode, stats=True), b.render(cam_p, mode, stats=True)
        assert oa["stats"][1] == 0                       # replay inside the harness agrees with the real octreeRaySkip
        for k in ("rgba", "id", "t"):
            assert_bit_equal(oa[k], ob[k], "mode %d %s" % (mode, k))
    if ma.n:
        oa, ob = ma.render(cam_r, 1, 1e-3 * vox, stats=True), mb.render(cam_p, 1, 1e-3 * vox, stats=True)
        for k in ("rgba", "id", "t"):
            assert_bit_equal(oa[k], ob[k], "bvh %s" % k)
        assert np.array_equal(oa["stats"], ob["stats"])


def test_port_vs_ref_axis_rays(port, ref):
    dims, gmin, vox, data = bind.sphere_grid(16)
    a, b = ref.octree(dims, gmin, vox, data), port.octree(dims, gmin, vox, data)
    a.build(); b.build()
    rng = np.random.default_rng(3)
    o = rng.uniform(-1, 1, (256, 3)).astype(np.float32)
    d = rng.normal(0, 1, (256, 3)).astype(np.float32)
    d[::4, 0] = 0.0; d[1::4, 1] = 0.0; d[2::8, 2] = -0.0; d[3::16] = [0, 0, 1]
    ta, ia = a.rayskip(o, d, 0.0, 1e30)
    tb, ib = b.rayskip(o, d, 0.0, 1e30)
    assert_bit_equal(ta, tb, "octreeRaySkip t"); assert np.array_equal(ia, ib); assert (ia != -2).all()
