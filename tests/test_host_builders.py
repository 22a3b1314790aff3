"""CPU only: the product's host builders (librto.so rto_host_*) against the oracle on the same inputs --
identical node arrays, triangle order, BVH shape and camera constants, including edge cases."""
import hashlib
import numpy as np
import pytest

from conftest import assert_bit_equal
from oracle import bind


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _random_grid(rto, dims, fill, seed):
    rng = np.random.default_rng(seed)
    data = (rng.random(dims[0] * dims[1] * dims[2]) < fill).astype(np.uint8)
    return rto.VoxelGrid(dims, (-3.0, 1.0, 0.5), 0.4, data)


@pytest.mark.parametrize("dims,fill", [((1, 1, 1), 1.0), ((1, 1, 1), 0.0), ((2, 2, 2), 0.5), ((7, 3, 5), 0.3), ((16, 16, 16), 0.05),
                                       ((33, 9, 17), 0.5), ((8, 8, 8), 1.0), ((8, 8, 8), 0.0), ((40, 40, 3), 0.9)])
def test_octree_mc_bvh_match_oracle(rto, checker, dims, fill):
    g = _random_grid(rto, dims, fill, seed=dims[0] * 131 + dims[1] * 17 + dims[2])
    nodes = rto.create_octree_from_voxel_grid(g)
    oc = checker.octree(g.dims, g.min, g.voxel_size, g.data)
    assert oc.build() == len(nodes)
    assert np.array_equal(nodes, oc.flat())
    tris = rto.marching_cubes_mesh(g, nodes)
    m = oc.mesh()
    assert_bit_equal(tris, m.tris(), "MC triangles")
    hb = rto.HostBVH(tris)
    m.build()
    boxes, meta = hb.export()
    rb, rm = m.export()
    if len(tris):
        assert_bit_equal(boxes, rb, "BVH boxes")
    assert np.array_equal(meta, rm)


def test_empty_grid_gives_no_octree(rto):
    g = rto.VoxelGrid((0, 4, 4), (0, 0, 0), 1.0, np.zeros(0, np.uint8))
    assert len(rto.create_octree_from_voxel_grid(g)) == 0     # createOctreeFromVoxelGrid returns nullptr (OctreeVoxel.cpp:766)


def test_sphere128_and_dt_checksums(rto, golden_meta, dt_grid_path):
    g = rto.generate_test_volume(128)
    nodes = rto.create_octree_from_voxel_grid(g)
    d = golden_meta["sphere128"]
    assert len(nodes) == d["nodes"] and sha(nodes) == d["flat_sha"]
    tris = rto.marching_cubes_mesh(g, nodes)
    assert len(tris) == d["tris"] and sha(tris) == d["tris_sha"]
    boxes, meta = rto.HostBVH(tris).export()
    assert sha(boxes) == d["bvh_boxes_sha"] and sha(meta) == d["bvh_meta_sha"]

    g = rto.VoxelGrid.load(dt_grid_path)
    d = golden_meta["dt"]
    assert list(g.dims) == d["dims"] and g.voxel_size == d["voxel"] and [float(x) for x in g.min] == d["gmin"]
    nodes = rto.create_octree_from_voxel_grid(g)
    assert len(nodes) == d["nodes"] and sha(nodes) == d["flat_sha"]
    tris = rto.marching_cubes_mesh(g, nodes)
    assert len(tris) == d["tris"] and sha(tris) == d["tris_sha"]
    boxes, meta = rto.HostBVH(tris).export()
    assert sha(boxes) == d["bvh_boxes_sha"] and sha(meta) == d["bvh_meta_sha"]


def test_bvh_with_duplicate_centroids(rto, checker):
    """std::sort on equal keys is only reproducible if the call sequence is identical: many exact ties."""
    rng = np.random.default_rng(11)
    base = rng.integers(0, 4, (300, 3)).astype(np.float32)               # lots of coincident centroids
    tris = np.concatenate([base, base + [1, 0, 0], base + [0, 1, 0]], axis=1).astype(np.float32)
    tris = np.concatenate([tris, tris[:100]], axis=0)                      # exact duplicates
    hb = rto.HostBVH(tris)
    m = checker.mesh(tris)
    m.build()
    boxes, meta = hb.export()
    rb, rm = m.export()
    assert_bit_equal(boxes, rb, "BVH boxes")
    assert np.array_equal(meta, rm)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5])
def test_tiny_bvh(rto, checker, n):
    rng = np.random.default_rng(n)
    tris = rng.normal(0, 1, (n, 9)).astype(np.float32)
    hb = rto.HostBVH(tris)
    m = checker.mesh(tris)
    m.build()
    boxes, meta = hb.export()
    rb, rm = m.export()
    assert_bit_equal(boxes, rb, "BVH boxes")
    assert np.array_equal(meta, rm)


@pytest.mark.parametrize("args", [(30, 40, 1.2, (0, 0, 0), 45.0, 1024, 768), (-75, 300, 2550.0, (10, -5, 3), 60.0, 1920, 1080),
                                  (0, 0, 0.1, (0, 0, 0), 45.0, 7, 5), (89, 180, 1000.0, (0, 100, 0), 30.0, 3840, 2160)])
def test_camera_constants(rto, checker, args):
    th, ph, r, tgt, fov, w, h = args
    cam, view = rto.Camera.from_degrees(th, ph, r, tgt).consts(fov, float(np.float32(w) / np.float32(h)), w, h)
    rcam, rview = checker.camera(th, ph, r, target=tgt, fov_deg=fov, width=w, height=h)
    assert bytes(cam) == bytes(rcam)
    assert_bit_equal(view, rview, "view matrix")


def test_grid_io_roundtrip(rto, tmp_path, dt_grid_path):
    g = rto.VoxelGrid.load(dt_grid_path)
    p = str(tmp_path / "cache.bin")
    g.save(p)
    g2 = rto.VoxelGrid.load(p)
    assert g2.dims == g.dims and np.array_equal(g2.min, g.min) and g2.voxel_size == g.voxel_size
    assert np.array_equal(g2.data, g.data)
    import gzip
    assert open(p, "rb").read() == gzip.open(dt_grid_path, "rb").read()    # byte-identical to the reference's file format
    with pytest.raises(rto.RtoError):
        rto.VoxelGrid.load(str(tmp_path / "missing.bin"))
    (tmp_path / "short.bin").write_bytes(open(p, "rb").read()[:1000])
    with pytest.raises(rto.RtoError):
        rto.VoxelGrid.load(str(tmp_path / "short.bin"))
