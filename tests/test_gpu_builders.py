"""GPU scene construction (rto_build.cu) against the host builders, which tests/test_oracle_pinning.py and
tests/test_host_builders.py pin to the reference: same octree array, same triangle soup, same device layout, bit for bit."""
import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def _grids(rto, dt_grid_path):
    rng = np.random.default_rng(7)
    out = {}
    out["sphere32"] = rto.generate_test_volume(32)
    out["sphere48_offgrid"] = rto.VoxelGrid((48, 48, 48), (-3.25, 0.5, 7.0), 0.37, rto.generate_test_volume(48).data)
    d = (20, 33, 7)
    out["ragged_noise"] = rto.VoxelGrid(d, (-1.0, -2.0, -3.0), 0.5, (rng.random(d[0] * d[1] * d[2]) < 0.3).astype(np.uint8))
    d = (17, 17, 17)
    out["odd_dense"] = rto.VoxelGrid(d, (0, 0, 0), 1.0, (rng.random(d[0] * d[1] * d[2]) < 0.9).astype(np.uint8))
    out["all_empty"] = rto.VoxelGrid((8, 8, 8), (0, 0, 0), 1.0, np.zeros(512, np.uint8))
    out["all_filled"] = rto.VoxelGrid((8, 8, 8), (0, 0, 0), 1.0, np.ones(512, np.uint8))
    out["all_filled_ragged"] = rto.VoxelGrid((5, 8, 3), (0, 0, 0), 1.0, np.ones(120, np.uint8))
    out["single_voxel"] = rto.VoxelGrid((1, 1, 1), (0, 0, 0), 1.0, np.ones(1, np.uint8))
    out["thin_slab"] = rto.VoxelGrid((16, 1, 16), (0, 0, 0), 2.0, (rng.random(256) < 0.5).astype(np.uint8))
    out["city128"] = rto.city_block_grid(128, 99, 8)
    out["dt"] = rto.VoxelGrid.load(dt_grid_path)
    return out


@pytest.fixture(scope="module")
def grids(rto, dt_grid_path):
    assert rto.lib().rto_init(0) == 0, rto.lib().rto_last_error().decode()
    return _grids(rto, dt_grid_path)


NAMES = ["sphere32", "sphere48_offgrid", "ragged_noise", "odd_dense", "all_empty", "all_filled", "all_filled_ragged", "single_voxel",
         "thin_slab", "city128", "dt"]


@pytest.mark.parametrize("name", NAMES)
def test_device_octree_equals_host_octree(rto, grids, name):
    g = grids[name]
    host = rto.create_octree_from_voxel_grid(g)
    dev = rto.create_octree_on_device(g)
    assert dev.shape == host.shape, "%s: %s vs %s nodes" % (name, dev.shape, host.shape)
    assert np.array_equal(dev, host), "%s: GPUNodes arrays differ" % name


@pytest.mark.parametrize("name", NAMES)
def test_device_mesh_equals_host_mesh(rto, grids, name):
    g = grids[name]
    host = rto.marching_cubes_mesh(g, rto.create_octree_from_voxel_grid(g))
    dev = rto.marching_cubes_mesh_on_device(g)
    assert dev.shape == host.shape, "%s: %d vs %d triangles" % (name, len(dev), len(host))
    assert_bit_equal(dev, host, name + " triangles")


@pytest.mark.parametrize("name", ["sphere32", "ragged_noise", "odd_dense", "all_filled", "single_voxel", "city128", "dt"])
def test_scene_from_grid_has_the_same_device_layout(rto, grids, name):
    g = grids[name]
    a = rto.Scene.octree(rto.create_octree_from_voxel_grid(g), g.min, g.voxel_size)
    b = rto.Scene.octree_from_grid(g)
    ia, ib = a.info(), b.info()
    assert ia["nodes"] == ib["nodes"] and ia["prims"] == ib["prims"] and ib["compact"] == 1 and ia["compact"] == 1
    for x, y, what in zip(a.octree_layout(), b.octree_layout(), ("desc", "up", "inner")):
        assert np.array_equal(x, y), "%s: %s differs" % (name, what)


def test_scene_from_grid_renders_like_the_oracle(rto, checker, grids):
    g = grids["sphere32"]
    sc = rto.Scene.octree_from_grid(g)
    cam, _ = rto.Camera.from_degrees(30, 40, 1.2).consts(45.0, float(np.float32(96) / np.float32(64)), 96, 64)
    rcam, _ = checker.camera(30, 40, 1.2, width=96, height=64)
    oc = checker.octree(g.dims, g.min, g.voxel_size, g.data)
    oc.build()
    for mode, key in ((rto.MODE_OCTREE_SKIP, 0), (rto.MODE_OCTREE_GLSL, 1)):
        got, want = sc.render(cam, mode), oc.render(rcam, key)
        assert np.array_equal(got["id"], want["id"])
        assert_bit_equal(got["t"], want["t"], "t")
        assert_bit_equal(got["rgba"], want["rgba"].reshape(-1, 4), "rgba")


def test_reserved_voxel_value_is_rejected(rto, grids):
    g = rto.VoxelGrid((4, 4, 4), (0, 0, 0), 1.0, np.full(64, 255, np.uint8))
    with pytest.raises(rto.RtoError):
        rto.create_octree_on_device(g)
    with pytest.raises(rto.RtoError):
        rto.create_octree_from_voxel_grid(g)


# ---- linear BVH built on the device: the fast route, not reference-shaped ----------------------------------------------
def _near_tie_report(dev, host):
    ids_differ = dev["id"] != host["id"]
    same = ~ids_differ
    # the same triangle gives the same t and colour to the last bit whatever tree found it
    assert np.array_equal(dev["t"][same].view(np.uint32), host["t"][same].view(np.uint32))
    assert np.array_equal(dev["rgba"][same].view(np.uint32), host["rgba"][same].view(np.uint32))
    return int(ids_differ.sum())


@pytest.mark.parametrize("name,theta,phi,radius,size", [("sphere32", 30, 40, 1.2, (256, 192)), ("ragged_noise", 20, 70, 30.0, (128, 128)),
                                                       ("city128", 35, 40, 115.0, (320, 240)), ("dt", 35, 40, 0.6 * 4250, (960, 540))])
def test_device_bvh_hits_equal_reference_tree_hits_except_near_ties(rto, grids, name, theta, phi, radius, size):
    g = grids[name]
    tris = rto.marching_cubes_mesh(g, rto.create_octree_from_voxel_grid(g))
    host = rto.Scene.bvh(tris)
    dev = rto.Scene.bvh_device(tris)
    grid_dev = rto.Scene.bvh_from_grid(g)
    assert dev.info()["prims"] == len(tris) == grid_dev.info()["prims"]
    w, h = size
    cam, _ = rto.Camera.from_degrees(theta, phi, radius).consts(45.0, float(np.float32(w) / np.float32(h)), w, h)
    bias = 1e-3 * g.voxel_size
    for flags in (0, rto.FLAG_SHADOWS, rto.FLAG_NO_PRUNE):
        a = host.render(cam, rto.MODE_BVH, flags, bias)
        b = dev.render(cam, rto.MODE_BVH, flags, bias)
        c = grid_dev.render(cam, rto.MODE_BVH, flags, bias)
        # triangles -> device BVH and grid -> device BVH are the same build on the same triangle soup
        assert np.array_equal(b["id"], c["id"]) and np.array_equal(b["t"].view(np.uint32), c["t"].view(np.uint32))
        if flags & rto.FLAG_SHADOWS:
            # a shadow ray starts at the hit point: only pixels whose primary hit agrees are comparable bit for bit
            same = a["id"] == b["id"]
            frac = 1.0 - same.mean()
            lit_differs = (a["rgba"][same] != b["rgba"][same]).any(axis=1).mean()
            assert frac <= 1e-4 and lit_differs <= 1e-4, (name, frac, lit_differs)
            continue
        # ids differ only where two triangles tie in t (shared edges) or a ray grazes the edge of a leaf box, whose extent depends
        # on which two triangles share the leaf (measured at 4 x 1080p: 8 / 27 / 79 / 130 of 8.3 M pixels on C1 / DT / C3 / C4)
        n_diff = _near_tie_report(b, a)
        assert n_diff <= 1e-4 * w * h, "%s: %d of %d hit ids differ from the reference-shaped tree" % (name, n_diff, w * h)


@pytest.mark.parametrize("name,theta,phi,radius", [("sphere32", 30, 40, 1.2), ("ragged_noise", 20, 70, 30.0), ("city128", 35, 40, 115.0), ("dt", 35, 40, 0.6 * 4250)])
def test_block_rebuild_of_the_device_tree_equals_the_serial_one(rto, grids, name, theta, phi, radius, tmp_path, monkeypatch):
    """The surface-area rebuild of the device-built tree runs as one block per subtree (rto_sahchunk.h sah_rebuild_chunk_block); the
    one-thread form, which tests/test_emu_traversal.py checks on the CPU, is kept behind RTO_SAH_SERIAL=1.  Both choose their splits from the
    same bins, so they build the same tree: the saved device layouts are compared byte for byte (they may differ only below ranges whose
    centroids all coincide, where any halving is valid -- then the frames still have to agree up to near-ties), and the frames bit for bit."""
    g = grids[name]
    tris = rto.marching_cubes_mesh(g, rto.create_octree_from_voxel_grid(g))
    block = rto.Scene.bvh_device(tris)
    monkeypatch.setenv("RTO_SAH_SERIAL", "1")
    serial = rto.Scene.bvh_device(tris)
    monkeypatch.delenv("RTO_SAH_SERIAL")
    pa, pb = str(tmp_path / "block.rtoscene"), str(tmp_path / "serial.rtoscene")
    block.save(pa); serial.save(pb)
    same_layout = open(pa, "rb").read() == open(pb, "rb").read()
    w, h = 480, 270
    cam, _ = rto.Camera.from_degrees(theta, phi, radius).consts(45.0, float(np.float32(w) / np.float32(h)), w, h)
    bias = 1e-3 * g.voxel_size
    for flags in (rto.FLAG_NO_PRUNE, rto.FLAG_SHADOWS):
        a, b = block.render(cam, rto.MODE_BVH, flags, bias), serial.render(cam, rto.MODE_BVH, flags, bias)
        if same_layout or flags == rto.FLAG_NO_PRUNE:       # the exact path does not depend on the tree at all
            assert np.array_equal(a["id"], b["id"]) and np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
            assert np.array_equal(a["rgba"].view(np.uint32), b["rgba"].view(np.uint32))
        else:
            assert (a["id"] != b["id"]).mean() <= 1e-4
    assert (a["id"] >= 0).any()
    print("%s: block and serial rebuild give %s device layouts" % (name, "byte-identical" if same_layout else "different (ranges whose box centres coincide, e.g. the two triangles of a Marching-Cubes quad)"))


def test_device_bvh_edge_cases(rto, grids):
    cam, _ = rto.Camera.from_degrees(30, 40, 6.0).consts(45.0, 1.0, 64, 64)
    # no triangles, one triangle, two, three (one and a half leaves)
    base = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0], [0, 0, 1, 1, 0, 1, 0, 1, 1], [0, 0, -1, 1, 0, -1, 0, 1, -1]], np.float32)
    for n in (0, 1, 2, 3):
        tris = base[:n]
        dev = rto.Scene.bvh_device(tris)
        b = dev.render(cam, rto.MODE_BVH, 0, 0.0)
        if n == 0:
            assert (b["id"] == -1).all()
            continue
        a = rto.Scene.bvh(tris).render(cam, rto.MODE_BVH, 0, 0.0)
        assert np.array_equal(a["id"], b["id"]) and np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    # all-empty grid: no surface, an empty scene that renders misses
    e = rto.Scene.bvh_from_grid(grids["all_empty"])
    assert e.info()["prims"] == 0
    assert (e.render(cam, rto.MODE_BVH, 0, 0.0)["id"] == -1).all()
    # the reference-tree-only entry points say so instead of answering from the wrong tree
    dev = rto.Scene.bvh_device(base)
    with pytest.raises(rto.RtoError) as err:
        dev.stats(cam, rto.MODE_BVH, 0, 0.0)
    assert err.value.code == 6


# ---- Dual Contouring mesh on the GPU (rto_dc.cu) -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", NAMES)
def test_device_dc_mesh_equals_host(rto, grids, name):
    """Same triangles in the same order as rto_host_dc_mesh (which tests/test_dc_mesh.py pins to the reference's own mesher)."""
    g = grids[name]
    nodes = rto.create_octree_from_voxel_grid(g)
    assert_bit_equal(rto.dual_contouring_mesh(g, nodes, algo="device"), rto.dual_contouring_mesh(g, nodes), "DC triangles " + name)


def test_device_dc_mesh_equals_the_golden_checksums(rto, grids):
    """Every case of tests/golden/golden_dc.json (generated from the reference's AdaptiveDualContouringRenderer.cpp), culled walks included."""
    import hashlib, json, os
    from dc_cases import CASES, GOLDEN, make_grid
    meta = json.load(open(os.path.join(GOLDEN, "golden_dc.json")))
    for name in sorted(CASES):
        case, want = CASES[name], meta[name]
        dims, gmin, voxel, data = make_grid(case)
        g = rto.VoxelGrid(dims, gmin, voxel, data)
        nodes = rto.create_octree_from_voxel_grid(g)
        vp = None if want["view_proj"] is None else np.array(want["view_proj"], np.float32)
        got = rto.dual_contouring_mesh(g, nodes, vp, case.get("margin", 50.0), algo="device")
        assert len(got) == want["tris"], name
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == want["sha"], name


def test_device_dc_mesh_fallback_rounds(rto, grids):
    """Noise grids: most boundary leaves fall back and their cached centres change what later leaves see (several rounds)."""
    rng = np.random.default_rng(5)
    for p in (0.05, 0.3, 0.6, 0.95):
        d = tuple(int(x) for x in rng.integers(5, 40, 3))
        g = rto.VoxelGrid(d, (0.5, -1.0, 2.0), 0.3, (rng.random(d[0] * d[1] * d[2]) < p).astype(np.uint8))
        nodes = rto.create_octree_from_voxel_grid(g)
        assert_bit_equal(rto.dual_contouring_mesh(g, nodes, algo="device"), rto.dual_contouring_mesh(g, nodes, algo="replay"), "noise p=%g" % p)


@pytest.mark.parametrize("name", ["sphere32", "city128", "dt"])
def test_scene_from_grid_dc_equals_scene_from_the_dc_soup(rto, grids, name):
    """grid -> octree -> DC mesh -> linear BVH without leaving the device renders exactly like the linear BVH over rto_host_dc_mesh's soup."""
    g = grids[name]
    nodes = rto.create_octree_from_voxel_grid(g)
    tris = rto.dual_contouring_mesh(g, nodes)
    a, b = rto.Scene.bvh_from_grid_dc(g), rto.Scene.bvh_device(tris)
    assert a.info()["prims"] == b.info()["prims"] == len(tris)
    ext = float(max(g.dims)) * g.voxel_size
    centre = tuple(float(g.min[i]) + 0.5 * g.dims[i] * g.voxel_size for i in range(3))
    cam, _ = rto.Camera.from_degrees(35, 40, 1.1 * ext, centre).consts(45.0, 1.5, 384, 256)
    fa, fb = a.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3 * g.voxel_size), b.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3 * g.voxel_size)
    assert (fa["id"] >= 0).any()
    for k in ("id", "t", "rgba"):
        assert_bit_equal(fa[k], fb[k], "%s %s" % (name, k))


def test_device_built_tree_exact_path_and_no_prune(rto, grids):
    """Rays that are not admitted to the fused node tests (zero direction components, far origins) and RTO_FLAG_NO_PRUNE walk the
    device-built tree with the reference's select-form box test; that tree is stored in the paired node layout, which this path has
    to unpack.  Same hit distances as the host route's scene (ids may differ at exact ties only)."""
    g = grids["sphere32"]
    nodes = rto.create_octree_from_voxel_grid(g)
    tris = rto.marching_cubes_mesh(g, nodes)
    dev, host = rto.Scene.bvh_device(tris), rto.Scene.bvh(tris)
    rng = np.random.default_rng(3)
    n = 20000
    o = (rng.random((n, 3)).astype(np.float32) * 3 - 1.5).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[::2, rng.integers(0, 3)] = 0.0                       # every second ray has a zero component: reciprocal is infinite
    d[::5] = np.array([0, 0, -1], np.float32)              # axis-parallel rays
    o[::7] *= 4000.0                                       # origins thousands of scene extents away
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for flags in (0, rto.FLAG_NO_PRUNE):
        td, idd = dev.trace_rays(o, d, rto.MODE_BVH, flags=flags)
        th, idh = host.trace_rays(o, d, rto.MODE_BVH, flags=flags)
        assert (idh >= 0).sum() > 500
        assert np.array_equal(idd >= 0, idh >= 0)
        assert_bit_equal(td, th, "t flags %d" % flags)
        assert (idd != idh).mean() < 1e-3
