"""GPU parity tests (run on the B200 box with -m gpu): the CUDA kernels, reached through the C ABI, against the CPU
oracle on the same inputs.  Bar: hit ids and leaf ids bit-exact; t and rgba bit-exact too (same IEEE operations in
the same order, no FMA) -- the 1e-4 tolerance of the north star is never needed.  The only documented exception is
the pruned BVH traversal at ill-conditioned grazing hits; tests allow <= 1e-4 of pixels there and
require RTO_FLAG_NO_PRUNE to be exact."""
import numpy as np
import pytest

from conftest import assert_bit_equal, cam_from_dict

pytestmark = pytest.mark.gpu


NEAR_TIE_FRAC = 1e-4        # BASELINE.json north_star: hit ids may differ at documented near-ties on at most 1e-4 of the pixels


def _cmp_frames(got, want, what, allow_frac=0.0, tally=None):
    """ids equal (or, with tally = [bad, total], counted for a cap over everything compared), t and rgba bit-equal where ids agree."""
    n = len(want["id"])
    bad = got["id"] != want["id"]
    frac = bad.sum() / max(n, 1)
    if tally is not None:
        tally[0] += int(bad.sum()); tally[1] += n
    else:
        assert frac <= allow_frac, "%s: %d of %d ids differ" % (what, int(bad.sum()), n)
    ok = ~bad
    assert_bit_equal(got["t"][ok], want["t"][ok], what + " t")
    assert_bit_equal(got["rgba"][ok], want["rgba"][ok], what + " rgba")
    return int(bad.sum())


# ---- golden vectors from the compiled reference -------------------------------------------------------------
@pytest.mark.parametrize("name", ["a", "b", "inside"])
def test_golden_sphere32(gpu, golden_sphere32, golden_meta, name):
    rto = gpu
    g = golden_sphere32
    grid = rto.generate_test_volume(32)
    nodes = rto.create_octree_from_voxel_grid(grid)
    assert np.array_equal(nodes, g["flat"])
    cam = cam_from_dict(rto.RtoCamera, golden_meta["sphere32_cams"][name])
    oc = rto.Scene.octree(nodes, grid.min, grid.voxel_size)
    assert oc.info()["compact"] == 1
    for mode, key in ((rto.MODE_OCTREE_SKIP, 0), (rto.MODE_OCTREE_GLSL, 1)):
        out = oc.render(cam, mode)
        want = {k: g["oct%d_%s_%s" % (key, name, k)] for k in ("rgba", "id", "t")}
        _cmp_frames(out, want, "octree mode %d" % key)
        assert oc.stats(cam, mode)[0] == g["oct%d_%s_stats" % (key, name)][0]
    tris = rto.marching_cubes_mesh(grid, nodes)
    sc = rto.Scene.bvh(tris)
    bias = 1e-3 * grid.voxel_size
    for flags in (0, 1):
        want = {k: g["bvh%d_%s_%s" % (flags, name, k)] for k in ("rgba", "id", "t")}
        _cmp_frames(sc.render(cam, rto.MODE_BVH, flags | rto.FLAG_NO_PRUNE, bias), want, "bvh exact flags %d" % flags)
        _cmp_frames(sc.render(cam, rto.MODE_BVH, flags, bias), want, "bvh pruned flags %d" % flags, allow_frac=1e-4)
        st = sc.stats(cam, rto.MODE_BVH, flags, bias)
        ws = g["bvh%d_%s_stats" % (flags, name)]
        assert st[0] == ws[0] and st[1] == ws[1]
        if flags:
            assert np.array_equal(st, ws)


def test_golden_sphere32_query_and_edge_rays(gpu, golden_sphere32):
    rto = gpu
    g = golden_sphere32
    grid = rto.generate_test_volume(32)
    nodes = rto.create_octree_from_voxel_grid(grid)
    tris = rto.marching_cubes_mesh(grid, nodes)
    sc = rto.Scene.bvh(tris)
    off, ids = sc.query(g["query_o"], g["query_d"])
    assert np.array_equal(off, g["query_off"]) and np.array_equal(ids, g["query_ids"])
    oc = rto.Scene.octree(nodes, grid.min, grid.voxel_size)
    t, ids = oc.trace_rays(g["edge_o"], g["edge_d"], rto.MODE_OCTREE_SKIP)
    assert_bit_equal(t, g["edge_t"], "octreeRaySkip on axis-parallel rays")
    assert np.array_equal(ids, g["edge_id"])


@pytest.mark.parametrize("name", ["far", "near"])
def test_golden_dt_rows(gpu, dt_scene, golden_dt, golden_meta, name):
    rto = gpu
    cam = cam_from_dict(rto.RtoCamera, golden_meta["dt_cams"][name])
    bias = 1e-3 * dt_scene["grid"].voxel_size
    for bi, (y0, y1) in enumerate(golden_dt["bands"]):
        y0, y1 = int(y0), int(y1)
        for mode, key in ((rto.MODE_OCTREE_SKIP, 0), (rto.MODE_OCTREE_GLSL, 1)):
            want = {k: golden_dt["oct%d_%s_%d_%s" % (key, name, bi, k)] for k in ("rgba", "id", "t")}
            _cmp_frames(dt_scene["oct"].render(cam, mode, y0=y0, y1=y1), want, "dt octree mode %d band %d" % (key, bi))
            assert dt_scene["oct"].stats(cam, mode, y0=y0, y1=y1)[0] == golden_dt["oct%d_%s_%d_stats" % (key, name, bi)][0]
        want = {k: golden_dt["bvh1_%s_%d_%s" % (name, bi, k)] for k in ("rgba", "id", "t")}
        _cmp_frames(dt_scene["bvh"].render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS | rto.FLAG_NO_PRUNE, bias, y0, y1), want, "dt bvh exact band %d" % bi)
        _cmp_frames(dt_scene["bvh"].render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, y0, y1), want, "dt bvh pruned band %d" % bi, allow_frac=1e-4)
        assert np.array_equal(dt_scene["bvh"].stats(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, y0, y1), golden_dt["bvh1_%s_%d_stats" % (name, bi)])


# ---- differential tests against the live oracle on seeded inputs --------------------------------------------
@pytest.mark.parametrize("dims,fill,seed", [((20, 13, 7), 0.35, 1), ((16, 16, 16), 0.08, 2), ((1, 1, 1), 1.0, 3), ((33, 9, 17), 0.6, 4), ((8, 8, 8), 0.0, 5)])
def test_random_grids_vs_oracle(gpu, checker, dims, fill, seed):
    rto = gpu
    rng = np.random.default_rng(seed)
    data = (rng.random(dims[0] * dims[1] * dims[2]) < fill).astype(np.uint8)
    grid = rto.VoxelGrid(dims, (-1.5, 0.25, 3.0), 0.37, data)
    nodes = rto.create_octree_from_voxel_grid(grid)
    oc_ref = checker.octree(grid.dims, grid.min, grid.voxel_size, grid.data)
    oc_ref.build()
    ext = max(dims) * grid.voxel_size
    tgt = tuple(float(grid.min[i] + dims[i] * grid.voxel_size / 2) for i in range(3))
    oc = rto.Scene.octree(nodes, grid.min, grid.voxel_size)
    tris = rto.marching_cubes_mesh(grid, nodes)
    sc = rto.Scene.bvh(tris)
    m_ref = oc_ref.mesh()
    m_ref.build()
    for (th, ph, r, w, h) in [(25, 130, 2.2 * ext, 160, 120), (-40, 10, 0.8 * ext, 120, 90), (5, 260, 0.2 * ext, 64, 64)]:
        cam, _ = rto.Camera.from_degrees(th, ph, r, tgt).consts(45.0, float(np.float32(w) / np.float32(h)), w, h)
        rcam, _ = checker.camera(th, ph, r, target=tgt, width=w, height=h)
        assert bytes(cam) == bytes(rcam)
        for mode, key in ((rto.MODE_OCTREE_SKIP, 0), (rto.MODE_OCTREE_GLSL, 1)):
            want = oc_ref.render(rcam, key, stats=True)
            _cmp_frames(oc.render(cam, mode), want, "octree mode %d" % key)
            assert oc.stats(cam, mode)[0] == want["stats"][0]
        for flags in (0, 1):
            want = m_ref.render(rcam, flags, 1e-3 * grid.voxel_size, stats=True)
            _cmp_frames(sc.render(cam, rto.MODE_BVH, flags | rto.FLAG_NO_PRUNE, 1e-3 * grid.voxel_size), want, "bvh exact")
            _cmp_frames(sc.render(cam, rto.MODE_BVH, flags, 1e-3 * grid.voxel_size), want, "bvh pruned", allow_frac=1e-4)
            if len(tris):
                st = sc.stats(cam, rto.MODE_BVH, flags, 1e-3 * grid.voxel_size)
                assert st[0] == want["stats"][0] and st[1] == want["stats"][1]


def test_general_octree_layout_matches_compact(gpu, checker):
    """Arrays that do not have the reference builder's shape (here: node order shuffled, a child removed) take the
    64-byte general path; on the builder's own array both paths must agree with each other and the oracle."""
    rto = gpu
    grid = rto.generate_test_volume(16)
    nodes = rto.create_octree_from_voxel_grid(grid)
    # permute node storage order (keeps the tree, breaks sibling contiguity) -> general layout; ids are permuted too
    rng = np.random.default_rng(5)
    perm = np.concatenate([[0], 1 + rng.permutation(len(nodes) - 1)])          # new index -> old index
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    shuffled = nodes[perm].copy()
    ch = shuffled[:, 7:15]
    ch[ch >= 0] = inv[ch[ch >= 0]]
    a = rto.Scene.octree(nodes, grid.min, grid.voxel_size)
    b = rto.Scene.octree(shuffled, grid.min, grid.voxel_size)
    assert a.info()["compact"] == 1 and b.info()["compact"] == 0
    cam, _ = rto.Camera.from_degrees(20, 50, 1.4).consts(45.0, 1.5, 150, 100)
    for mode in (rto.MODE_OCTREE_SKIP, rto.MODE_OCTREE_GLSL):
        oa, ob = a.render(cam, mode), b.render(cam, mode)
        assert_bit_equal(oa["t"], ob["t"], "t"); assert_bit_equal(oa["rgba"], ob["rgba"], "rgba")
        hit = oa["id"] >= 0
        assert np.array_equal(hit, ob["id"] >= 0)
        assert np.array_equal(oa["id"][hit], perm[ob["id"][hit]])
        assert a.stats(cam, mode)[0] == b.stats(cam, mode)[0]


def test_bvh_query_matches_oracle(gpu, checker):
    rto = gpu
    rng = np.random.default_rng(9)
    tris = rng.normal(0, 1, (500, 9)).astype(np.float32) * 0.2 + rng.normal(0, 1, (500, 1)).astype(np.float32)
    bvh = rto.BVH(tris)
    m = checker.mesh(tris)
    m.build()
    o = rng.normal(0, 3, (300, 3)).astype(np.float32)
    d = rng.normal(0, 1, (300, 3)).astype(np.float32)
    d[::7, 0] = 0.0                       # 1/0 = inf directions (BVH::query has no guard, BVH.cpp:110)
    d[3::11, 1] = -0.0
    off, ids = bvh.query_batch(o, d)
    roff, rids = m.query(o, d)
    assert np.array_equal(off, roff) and np.array_equal(ids, rids)
    assert np.array_equal(bvh.query(o[0], d[0]), rids[roff[0]:roff[1]])


@pytest.mark.parametrize("n", [0, 1, 2, 3])
def test_tiny_meshes(gpu, checker, n):
    rto = gpu
    rng = np.random.default_rng(n)
    tris = (rng.normal(0, 1, (n, 9)) * 0.5).astype(np.float32)
    sc = rto.Scene.bvh(tris)
    m = checker.mesh(tris)
    m.build()
    cam, _ = rto.Camera.from_degrees(10, 20, 4.0).consts(45.0, 1.0, 64, 64)
    rcam, _ = checker.camera(10, 20, 4.0, width=64, height=64, aspect=1.0)
    want = m.render(rcam, 1, 1e-3)
    _cmp_frames(sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS | rto.FLAG_NO_PRUNE, 1e-3), want, "tiny mesh")
    _cmp_frames(sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3), want, "tiny mesh pruned")


def test_degenerate_and_duplicate_triangles(gpu, checker):
    rto = gpu
    rng = np.random.default_rng(21)
    base = rng.integers(-2, 3, (200, 3)).astype(np.float32)
    tris = np.concatenate([base, base + [1, 0, 0], base + [0, 1, 0]], axis=1).astype(np.float32)
    tris = np.concatenate([tris, tris[:50], np.zeros((5, 9), np.float32)], axis=0)      # duplicates + zero-area
    sc = rto.Scene.bvh(tris)
    m = checker.mesh(tris)
    m.build()
    cam, _ = rto.Camera.from_degrees(33, 47, 9.0).consts(45.0, 1.0, 128, 128)
    rcam, _ = checker.camera(33, 47, 9.0, width=128, height=128, aspect=1.0)
    want = m.render(rcam, 1, 1e-3)
    _cmp_frames(sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS | rto.FLAG_NO_PRUNE, 1e-3), want, "duplicates exact")
    _cmp_frames(sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-3), want, "duplicates pruned")   # ties: first candidate wins


def test_row_bands_and_batches_compose(gpu):
    """A frame rendered in bands / as part of a camera batch equals the frame rendered at once (sharding invariant)."""
    rto = gpu
    grid = rto.generate_test_volume(32)
    nodes = rto.create_octree_from_voxel_grid(grid)
    tris = rto.marching_cubes_mesh(grid, nodes)
    sc = rto.Scene.bvh(tris)
    cam, _ = rto.Camera.from_degrees(30, 40, 1.2).consts(45.0, 4 / 3, 200, 150)
    full = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-5)
    parts = [sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, 1e-5, y0, y1) for (y0, y1) in [(0, 37), (37, 38), (38, 150)]]
    for k in ("rgba", "id", "t"):
        assert_bit_equal(np.concatenate([p[k] for p in parts]), full[k], "bands " + k)


def test_error_paths(gpu):
    rto = gpu
    grid = rto.generate_test_volume(8)
    nodes = rto.create_octree_from_voxel_grid(grid)
    oc = rto.Scene.octree(nodes, grid.min, grid.voxel_size)
    cam, _ = rto.Camera.from_degrees(30, 40, 1.2).consts(45.0, 1.0, 32, 32)
    with pytest.raises(rto.RtoError):
        oc.render(cam, rto.MODE_BVH)                      # wrong mode for the scene kind
    with pytest.raises(rto.RtoError):
        oc.render(cam, rto.MODE_OCTREE_GLSL, y0=10, y1=5)
    with pytest.raises(rto.RtoError):
        rto.Scene.octree(nodes[:0], grid.min, grid.voxel_size)
    bad = nodes.copy(); bad[0, 7] = len(nodes) + 5
    with pytest.raises(rto.RtoError):
        rto.Scene.octree(bad, grid.min, grid.voxel_size)
    rt = rto.RayTracerBVH()
    assert rt.render_scene_compute(rto.Camera(0.5, 0.7, 1.2), 32, 32, 1.0, 45.0) is None     # no data: frame untouched


# ---- full-size properties (BASELINE.json configs) ---------------------------------------------------------------
def test_full_size_dt_1080p(gpu, dt_scene, checker, dt_grid_path):
    """C2 at full size: the whole 1920x1080 frame, primary + shadow, vs the oracle on every 16th scanline band, plus
    size-independent properties on the full frame."""
    rto = gpu
    cam, _ = rto.Camera.from_degrees(35, 40, 0.6 * 4250).consts(45.0, float(np.float32(1920) / np.float32(1080)), 1920, 1080)
    bias = 1e-3 * dt_scene["grid"].voxel_size
    full = dt_scene["bvh"].render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
    exact = dt_scene["bvh"].render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS | rto.FLAG_NO_PRUNE, bias)
    n = 1920 * 1080
    mism = int((full["id"] != exact["id"]).sum())
    assert mism <= 1e-4 * n, "pruned traversal differs from the exact candidate replay on %d pixels" % mism
    hit = full["id"] >= 0
    assert hit.sum() > 0.3 * n
    assert ((full["t"][hit] > 1e-4) & (full["t"][hit] < 1e30)).all() and (full["t"][~hit] == np.float32(1e30)).all()
    assert (full["rgba"][:, 3] == 1.0).all() and (full["rgba"][~hit, :3] == 0).all()
    assert full["id"].max() < len(dt_scene["tris"])
    # the hit point lies on the reported triangle's plane (independent check of id/t consistency)
    # oracle on scanline bands
    from oracle import bind
    oc_ref = checker.octree(*bind.load_scene_cache(dt_grid_path))
    oc_ref.build()
    m_ref = oc_ref.mesh(); m_ref.build()
    rcam, _ = checker.camera(35, 40, 0.6 * 4250, width=1920, height=1080)
    assert bytes(rcam) == bytes(cam)
    tally = [0, 0]
    for y0 in range(8, 1080, 135):
        want = m_ref.render(rcam, 1, bias, y0, y0 + 2)
        sl = slice(y0 * 1920, (y0 + 2) * 1920)
        got = {k: exact[k][sl] for k in ("rgba", "id", "t")}
        _cmp_frames(got, want, "dt exact rows %d" % y0)
        got = {k: full[k][sl] for k in ("rgba", "id", "t")}
        _cmp_frames(got, want, "dt pruned rows %d" % y0, tally=tally)
    assert tally[0] <= NEAR_TIE_FRAC * tally[1], "pruned traversal vs oracle: %d of %d pixels differ" % tuple(tally)


def test_full_size_sphere128_and_octree_modes(gpu, checker):
    """C1 (sphere 128^3 MC mesh, 1024x768) and both octree modes at full size against the oracle on bands."""
    rto = gpu
    grid = rto.generate_test_volume(128)
    nodes = rto.create_octree_from_voxel_grid(grid)
    tris = rto.marching_cubes_mesh(grid, nodes)
    sc = rto.Scene.bvh(tris)
    oc = rto.Scene.octree(nodes, grid.min, grid.voxel_size)
    cam, _ = rto.Camera.from_degrees(30, 40, 1.2).consts(45.0, float(np.float32(1024) / np.float32(768)), 1024, 768)
    rcam, _ = checker.camera(30, 40, 1.2, width=1024, height=768)
    oc_ref = checker.octree(grid.dims, grid.min, grid.voxel_size, grid.data); oc_ref.build()
    m_ref = oc_ref.mesh(); m_ref.build()
    bias = 1e-3 * grid.voxel_size
    full = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
    fa, fb = oc.render(cam, rto.MODE_OCTREE_SKIP), oc.render(cam, rto.MODE_OCTREE_GLSL)
    # mode A returns a solid leaf at least as far as ... no ordering guarantee between modes, but both hit the same pixels
    assert np.array_equal(fa["id"] >= 0, fb["id"] >= 0)
    tally = [0, 0]
    for y0 in range(5, 768, 96):
        sl = slice(y0 * 1024, (y0 + 3) * 1024)
        _cmp_frames({k: full[k][sl] for k in full}, m_ref.render(rcam, 1, bias, y0, y0 + 3), "sphere bvh rows %d" % y0, tally=tally)
        _cmp_frames({k: fa[k][sl] for k in fa}, oc_ref.render(rcam, 0, y0, y0 + 3), "sphere mode A rows %d" % y0)
        _cmp_frames({k: fb[k][sl] for k in fb}, oc_ref.render(rcam, 1, y0, y0 + 3), "sphere mode B rows %d" % y0)
    assert tally[0] <= NEAR_TIE_FRAC * tally[1], "pruned traversal vs oracle: %d of %d pixels differ" % tuple(tally)


def test_city_block_octree_and_mesh_vs_oracle(gpu, checker):
    """C3-shaped input (synthetic city blocks; 192^3 here so that the oracle's top-down octree build stays in seconds), both
    octree modes and the mesh path against the oracle on scanline bands, from an oblique and from an axis-aligned camera
    (the latter puts the sign changes of 1/d -- the octant boundaries the kernels specialise on -- in the middle of the image)."""
    rto = gpu
    grid = rto.city_block_grid(192, 4242, 12)
    nodes = rto.create_octree_from_voxel_grid(grid)
    oc = rto.Scene.octree_from_grid(grid)
    tris = rto.marching_cubes_mesh(grid, nodes)
    sc = rto.Scene.bvh(tris)
    oc_ref = checker.octree(grid.dims, grid.min, grid.voxel_size, grid.data); oc_ref.build()
    m_ref = oc_ref.mesh(); m_ref.build()
    W, H = 640, 360
    bias = 1e-3 * grid.voxel_size
    tally = [0, 0]
    for theta, phi in ((35, 40), (20, 90), (60, 180), (1, 0)):
        cam, _ = rto.Camera.from_degrees(theta, phi, 0.9 * 192).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)
        rcam, _ = checker.camera(theta, phi, 0.9 * 192, width=W, height=H)
        assert bytes(rcam) == bytes(cam)
        fa, fb = oc.render(cam, rto.MODE_OCTREE_SKIP), oc.render(cam, rto.MODE_OCTREE_GLSL)
        fm = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
        for y0 in range(2, H, 45):
            sl = slice(y0 * W, (y0 + 2) * W)
            _cmp_frames({k: fa[k][sl] for k in fa}, oc_ref.render(rcam, 0, y0, y0 + 2), "city mode A rows %d cam %d/%d" % (y0, theta, phi))
            _cmp_frames({k: fb[k][sl] for k in fb}, oc_ref.render(rcam, 1, y0, y0 + 2), "city mode B rows %d cam %d/%d" % (y0, theta, phi))
            _cmp_frames({k: fm[k][sl] for k in fm}, m_ref.render(rcam, 1, bias, y0, y0 + 2), "city mesh rows %d cam %d/%d" % (y0, theta, phi), tally=tally)
    assert tally[0] <= NEAR_TIE_FRAC * tally[1], "pruned traversal vs oracle: %d of %d pixels differ" % tuple(tally)


def test_dual_contouring_mesh_renders_like_the_oracle(gpu, checker, dt_grid_path):
    """C4-shaped path at a size the oracle finishes in seconds: Dual-Contouring soups (rto_host_dc_mesh, bit-identical to the reference's
    mesher: tests/test_dc_mesh.py) of a 192^3 city-block grid and of the DT grid through the BVH kernel with shadows, against the
    oracle's BVH::query + Moller-Trumbore on the same triangles.  DC soups are harder on the tie rules than Marching-Cubes soups:
    quads are split along a diagonal both triangles share and neighbouring cells emit coincident triangles."""
    rto = gpu
    W, H = 640, 360
    for name, grid, radius in (("city", rto.city_block_grid(192, 777, 12), 0.9 * 192), ("dt", rto.VoxelGrid.load(dt_grid_path), 0.6 * 4250)):
        nodes = rto.create_octree_from_voxel_grid(grid)
        tris = rto.dual_contouring_mesh(grid, nodes)
        sc = rto.Scene.bvh(tris)
        m_ref = checker.mesh(tris); m_ref.build()
        bias = 1e-3 * grid.voxel_size
        tally = [0, 0]
        for theta, phi in ((35, 40), (60, 190)):
            cam, _ = rto.Camera.from_degrees(theta, phi, radius).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)
            rcam, _ = checker.camera(theta, phi, radius, width=W, height=H)
            fm = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
            fx = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS | rto.FLAG_NO_PRUNE, bias)
            assert (fm["id"] >= 0).mean() > 0.05
            for y0 in range(2, H, 45):
                sl = slice(y0 * W, (y0 + 2) * W)
                want = m_ref.render(rcam, 1, bias, y0, y0 + 2)
                _cmp_frames({k: fx[k][sl] for k in fx}, want, "%s DC exact replay rows %d cam %d/%d" % (name, y0, theta, phi))
                _cmp_frames({k: fm[k][sl] for k in fm}, want, "%s DC rows %d cam %d/%d" % (name, y0, theta, phi), tally=tally)
        assert tally[0] <= NEAR_TIE_FRAC * tally[1], "%s: pruned traversal vs oracle: %d of %d pixels differ" % (name, tally[0], tally[1])
        m_ref.free()


def test_deep_trees_stay_within_the_traversal_stack(gpu, monkeypatch):
    """The kernels keep a fixed number of postponed subtrees.  A SAH tree deeper than the builders' limit is replaced by a balanced
    median-split tree over the same triangles; forcing that fallback (RTO_BVH_MAX_DEPTH) must not change a single result."""
    rto = gpu
    grid = rto.generate_test_volume(32)
    tris = rto.marching_cubes_mesh(grid, rto.create_octree_from_voxel_grid(grid))
    cam, _ = rto.Camera.from_degrees(30, 40, 1.2).consts(45.0, 1.0, 160, 160)
    bias = 1e-3 * grid.voxel_size
    a = rto.Scene.bvh(tris).render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
    monkeypatch.setenv("RTO_BVH_MAX_DEPTH", "6")            # any SAH tree over 7 936 triangles is deeper than 6
    sc = rto.Scene.bvh(tris)
    monkeypatch.delenv("RTO_BVH_MAX_DEPTH")
    b = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
    c = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS | rto.FLAG_NO_PRUNE, bias)
    for k in ("id", "t", "rgba"):
        assert_bit_equal(a[k], b[k], "median-split fallback " + k)
        assert_bit_equal(a[k], c[k], "exact replay " + k)
    # strongly graded sizes (each triangle 1.2x smaller and closer to the origin than the last): whatever shape the builders give
    # the tree, pruned, exact and device-built traversals agree
    n = 64
    k = np.arange(n, dtype=np.float64)
    x = (1.2 ** -k * 1.0e5).astype(np.float32)
    s = (x * np.float32(0.15)).astype(np.float32)
    tris = np.zeros((n, 9), np.float32)
    tris[:, 0] = x; tris[:, 3] = x + s; tris[:, 6] = x; tris[:, 7] = s
    sc, dev = rto.Scene.bvh(tris), rto.Scene.bvh_device(tris)
    rng = np.random.default_rng(3)
    pick = rng.integers(0, n, 2048)
    target = np.stack([x[pick] + s[pick] * 0.25, s[pick] * 0.25, np.zeros(len(pick), np.float32)], 1).astype(np.float32)
    origin = (target + np.array([0.3, 0.2, 5.0], np.float32) * s[pick, None]).astype(np.float32)
    d = (target - origin).astype(np.float32)
    t0, i0 = sc.trace_rays(origin, d, rto.MODE_BVH, rto.FLAG_NO_PRUNE)
    t1, i1 = sc.trace_rays(origin, d, rto.MODE_BVH, 0)
    t2, i2 = dev.trace_rays(origin, d, rto.MODE_BVH, 0)
    assert (i0 >= 0).mean() > 0.5                           # (the rule's |det| >= 1e-8 rejects the very smallest triangles)
    assert np.array_equal(i0, i1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    assert np.array_equal(i0, i2) and np.array_equal(t0.view(np.uint32), t2.view(np.uint32))


def test_bvh_awkward_rays_match_the_exact_replay(gpu):
    """Rays that leave the fast paths of the production traversal -- zero, denormal-small and huge direction components (1/d zero,
    infinite or overflowing in o/d), origins far outside (beyond the fused tests' admission range), inside the mesh and on
    vertices -- must give exactly what the exact replay of BVH::query (RTO_FLAG_NO_PRUNE, reference tree, select-form tests) gives."""
    rto = gpu
    grid = rto.generate_test_volume(32)
    tris = rto.marching_cubes_mesh(grid, rto.create_octree_from_voxel_grid(grid))
    sc = rto.Scene.bvh(tris)
    rng = np.random.default_rng(11)
    n = 20000
    o = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    d = rng.normal(0, 1, (n, 3)).astype(np.float32)
    kinds = rng.integers(0, 8, n)
    ax = rng.integers(0, 3, n)
    small = np.array([0.0, -0.0, 1e-36, -1e-36, 1e-20, 1e-42, 3e38, -3e38], np.float32)
    for i in range(n):
        k = kinds[i]
        if k == 0:
            d[i, ax[i]] = small[rng.integers(0, 6)]                       # one (near-)zero component
        elif k == 1:
            d[i] = 0; d[i, ax[i]] = rng.choice([-1.0, 1.0])               # axis-parallel
        elif k == 2:
            o[i] *= np.float32(30.0 if i % 2 else 3000.0)                 # beyond the fused tests' admission range (16 scene extents), aimed at the scene
            d[i] = -o[i] + rng.normal(0, 0.3, 3)
        elif k == 3:
            o[i] = rng.uniform(-0.1, 0.1, 3)                              # inside the hollow sphere
        elif k == 4:
            o[i] = tris[rng.integers(0, len(tris)), 0:3]                  # on a vertex
        elif k == 5:
            d[i, ax[i]] = small[6 + rng.integers(0, 2)]                   # one huge component
        elif k == 6:
            o[i, ax[i]] = np.float32(3e37); d[i, ax[i]] = np.float32(-1.0); d[i, (ax[i] + 1) % 3] = np.float32(1e-30)   # o/d overflows
    t0, i0 = sc.trace_rays(o, d, rto.MODE_BVH, rto.FLAG_NO_PRUNE)
    t1, i1 = sc.trace_rays(o, d, rto.MODE_BVH, 0)
    assert (i0 >= 0).sum() > 1000
    bad = np.nonzero(i0 != i1)[0]
    assert len(bad) == 0, "%d rays differ, first: o=%s d=%s exact=(%d, %g) fast=(%d, %g)" % (len(bad), o[bad[0]], d[bad[0]], i0[bad[0]], t0[bad[0]], i1[bad[0]], t1[bad[0]])
    assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))


def test_ray_coherence_sorting_changes_nothing_but_the_order(gpu, dt_scene):
    """RTO_FLAG_SORT_RAYS traces an explicit ray list in the order of a device radix sort on (octant, Morton cell of the entry point,
    coarse direction); every result must land at its ray's own index with the same bits, for all three traversals, including rays
    that miss the scene box, start inside it, or have zero direction components."""
    rto = gpu
    g, sc, oc = dt_scene["grid"], dt_scene["bvh"], dt_scene["oct"]
    rng = np.random.default_rng(11)
    n = 200_000
    ext = np.array(g.dims, np.float32) * np.float32(g.voxel_size)
    lo = np.array(g.min, np.float32)
    o = (lo + (rng.random((n, 3)).astype(np.float32) * 3 - 1) * ext).astype(np.float32)            # inside and around the grid
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[::17, 0] = 0.0; d[::23, 1] = 0.0; d[::29, 2] = 0.0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for scene, mode in ((sc, rto.MODE_BVH), (oc, rto.MODE_OCTREE_SKIP), (oc, rto.MODE_OCTREE_GLSL)):
        t0, i0 = scene.trace_rays(o, d, mode)
        t1, i1 = scene.trace_rays(o, d, mode, flags=rto.FLAG_SORT_RAYS)
        assert (i0 >= 0).mean() > 0.02
        assert np.array_equal(i0, i1)
        assert_bit_equal(t0, t1, "t mode %d" % mode)
    t1, i1 = sc.trace_rays(o[:1], d[:1], rto.MODE_BVH, flags=rto.FLAG_SORT_RAYS)                     # a single ray is not sorted
    assert i1[0] == sc.trace_rays(o[:1], d[:1], rto.MODE_BVH)[1][0]


def test_device_radix_sort_is_a_stable_sort(gpu):
    """csrc/rto_sort.cuh (the hand-written sort behind RTO_FLAG_SORT_RAYS) through rto_device_sort_pairs: keys come back ascending and the
    values follow them in the order of a STABLE sort -- equal keys keep their input order -- for sizes around the 4096-key tiles and the
    512-key warp runs, for keys with few distinct values, one value, all 32 bits in use, and already sorted / reversed input."""
    import ctypes
    rto = gpu
    rng = np.random.default_rng(11)
    cases = []
    for n in (0, 1, 2, 3, 31, 32, 33, 511, 512, 513, 4095, 4096, 4097, 8192, 12345, 100000, 1 << 20, (1 << 21) + 777):
        cases.append(rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))
    cases.append(rng.integers(0, 3, 50000, dtype=np.uint64).astype(np.uint32))                      # three distinct keys
    cases.append(np.full(70000, 0xdeadbeef, np.uint32))                                             # one key
    cases.append((rng.integers(0, 256, 300000, dtype=np.uint64) << 24).astype(np.uint32))           # only the top digit differs
    cases.append(np.arange(200000, dtype=np.uint32))                                                # sorted
    cases.append(np.arange(200000, dtype=np.uint32)[::-1].copy())                                   # reversed
    cases.append((rng.integers(0, 1 << 16, 400000, dtype=np.uint64) * 65537 % (1 << 32)).astype(np.uint32))
    for keys in cases:
        n = len(keys)
        k = keys.copy(); v = np.arange(n, dtype=np.uint32)
        rc = rto.lib().rto_device_sort_pairs(k.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p), n)
        assert rc == 0, rto.lib().rto_last_error().decode()
        order = np.argsort(keys, kind="stable").astype(np.uint32)
        assert np.array_equal(v, order), "n = %d: permutation differs from a stable sort" % n
        assert np.array_equal(k, keys[order])
