"""CPU only: the product's traversal code itself (rto_kernels.cuh, compiled for the host by tests/emu) against the oracle.
The kernels' functions are written host+device, so their logic -- both octree modes in the per-node and the batched
("fast") form, the general 64-byte layout, the pruned and the exact BVH traversal, shadow rays -- is checked here in the
GPU-less container on small scenes; the -m gpu tests then check the same code as compiled for sm_100a."""
import numpy as np
import pytest

from conftest import assert_bit_equal, cam_from_dict


@pytest.fixture(scope="module")
def emu(rto):
    import emu as E
    E.lib()
    return E


def _same(got, want, what):
    assert np.array_equal(got["id"], want["id"]), "%s: %d ids differ" % (what, int((got["id"] != want["id"]).sum()))
    assert_bit_equal(got["t"], want["t"], what + " t")
    assert_bit_equal(got["rgba"], want["rgba"], what + " rgba")


@pytest.mark.parametrize("name", ["a", "b", "inside"])
def test_emu_octree_golden_sphere32(rto, emu, golden_sphere32, golden_meta, name):
    g = golden_sphere32
    grid = rto.generate_test_volume(32)
    nodes = rto.create_octree_from_voxel_grid(grid)
    oc = emu.Octree(nodes, grid.min, grid.voxel_size)
    cam = cam_from_dict(rto.RtoCamera, golden_meta["sphere32_cams"][name])
    for mode, key in ((rto.MODE_OCTREE_SKIP, 0), (rto.MODE_OCTREE_GLSL, 1)):
        want = {k: g["oct%d_%s_%s" % (key, name, k)] for k in ("rgba", "id", "t")}
        slow = oc.render(cam, mode, count=True)
        _same(slow, want, "per-node mode %d" % key)
        assert slow["visits"] == int(g["oct%d_%s_stats" % (key, name)][0])
        _same(oc.render(cam, mode, count=False), want, "batched mode %d" % key)


@pytest.mark.parametrize("dims,fill,seed", [((20, 13, 7), 0.35, 1), ((16, 16, 16), 0.08, 2), ((1, 1, 1), 1.0, 3), ((33, 9, 17), 0.6, 4), ((8, 8, 8), 0.0, 5), ((40, 40, 3), 0.9, 6)])
def test_emu_random_grids(rto, emu, checker, dims, fill, seed):
    rng = np.random.default_rng(seed)
    data = (rng.random(dims[0] * dims[1] * dims[2]) < fill).astype(np.uint8)
    grid = rto.VoxelGrid(dims, (-1.5, 0.25, 3.0), 0.37, data)
    nodes = rto.create_octree_from_voxel_grid(grid)
    oc_ref = checker.octree(grid.dims, grid.min, grid.voxel_size, grid.data)
    oc_ref.build()
    oc = emu.Octree(nodes, grid.min, grid.voxel_size)
    tris = rto.marching_cubes_mesh(grid, nodes)
    m_ref = oc_ref.mesh()
    m_ref.build()
    bv = emu.Bvh(tris)
    ext = max(dims) * grid.voxel_size
    tgt = tuple(float(grid.min[i] + dims[i] * grid.voxel_size / 2) for i in range(3))
    # the third camera sits exactly on voxel planes (integer multiples of the voxel size from the grid corner): zero slab distances
    for (th, ph, r, w, h, t) in [(25, 130, 2.2 * ext, 96, 72, tgt), (-40, 10, 0.8 * ext, 64, 48, tgt), (0, 0, 0.37 * 4, 48, 48, (-1.5 + 0.37 * 2, 0.25 + 0.37 * 2, 3.0 + 0.37))]:
        cam, _ = rto.Camera.from_degrees(th, ph, r, t).consts(45.0, float(np.float32(w) / np.float32(h)), w, h)
        rcam, _ = checker.camera(th, ph, r, target=t, width=w, height=h)
        assert bytes(cam) == bytes(rcam)
        for mode, key in ((rto.MODE_OCTREE_SKIP, 0), (rto.MODE_OCTREE_GLSL, 1)):
            want = oc_ref.render(rcam, key, stats=True)
            slow = oc.render(cam, mode, count=True)
            _same(slow, want, "per-node mode %d" % key)
            assert slow["visits"] == int(want["stats"][0])
            _same(oc.render(cam, mode, count=False), want, "batched mode %d" % key)
        for flags in (0, 1):
            want = m_ref.render(rcam, flags, 1e-3 * grid.voxel_size)
            _same(bv.render(cam, flags | rto.FLAG_NO_PRUNE, 1e-3 * grid.voxel_size), want, "bvh exact")
            _same(bv.render(cam, flags, 1e-3 * grid.voxel_size), want, "bvh pruned (SAH topology)")


def test_emu_general_layout(rto, emu):
    grid = rto.generate_test_volume(16)
    nodes = rto.create_octree_from_voxel_grid(grid)
    rng = np.random.default_rng(5)
    perm = np.concatenate([[0], 1 + rng.permutation(len(nodes) - 1)])
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    shuffled = nodes[perm].copy()
    ch = shuffled[:, 7:15]
    ch[ch >= 0] = inv[ch[ch >= 0]]
    a, b = emu.Octree(nodes, grid.min, grid.voxel_size), emu.Octree(shuffled, grid.min, grid.voxel_size)
    cam, _ = rto.Camera.from_degrees(20, 50, 1.4).consts(45.0, 1.5, 90, 60)
    for mode in (rto.MODE_OCTREE_SKIP, rto.MODE_OCTREE_GLSL):
        oa, ob = a.render(cam, mode), b.render(cam, mode, count=True)
        assert_bit_equal(oa["t"], ob["t"], "t")
        hit = oa["id"] >= 0
        assert np.array_equal(hit, ob["id"] >= 0) and np.array_equal(oa["id"][hit], perm[ob["id"][hit]])


def test_emu_axis_parallel_and_zero_direction_rays(rto, emu, checker):
    """1/0 directions: mode B is unguarded (inf / NaN slab distances), mode A clamps to +-1e10 -- both must follow the oracle."""
    grid = rto.generate_test_volume(16)
    nodes = rto.create_octree_from_voxel_grid(grid)
    oc = emu.Octree(nodes, grid.min, grid.voxel_size)
    oc_ref = checker.octree(grid.dims, grid.min, grid.voxel_size, grid.data)
    oc_ref.build()
    rng = np.random.default_rng(3)
    o = rng.uniform(-1, 1, (512, 3)).astype(np.float32)
    o[::5] = np.round(o[::5] * 16) / 16                      # origins on voxel planes
    d = rng.normal(0, 1, (512, 3)).astype(np.float32)
    d[::4, 0] = 0.0; d[1::4, 1] = 0.0; d[2::8, 2] = -0.0; d[3::16] = [0, 0, 1]
    want_t, want_id = oc_ref.rayskip(o, d, 0.0, 1e30)
    for count in (True, False):
        t, ids = oc.trace(o, d, rto.MODE_OCTREE_SKIP, count=count)
        assert_bit_equal(t, want_t, "octreeRaySkip t"); assert np.array_equal(ids, want_id)
    ta, ia = oc.trace(o, d, rto.MODE_OCTREE_GLSL, count=True)
    tb, ib = oc.trace(o, d, rto.MODE_OCTREE_GLSL, count=False)
    assert_bit_equal(ta, tb, "mode B per-node vs batched"); assert np.array_equal(ia, ib)


def test_emu_wide_quantised_tree_gives_the_same_frames(rto, emu, checker, monkeypatch):
    """The 4-wide form of the production tree (16-bit boxes on one grid, dequantised by one fused multiply-add per plane) decides
    nothing by itself: candidates are settled at the leaves, so ids, t and colours equal the binary form's -- and the oracle's -- bit
    for bit, from outside, from inside the mesh, along axes, and for scenes far from the origin (where the margin of the quantised
    boxes is set by the size of the coordinates, not by the size of the scene)."""
    monkeypatch.setenv("RTO_BVH_WIDE", "1")
    for shift, dim in (((0.0, 0.0, 0.0), 24), ((5.0e4, -3.0e4, 7.0e4), 16)):
        grid = rto.generate_test_volume(dim)
        gm = tuple(float(np.float32(grid.min[i]) + np.float32(shift[i])) for i in range(3))
        grid = rto.VoxelGrid(grid.dims, gm, grid.voxel_size, grid.data)
        nodes = rto.create_octree_from_voxel_grid(grid)
        tris = rto.marching_cubes_mesh(grid, nodes)
        bv = emu.Bvh(tris)
        assert bv.wide_nodes > len(tris) // 8, "no wide tree was built"
        m_ref = checker.mesh(tris); m_ref.build()
        tgt = tuple(gm[i] + 0.5 for i in range(3))
        bias = 1e-3 * grid.voxel_size
        for (th, ph, r, w, h) in ((30, 40, 1.2, 80, 60), (0, 0, 1.5, 48, 48), (-60, 200, 0.3, 48, 36), (89, 10, 2.0, 40, 40)):
            cam, _ = rto.Camera.from_degrees(th, ph, r, tgt).consts(45.0, float(np.float32(w) / np.float32(h)), w, h)
            rcam, _ = checker.camera(th, ph, r, target=tgt, width=w, height=h)
            for flags in (0, 1):
                want = m_ref.render(rcam, flags, bias)
                _same(bv.render(cam, flags, bias), want, "binary form")
                _same(bv.render(cam, flags | emu.Bvh.WIDE, bias), want, "wide form, shift %s camera %s" % (shift, (th, ph, r)))
        m_ref.free()


def test_emu_surface_area_rebuild_of_radix_subtrees(rto, emu):
    """rto_sahchunk.h, the per-thread surface-area rebuild of the bottom of the device-built BVH, on the CPU: for every size and for both
    places the subtree's root can sit, the result is a binary tree over exactly the given leaves inside exactly the slots the radix
    subtree owned (checked by emu_sah_chunk), and its boxes are smaller in total than those of halving the Morton-sorted list."""
    rng = np.random.default_rng(3)

    def halving_cost(boxes):
        if len(boxes) == 1:
            return 0.0
        lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
        d = (hi - lo).astype(np.float64)
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0] + halving_cost(boxes[:len(boxes) // 2]) + halving_cost(boxes[len(boxes) // 2:])
    for m in (2, 3, 4, 5, 7, 16, 33, 100, 255, 256, 1000, 4096):
        for kind in ("random", "plane", "identical", "line"):
            c = rng.random((m, 3)).astype(np.float32)
            if kind == "plane":
                c[:, 1] = 0.25
            elif kind == "identical":
                c[:] = 0.5
            elif kind == "line":
                c[:, 1:] = c[:, :1]
            h = (rng.random((m, 3)) * 0.02).astype(np.float32)
            boxes = np.concatenate([c - h, c + h], 1).astype(np.float32)
            for end in (False, True):
                cost = emu.sah_chunk(boxes, first=int(rng.integers(0, 50)), root_at_end=end)
                assert cost >= 0, "broken tree (code %g) for m=%d %s root_at_end=%s" % (cost, m, kind, end)
                if kind == "random" and m >= 16:
                    assert cost < 0.9 * halving_cost(boxes), "surface-area splits no better than halving a random list (m=%d)" % m
