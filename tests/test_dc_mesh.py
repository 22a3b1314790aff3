"""Adaptive Dual Contouring mesh (rto_host_dc_mesh) against the reference's own AdaptiveDualContouringRenderer.cpp (compiled in place,
driven by renderOctree's traversal) and against golden checksums generated from it (tests/golden/make_golden_dc.py).  Bit-exact:
same triangles, same order.  The builder keeps the reference's visit order (its dual-vertex cache makes the mesh depend on it)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import assert_bit_equal
from dc_cases import CASES, GOLDEN, SLOW_IN_REFERENCE, make_grid, view_proj_for

META = json.load(open(os.path.join(GOLDEN, "golden_dc.json")))


def build(rto, case):
    dims, gmin, voxel, data = make_grid(case)
    g = rto.VoxelGrid(dims, gmin, voxel, data)
    return g, rto.create_octree_from_voxel_grid(g)


@pytest.mark.parametrize("algo", ["default", "replay"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_dc_mesh_equals_the_golden_checksums(rto, name, algo):
    """Both host formulations (order-free first-toucher minima; replay of the cache protocol in visit order) against the reference's output."""
    case, want = CASES[name], META[name]
    g, nodes = build(rto, case)
    vp = None if want["view_proj"] is None else np.array(want["view_proj"], np.float32)
    got = rto.dual_contouring_mesh(g, nodes, vp, case.get("margin", 50.0), algo=algo)
    assert len(nodes) == want["nodes"]
    assert len(got) == want["tris"]
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == want["sha"]
    if name == "sphere32":
        assert_bit_equal(got, np.load(os.path.join(GOLDEN, "golden_dc_sphere32.npz"))["tris"], "sphere32 DC triangles")


@pytest.mark.parametrize("name", sorted(set(CASES) - SLOW_IN_REFERENCE))
def test_dc_mesh_equals_the_compiled_reference(rto, ref, name):
    case = CASES[name]
    dims, gmin, voxel, data = make_grid(case)
    oc = ref.octree(dims, gmin, voxel, data); oc.build()
    g, nodes = build(rto, case)
    vp = view_proj_for(ref, case)
    if vp is not None:                       # the host view-projection equals the one the golden file recorded
        assert_bit_equal(vp, np.array(META[name]["view_proj"], np.float32), "view_proj")
    want = oc.dc_mesh(vp, case.get("margin", 50.0)).tris()
    got = rto.dual_contouring_mesh(g, nodes, vp, case.get("margin", 50.0))
    oc.free()
    assert_bit_equal(got, want, name)


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_restatement_equals_the_golden_checksums(port, name):
    """oracle/oracle_port.cpp's literal, sequential restatement of the mesher (the checker where libref.so is absent) is pinned to the
    reference's output too."""
    case, want = CASES[name], META[name]
    oc = port.octree(*make_grid(case)); oc.build()
    vp = None if want["view_proj"] is None else np.array(want["view_proj"], np.float32)
    got = oc.dc_mesh(vp, case.get("margin", 50.0)).tris()
    oc.free()
    assert len(got) == want["tris"]
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == want["sha"]


def test_dc_mesh_refuses_octrees_beyond_the_reference_key_range(rto):
    """Cell keys are x << 20 | y << 10 | z in the reference (AdaptiveDualContouringRenderer.cpp:553-555): they alias past 1024 voxels."""
    nodes = np.zeros((1, 15), np.int32)
    nodes[0, 3] = 2048; nodes[0, 4] = 1; nodes[0, 7:] = -1
    g = rto.VoxelGrid((1, 1, 1), (0, 0, 0), 1.0, np.zeros(1, np.uint8))
    with pytest.raises(rto.RtoError):
        rto.dual_contouring_mesh(g, nodes)


@pytest.mark.parametrize("seed", range(8))
def test_dc_formulations_agree_on_random_grids(rto, seed):
    """Noise grids make every boundary leaf a fallback candidate and chain the fallback rounds of the order-free formulation."""
    rng = np.random.default_rng(100 + seed)
    dims = tuple(int(x) for x in rng.integers(2, 48, 3))
    p = [0.01, 0.05, 0.2, 0.5, 0.7, 0.9, 0.99, 0.35][seed]
    g = rto.VoxelGrid(dims, (1.0, -2.0, 0.5), 0.25, (rng.random(dims[0] * dims[1] * dims[2]) < p).astype(np.uint8))
    nodes = rto.create_octree_from_voxel_grid(g)
    assert_bit_equal(rto.dual_contouring_mesh(g, nodes), rto.dual_contouring_mesh(g, nodes, algo="replay"), "seed %d" % seed)


def test_dc_mesh_of_nothing(rto):
    g = rto.VoxelGrid((4, 4, 4), (0, 0, 0), 1.0, np.zeros(64, np.uint8))
    assert len(rto.dual_contouring_mesh(g, np.zeros((0, 15), np.int32))) == 0
    assert len(rto.dual_contouring_mesh(g, rto.create_octree_from_voxel_grid(g))) == 0


@pytest.mark.parametrize("name", ["sphere32", "sphere48_culled", "noise_half", "noise_dense", "boxes_ragged", "dt_culled_near"])
def test_dc_normals_equal_the_reference_mctriangles(rto, ref, name):
    """rto_host_dc_mesh_normals: the normal stored with every MCTriangle (flat normal, negated for solid leaves; face normals for the
    fallback fans, signed zeros included), against the reference's own records."""
    case = CASES[name]
    dims, gmin, voxel, data = make_grid(case)
    oc = ref.octree(dims, gmin, voxel, data); oc.build()
    g, nodes = build(rto, case)
    vp = view_proj_for(ref, case)
    want = oc.dc_mesh_full(vp, case.get("margin", 50.0))
    oc.free()
    tris, normals = rto.dual_contouring_mesh_with_normals(g, nodes, vp, case.get("margin", 50.0))
    assert_bit_equal(tris, want[:, :9], name + " vertices")
    for k in range(3):
        assert_bit_equal(normals, want[:, 9 + 3 * k:12 + 3 * k], name + " normal %d" % k)


def test_triangle_cache_file_round_trip(rto, tmp_path):
    """saveTriangleCache / loadTriangleCache (main.cpp:27-67): size_t count + count x 72-byte MCTriangle."""
    case = CASES["sphere32"]
    g, nodes = build(rto, case)
    tris, normals = rto.dual_contouring_mesh_with_normals(g, nodes)
    path = str(tmp_path / "dc_triangles_1.bin")
    rto.save_triangle_cache(path, tris, normals)
    raw = open(path, "rb").read()
    want = np.concatenate([tris, normals, normals, normals], axis=1).astype(np.float32)
    assert raw == np.uint64(len(tris)).tobytes() + want.tobytes()
    t2, n2 = rto.load_triangle_cache(path)
    assert_bit_equal(t2, tris, "tris"); assert_bit_equal(n2, want[:, 9:], "normals")
    # no normals given: the flat normal of the geometry (what localMC stores with Marching-Cubes triangles)
    rto.save_triangle_cache(path, tris[:10])
    t3, n3 = rto.load_triangle_cache(path)
    e1, e2 = tris[:10, 3:6] - tris[:10, 0:3], tris[:10, 6:9] - tris[:10, 0:3]
    c = np.cross(e1.astype(np.float64), e2.astype(np.float64)); c /= np.linalg.norm(c, axis=1, keepdims=True)
    assert len(t3) == 10 and np.allclose(n3[:, :3], c, atol=1e-5) and np.array_equal(n3[:, :3], n3[:, 3:6])
    # truncated and missing files are errors, an empty cache is not
    open(path, "wb").write(raw[:100])
    with pytest.raises(rto.RtoError):
        rto.load_triangle_cache(path)
    with pytest.raises(rto.RtoError):
        rto.load_triangle_cache(str(tmp_path / "nope.bin"))
    rto.save_triangle_cache(path, np.zeros((0, 9), np.float32))
    assert len(rto.load_triangle_cache(path)[0]) == 0
