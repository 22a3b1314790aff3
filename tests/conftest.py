import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rto():
    """The product package with librto.so built in-tree (nvcc cross-compiles without a GPU)."""
    from ray_tracing_octrees_b200 import build
    build.build()
    import ray_tracing_octrees_b200 as pkg
    pkg.lib()
    return pkg


@pytest.fixture(scope="session")
def gpu(rto):
    """The product package on cuda:0 (fails, not skips, without a device: -m gpu tests are the parity tests proper)."""
    rc = rto.lib().rto_init(0)
    assert rc == 0, rto.lib().rto_last_error()
    return rto


@pytest.fixture(scope="session")
def dt_scene(gpu, dt_grid_path):
    """The DT Calgary scene of C2 / C5: the reference's voxel grid -> octree -> Marching-Cubes mesh -> BVH scene + octree scene."""
    rto = gpu
    grid = rto.VoxelGrid.load(dt_grid_path)
    nodes = rto.create_octree_from_voxel_grid(grid)
    tris = rto.marching_cubes_mesh(grid, nodes)
    return dict(grid=grid, nodes=nodes, tris=tris, oct=rto.Scene.octree(nodes, grid.min, grid.voxel_size), bvh=rto.Scene.bvh(tris))


@pytest.fixture(scope="session")
def golden_fullsize():
    return json.load(open(os.path.join(GOLDEN, "golden_fullsize.json")))


@pytest.fixture(scope="session")
def port():
    from oracle import bind
    return bind.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import bind
    if not bind.ref_available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    return bind.ref()


@pytest.fixture(scope="session")
def checker():
    """Strongest CPU checker available: compiled reference if present, else the port."""
    from oracle import bind
    return bind.best()


@pytest.fixture(scope="session")
def golden_meta():
    return json.load(open(os.path.join(GOLDEN, "golden_meta.json")))


@pytest.fixture(scope="session")
def golden_sphere32():
    return dict(np.load(os.path.join(GOLDEN, "golden_sphere32.npz")))


@pytest.fixture(scope="session")
def golden_dt():
    return dict(np.load(os.path.join(GOLDEN, "golden_dt_rows.npz")))


@pytest.fixture(scope="session")
def dt_grid_path():
    return os.path.join(GOLDEN, "dt_sceneCache.bin.gz")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bit_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    bad = bits(a) != bits(b)
    assert not bad.any(), "%s: %d of %d elements differ (first at %s)" % (what, int(bad.sum()), bad.size, np.argwhere(bad)[0])


def cam_from_dict(cls, d):
    cam = cls()
    for i in range(3):
        cam.camPos[i] = d["camPos"][i]
    for i in range(16):
        cam.invView[i] = d["invView"][i]
    cam.tanHalfFov, cam.aspect, cam.width, cam.height = d["tanHalfFov"], d["aspect"], d["width"], d["height"]
    return cam
