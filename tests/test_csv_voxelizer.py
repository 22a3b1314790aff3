"""The CSV voxeliser: loadCSVDataIntoVoxelGrid of the reference (BuildingLoader.cpp:153-290, compiled in place
into oracle/_ref/libref.so) and the port's restatement against the product's host implementation (CPU) and device fill (GPU)."""
import numpy as np
import pytest

from csv_scenes import FACE_HEADER, VERT_HEADER, write_city_csv

CASES = [dict(seed=1, buildings=6, extent=150.0, voxel=5.0), dict(seed=2, buildings=20, extent=600.0, voxel=10.0),
         dict(seed=3, buildings=4, extent=80.0, voxel=1.5, messy=False), dict(seed=4, buildings=30, extent=9000.0, voxel=5.0)]   # the last one hits the 1000-cell cap


def _grid_of(oct_handle):
    return oct_handle.dims, np.asarray(oct_handle.gmin, np.float32), np.float32(oct_handle.voxel), oct_handle.grid_data()


def _assert_same(grid, want, what):
    dims, gmin, voxel, data = want
    assert grid is not None, what
    assert tuple(grid.dims) == tuple(dims), "%s: dims %s vs %s" % (what, grid.dims, dims)
    assert np.array_equal(np.asarray(grid.min, np.float32).view(np.uint32), gmin.view(np.uint32)), what + ": grid min"
    assert np.float32(grid.voxel_size).view(np.uint32) == voxel.view(np.uint32), what + ": voxel size"
    assert np.array_equal(grid.data, data), "%s: %d voxels differ" % (what, int((grid.data != data).sum()))


@pytest.mark.parametrize("case", CASES)
def test_host_voxeliser_equals_the_oracle(rto, checker, port, tmp_path, case):
    c = dict(case)
    voxel = c.pop("voxel")
    vp, fp = write_city_csv(tmp_path, **c)
    got = rto.load_csv_data_into_voxel_grid(vp, fp, voxel)
    want = _grid_of(checker.octree(csv=(vp, fp, voxel)))
    assert want[3].sum() > 0
    _assert_same(got, want, "host vs %s oracle" % checker.kind)
    _assert_same(got, _grid_of(port.octree(csv=(vp, fp, voxel))), "host vs port oracle")


def test_voxeliser_empty_and_missing_inputs(rto, checker, tmp_path):
    vp, fp = str(tmp_path / "v.csv"), str(tmp_path / "f.csv")
    open(vp, "w").write(VERT_HEADER)
    open(fp, "w").write(FACE_HEADER)
    assert rto.load_csv_data_into_voxel_grid(vp, fp, 5.0) is None                      # header only: empty grid
    assert checker.octree(csv=(vp, fp, 5.0)).dims == (0, 0, 0)
    assert rto.load_csv_data_into_voxel_grid(str(tmp_path / "nope.csv"), fp, 5.0) is None
    with pytest.raises(rto.RtoError):
        rto.load_csv_data_into_voxel_grid(vp, fp, 0.0)


def test_voxelised_grid_feeds_the_path(rto, tmp_path):
    """CSV -> grid -> octree -> sceneCache.bin round trip: the grid is a normal VoxelGrid for everything downstream."""
    vp, fp = write_city_csv(tmp_path, seed=5, buildings=8, extent=200.0)
    g = rto.load_csv_data_into_voxel_grid(vp, fp, 5.0)
    nodes = rto.create_octree_from_voxel_grid(g)
    assert len(nodes) > 1 and nodes[0][3] >= max(g.dims)
    path = str(tmp_path / "sceneCache.bin")
    g.save(path)
    back = rto.VoxelGrid.load(path)
    assert back.dims == g.dims and np.array_equal(back.data, g.data)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_voxeliser_equals_host(rto, tmp_path, case):
    assert rto.lib().rto_init(0) == 0
    c = dict(case)
    voxel = c.pop("voxel")
    vp, fp = write_city_csv(tmp_path, **c)
    host = rto.load_csv_data_into_voxel_grid(vp, fp, voxel)
    dev = rto.load_csv_data_into_voxel_grid(vp, fp, voxel, device=True)
    assert dev.dims == host.dims and np.array_equal(dev.min.view(np.uint32), host.min.view(np.uint32)) and dev.voxel_size == host.voxel_size
    assert np.array_equal(dev.data, host.data), "%d voxels differ" % int((dev.data != host.data).sum())
