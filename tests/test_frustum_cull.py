"""Frustum culling of the flattened octree (the CPU part of RayTracerBVH::renderSceneComputeWithCulling, RayTracerBVH.cpp:724-813,
the variant the reference's main loop calls): the product's host and device culls against the reference's own code (compiled in
place; its SSBO upload is captured) and the port's restatement, then the culled array rendered through the GLSL-semantics kernel."""
import numpy as np
import pytest

CAMS = [(35, 40, 0.6 * 4250), (10, 0, 300.0), (60, 10, 0.35 * 4250), (5, 200, 50.0), (80, 123, 900.0)]


def _dt(rto, dt_grid_path):
    g = rto.VoxelGrid.load(dt_grid_path)
    return g, rto.create_octree_from_voxel_grid(g)


@pytest.mark.parametrize("cam", CAMS)
def test_host_cull_equals_the_oracles(rto, checker, port, dt_grid_path, cam):
    from oracle import bind
    g, nodes = _dt(rto, dt_grid_path)
    theta, phi, radius = cam
    aspect = float(np.float32(1920) / np.float32(1080))
    c, view = rto.Camera.from_degrees(theta, phi, radius).consts(45.0, aspect, 1920, 1080)
    vp = rto.view_proj(view, 45.0, aspect)
    culled, back = rto.frustum_cull(nodes, g, vp, 150.0)
    for orc in (checker, port):
        oc = orc.octree(*bind.load_scene_cache(dt_grid_path)); oc.build()
        want = oc.cull(theta, phi, radius, 45.0, aspect, 1920, 1080)
        assert len(want) == len(culled), "%s: %d vs %d nodes kept" % (orc.kind, len(want), len(culled))
        assert np.array_equal(want, culled), orc.kind
    assert 0 < len(culled) <= len(nodes)
    assert np.array_equal(nodes[back][:, :7], culled[:, :7])          # same nodes, in index order
    assert (np.diff(back) > 0).all()


def test_cull_small_grid_and_everything_outside(rto, checker):
    g = rto.generate_test_volume(32)
    nodes = rto.create_octree_from_voxel_grid(g)
    # the sphere grid spans 1 unit: with the reference's margin of 150 nothing is ever dropped ...
    c, view = rto.Camera.from_degrees(30, 40, 1.2).consts(45.0, 1.0, 64, 64)
    culled, back = rto.frustum_cull(nodes, g, rto.view_proj(view, 45.0, 1.0), 150.0)
    assert np.array_equal(culled, nodes) and np.array_equal(back, np.arange(len(nodes)))
    oc = checker.octree(g.dims, g.min, g.voxel_size, g.data); oc.build()
    assert np.array_equal(oc.cull(30, 40, 1.2, 45.0, 1.0, 64, 64), nodes)
    # ... with no margin and the camera looking away from the grid everything is
    c, view = rto.Camera.from_degrees(0, 180, 1.0, target=(0, 0, 50.0)).consts(45.0, 1.0, 64, 64)      # eye at z = 49 looking towards +z
    culled, back = rto.frustum_cull(nodes, g, rto.view_proj(view, 45.0, 1.0), 0.0)
    assert len(culled) == 0 and len(back) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("cam", CAMS[:3])
def test_device_cull_equals_host_and_culled_scene_renders_like_the_oracle(rto, checker, dt_grid_path, cam):
    from oracle import bind
    assert rto.lib().rto_init(0) == 0
    g, nodes = _dt(rto, dt_grid_path)
    theta, phi, radius = cam
    W, H = 480, 270
    aspect = float(np.float32(W) / np.float32(H))
    c, view = rto.Camera.from_degrees(theta, phi, radius).consts(45.0, aspect, W, H)
    vp = rto.view_proj(view, 45.0, aspect)
    host, hback = rto.frustum_cull(nodes, g, vp, 150.0)
    dev, dback = rto.frustum_cull(nodes, g, vp, 150.0, device=True)
    assert np.array_equal(host, dev) and np.array_equal(hback, dback)
    # what the compute shader would draw from the culled SSBO: the GLSL traversal over the culled array (general layout: it has holes)
    sc = rto.Scene.octree(host, g.min, g.voxel_size)
    got = sc.render(c, rto.MODE_OCTREE_GLSL)
    # culling never changes WHICH surface a pixel of this camera sees unless the 512-step budget bites: compare with the full array
    full = rto.Scene.octree(nodes, g.min, g.voxel_size).render(c, rto.MODE_OCTREE_GLSL)
    hit = got["id"] >= 0
    assert np.array_equal(hback[got["id"][hit]], full["id"][hit]) or (hback[got["id"][hit]] != full["id"][hit]).mean() < 0.02
