"""VolumeRaycastRenderer's per-frame use of octreeRaySkip (VolumeRaycastRenderer.cpp:1598-1664): 49 probe rays -> 15th percentile x 0.75
-> temporal blend.  The probe rays and the percentile/blend are host code (CPU tests against the oracle, whose 49 octreeRaySkip
calls are the reference's own function); the whole estimate with the rays traced on the GPU is a -m gpu test."""
import numpy as np
import pytest

CAMS = [(35, 40, 0.6 * 4250), (10, 0, 300.0), (60, 10, 0.35 * 4250), (5, 200, 50.0), (80, 123, 900.0)]


@pytest.mark.parametrize("cam", CAMS)
def test_probe_rays_and_percentile_equal_the_oracle(rto, checker, port, dt_grid_path, cam):
    from oracle import bind
    theta, phi, radius = cam
    aspect = float(np.float32(1300) / np.float32(1300))
    c, view = rto.Camera.from_degrees(theta, phi, radius).consts(45.0, aspect, 64, 64)
    o, d = rto.skip_probe_rays(view, np.array(list(c.camPos), np.float32), aspect)
    for orc in (checker, port):
        oc = orc.octree(*bind.load_scene_cache(dt_grid_path)); oc.build()
        last = 0.0
        for frame in range(3):                               # the temporal blend carries state from frame to frame
            want, t, wo, wd = oc.skip_distance(theta, phi, radius, aspect, last)
            assert np.array_equal(o.view(np.uint32), wo.view(np.uint32)), orc.kind
            assert np.array_equal(d.view(np.uint32), wd.view(np.uint32)), orc.kind
            got = rto.skip_distance_from_probes(t, last)
            assert np.float32(got).view(np.uint32) == np.float32(want).view(np.uint32), (orc.kind, got, want)
            last = want


@pytest.mark.gpu
@pytest.mark.parametrize("cam", CAMS)
def test_skip_distance_on_gpu(rto, checker, dt_grid_path, cam):
    from oracle import bind
    assert rto.lib().rto_init(0) == 0
    g = rto.VoxelGrid.load(dt_grid_path)
    sc = rto.Scene.octree_from_grid(g)
    oc = checker.octree(*bind.load_scene_cache(dt_grid_path)); oc.build()
    theta, phi, radius = cam
    aspect = 1.0
    c, view = rto.Camera.from_degrees(theta, phi, radius).consts(45.0, aspect, 64, 64)
    last = 0.0
    for frame in range(3):
        want, t, _, _ = oc.skip_distance(theta, phi, radius, aspect, last)
        got, probes = sc.skip_distance(view, np.array(list(c.camPos), np.float32), aspect, last)
        assert np.array_equal(probes.view(np.uint32), t.view(np.uint32))
        assert np.float32(got).view(np.uint32) == np.float32(want).view(np.uint32)
        last = want
