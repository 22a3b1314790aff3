"""The Dual-Contouring test cases shared by tests/test_dc_mesh.py and tests/golden/make_golden_dc.py."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "sphere32": dict(kind="sphere", dim=32),
    "sphere64": dict(kind="sphere", dim=64),
    "sphere48_culled": dict(kind="sphere", dim=48, cam=(30, 40, 1.2), margin=0.02),
    "sphere48_inside": dict(kind="sphere", dim=48, cam=(10, 200, 0.3), margin=0.0),
    "noise_sparse": dict(kind="noise", dims=(37, 26, 28), p=0.02, seed=1),
    "noise_half": dict(kind="noise", dims=(23, 14, 9), p=0.5, seed=2),
    "noise_dense": dict(kind="noise", dims=(32, 39, 19), p=0.97, seed=3),
    "noise_thin": dict(kind="noise", dims=(3, 35, 6), p=0.3, seed=4),
    "boxes": dict(kind="boxes", dims=(67, 68, 58), seed=5),
    "boxes_ragged": dict(kind="boxes", dims=(68, 25, 20), seed=6),
    "one_voxel": dict(kind="full", dims=(1, 1, 1)),
    "full_ragged": dict(kind="full", dims=(5, 9, 2)),
    "dt": dict(kind="dt"),
    "dt_culled_far": dict(kind="dt", cam=(35, 40, 0.6 * 4250), margin=50.0),
    "dt_culled_near": dict(kind="dt", cam=(10, 0, 300.0), margin=50.0),
}
SLOW_IN_REFERENCE = {"dt", "dt_culled_far"}      # 45 s and 23 s in the reference's own code: checked through the golden sha only


def make_grid(case):
    from oracle import bind
    k = case["kind"]
    if k == "sphere":
        return bind.sphere_grid(case["dim"])
    if k == "dt":
        return bind.load_scene_cache(os.path.join(GOLDEN, "dt_sceneCache.bin.gz"))
    dims = case["dims"]
    n = dims[0] * dims[1] * dims[2]
    if k == "full":
        return dims, (0.0, 0.0, 0.0), 1.0, np.ones(n, np.uint8)
    rng = np.random.default_rng(case["seed"])
    if k == "noise":
        return dims, (-3.0, 1.5, 10.0), 0.37, (rng.random(n) < case["p"]).astype(np.uint8)
    d = np.zeros(dims[::-1], np.uint8)
    for _ in range(12):
        a = [int(rng.integers(0, m)) for m in dims]
        b = [int(rng.integers(1, 24)) for _ in dims]
        d[a[2]:a[2] + b[2], a[1]:a[1] + b[1], a[0]:a[0] + b[0]] = 1
    return dims, (-10.0, -2.0, 5.0), 1.0, d.ravel()


def view_proj_for(backend, case):
    """proj * view as renderOctree forms it (main.cpp:122-125): perspective(45 deg, aspect, 0.01, 5000) * camera.getView(); aspect 1.5."""
    if "cam" not in case:
        return None
    import ray_tracing_octrees_b200 as rto
    th, ph, r = case["cam"]
    _, view = backend.camera(th, ph, r, width=96, height=64, aspect=1.5)
    return rto.view_proj(view, 45.0, 1.5, 0.01, 5000.0)
