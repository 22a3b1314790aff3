#!/usr/bin/env python
"""Scene construction: host builders vs the GPU builders (rto_build.cu) on the configs' grids; checks equality and prints times.
    python tools/bench_builders.py [--only dt,c3,c4] [--out FILE]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracing_octrees_b200 as rto

DT_GRID = os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz")


def t(fn, reps=1):
    best, out = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="dt,c3")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    assert rto.lib().rto_init(0) == 0
    rto.create_octree_on_device(rto.generate_test_volume(16))        # CUDA context + module load outside the timings
    rep = {}
    for name in a.only.split(","):
        if name == "dt":
            g = rto.VoxelGrid.load(DT_GRID)
        elif name == "c3":
            g = rto.city_block_grid(512, 1234, 32)
        elif name == "c4":
            g = rto.city_block_grid(1024, 4321, 64)
        elif name == "c1":
            g = rto.generate_test_volume(128)
        else:
            raise SystemExit("unknown grid " + name)
        r = {"dims": list(g.dims), "voxels_MB": g.data.size / 1e6}
        th, nodes_h = t(lambda: rto.create_octree_from_voxel_grid(g))
        td, nodes_d = t(lambda: rto.create_octree_on_device(g), 2)
        r["octree_nodes"] = int(len(nodes_h)); r["octree_host_s"] = th; r["octree_device_s_incl_60B_per_node_readback"] = td
        r["octree_equal"] = bool(np.array_equal(nodes_h, nodes_d))
        ts, sc = t(lambda: rto.Scene.octree_from_grid(g), 2)
        r["octree_scene_from_grid_s"] = ts
        th2, sc2 = t(lambda: rto.Scene.octree(nodes_h, g.min, g.voxel_size))
        r["octree_scene_from_host_nodes_s"] = th2
        del nodes_d, sc, sc2
        tm, tris_h = t(lambda: rto.marching_cubes_mesh(g, nodes_h))
        tmd, tris_d = t(lambda: rto.marching_cubes_mesh_on_device(g), 2)
        r["triangles"] = int(len(tris_h)); r["mc_host_s"] = tm; r["mc_device_s_incl_readback"] = tmd
        r["mc_equal"] = bool(tris_h.shape == tris_d.shape and np.array_equal(tris_h.view(np.uint32), tris_d.view(np.uint32)))
        rep[name] = r
        print(name, json.dumps(r), flush=True)
    if a.out:
        json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
