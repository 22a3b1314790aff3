#!/usr/bin/env python
"""Diagnostic: GPU octree traversal vs the CPU emulation of the same code, per mode, on a tiny scene (prints mismatches)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ray_tracing_octrees_b200 as rto
import emu
assert rto.lib().rto_init(0) == 0
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = rto.generate_test_volume(dim)
nodes = rto.create_octree_from_voxel_grid(g)
print("nodes", len(nodes))
oc = rto.Scene.octree(nodes, g.min, g.voxel_size)
eo = emu.Octree(nodes, g.min, g.voxel_size)
cam, _ = rto.Camera.from_degrees(30, 40, 1.2).consts(45.0, 1.0, 32, 32)
for mode, name in ((rto.MODE_OCTREE_GLSL, "B"), (rto.MODE_OCTREE_SKIP, "A")):
    want = eo.render(cam, mode, count=False)
    got = oc.render(cam, mode)
    bad = np.nonzero(got["id"] != want["id"])[0]
    tb = np.nonzero(got["t"].view(np.uint32) != want["t"].view(np.uint32))[0]
    print("mode", name, "id mismatches", len(bad), "t mismatches", len(tb), "of", len(want["id"]))
    for i in bad[:5]:
        print("   pixel", i, "gpu id/t", got["id"][i], got["t"][i], "emu id/t", want["id"][i], want["t"][i])
    sys.stdout.flush()
for mode, name in ((rto.MODE_OCTREE_GLSL, "B"), (rto.MODE_OCTREE_SKIP, "A")):
    print("stats (per-node path) gpu", int(oc.stats(cam, mode)[0]), "emu", eo.render(cam, mode, count=True)["visits"])
tris = rto.marching_cubes_mesh(g, nodes)
sc = rto.Scene.bvh(tris)
eb = emu.Bvh(tris)
for flags in (0, 1, 2, 3):
    got = sc.render(cam, rto.MODE_BVH, flags, 1e-4); want = eb.render(cam, flags, 1e-4)
    print("bvh flags", flags, "id mismatches", int((got["id"] != want["id"]).sum()), "hits", int((want["id"] >= 0).sum()))
