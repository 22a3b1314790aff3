#!/usr/bin/env python
"""Golden vectors for the Dual-Contouring mesh, generated from the COMPILED REFERENCE (oracle/_ref/libref.so: the reference's own
AdaptiveDualContouringRenderer.cpp driven by renderOctree's traversal, oracle/ref_harness.cpp::ref_dc_mesh_from_octree).

Run in the build container only (needs /root/reference; about 2 minutes):   python tests/golden/make_golden_dc.py
Outputs:  golden_dc_sphere32.npz  (the whole triangle soup of the 32^3 shell sphere)
          golden_dc.json          (triangle counts + sha256 of the soups of the larger cases, and the view-projection matrices used)
"""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import bind as B
from dc_cases import CASES, make_grid, view_proj_for

OUT = os.path.dirname(os.path.abspath(__file__))
R = B.ref()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


meta = {}
for name, case in CASES.items():
    dims, gmin, voxel, data = make_grid(case)
    oc = R.octree(dims, gmin, voxel, data); oc.build()
    vp = view_proj_for(R, case)
    tris = oc.dc_mesh(vp, case.get("margin", 50.0)).tris()
    meta[name] = dict(tris=int(len(tris)), sha=sha(tris), nodes=int(oc.num_nodes), view_proj=None if vp is None else [float(x) for x in vp])
    if name == "sphere32":
        np.savez_compressed(os.path.join(OUT, "golden_dc_sphere32.npz"), tris=tris)
    oc.free()
    print(name, meta[name]["tris"], meta[name]["sha"][:16])
json.dump(meta, open(os.path.join(OUT, "golden_dc.json"), "w"), indent=1, sort_keys=True)
