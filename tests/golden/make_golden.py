#!/usr/bin/env python
"""Regenerate the golden fixtures under tests/golden/ from the COMPILED REFERENCE (oracle/_ref/libref.so).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs (all small, committed):
  dt_sceneCache.bin.gz   the reference's own voxelised DT Calgary grid (/root/reference/sceneCache.bin), gzip -9,
                         byte-identical after decompression (sha256 recorded in golden_meta.json)
  golden_sphere32.npz    shell-sphere 32^3: flattened octree, MC mesh, BVH export, renders of all three modes
  golden_dt_rows.npz     DT scene at 1920x1080, two cameras, a few scanline bands, all three modes + work counters
  golden_meta.json       checksums of the big structures (sphere128 / DT octree, mesh, BVH) and camera constants
"""
import gzip, hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bind as B

OUT = os.path.dirname(os.path.abspath(__file__))
R = B.ref()
SRC = "/root/reference/sceneCache.bin"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cam_dict(cam):
    return dict(camPos=[float(x) for x in cam.camPos], invView=[float(x) for x in cam.invView],
                tanHalfFov=float(cam.tanHalfFov), aspect=float(cam.aspect), width=cam.width, height=cam.height)


meta = {}
raw = open(SRC, "rb").read()
with open(os.path.join(OUT, "dt_sceneCache.bin.gz"), "wb") as f:
    with gzip.GzipFile(fileobj=f, mode="wb", compresslevel=9, mtime=0) as z:
        z.write(raw)
meta["dt_sceneCache_sha256"] = hashlib.sha256(raw).hexdigest()

# ---- sphere 32: everything, small enough to store whole -------------------------------------------------
dims, gmin, vox, data = B.sphere_grid(32)
oc = R.octree(dims, gmin, vox, data); oc.build()
mesh = oc.mesh(); mesh.build()
boxes, bmeta = mesh.export()
store = dict(voxels=np.packbits(data), flat=oc.flat(), tris=mesh.tris(), bvh_boxes=boxes, bvh_meta=bmeta)
cams = {}
for name, (th, ph, r, w, h) in {"a": (30, 40, 1.2, 96, 64), "b": (-20, 200, 0.9, 64, 96), "inside": (10, 75, 0.25, 48, 48)}.items():
    cam, view = R.camera(th, ph, r, width=w, height=h)
    cams[name] = cam_dict(cam)
    store["view_" + name] = view
    for mode in (0, 1):
        o = oc.render(cam, mode, stats=True)
        for k in ("rgba", "id", "t"):
            store["oct%d_%s_%s" % (mode, name, k)] = o[k]
        store["oct%d_%s_stats" % (mode, name)] = o["stats"]
    for flags in (0, 1):
        o = mesh.render(cam, flags, 1e-3 * vox, stats=True)
        for k in ("rgba", "id", "t"):
            store["bvh%d_%s_%s" % (flags, name, k)] = o[k]
        store["bvh%d_%s_stats" % (flags, name)] = o["stats"]
# BVH::query candidates for a fan of rays
rng = np.random.default_rng(7)
qo = np.tile(np.array([[0.7, 0.6, 0.8]], np.float32), (64, 1))
qd = (-qo + rng.normal(0, 0.25, (64, 3))).astype(np.float32)
off, ids = mesh.query(qo, qd)
store.update(query_o=qo, query_d=qd, query_off=off, query_ids=ids)
# axis-parallel / zero-component rays through octreeRaySkip (exercises the 1e-10 reciprocal clamp)
eo = np.array([[0.0, 0.0, 2.0], [0.1, 0.05, -2.0], [2.0, 0.0, 0.0], [0.0, 2.0, 0.01], [1.5, 1.5, 1.5], [0.0, 0.0, 0.0]], np.float32)
ed = np.array([[0, 0, -1], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [-1, -1, -1], [1e-12, 1, 0]], np.float32)
et, eid = oc.rayskip(eo, ed)
store.update(edge_o=eo, edge_d=ed, edge_t=et, edge_id=eid)
np.savez_compressed(os.path.join(OUT, "golden_sphere32.npz"), **store)
meta["sphere32_cams"] = cams

# ---- sphere 128 and DT: checksums of structures + banded renders ------------------------------------------
dims, gmin, vox, data = B.sphere_grid(128)
oc128 = R.octree(dims, gmin, vox, data); n = oc128.build()
m128 = oc128.mesh(); m128.build()
bb, bm = m128.export()
meta["sphere128"] = dict(nodes=int(n), flat_sha=sha(oc128.flat()), tris=int(m128.n), tris_sha=sha(m128.tris()),
                         bvh_nodes=int(len(bb)), bvh_boxes_sha=sha(bb), bvh_meta_sha=sha(bm))

dt = R.octree(path=SRC); n = dt.build()
mdt = dt.mesh(); mdt.build()
bb, bm = mdt.export()
meta["dt"] = dict(dims=list(dt.dims), gmin=[float(x) for x in dt.gmin], voxel=dt.voxel, nodes=int(n), flat_sha=sha(dt.flat()),
                  tris=int(mdt.n), tris_sha=sha(mdt.tris()), bvh_nodes=int(len(bb)), bvh_boxes_sha=sha(bb), bvh_meta_sha=sha(bm))
store = {}
dcams = {}
bands = [(200, 204), (536, 544), (900, 904)]
for name, (th, ph, r) in {"far": (35, 40, 0.6 * 4250), "near": (60, 10, 0.35 * 4250)}.items():
    cam, view = R.camera(th, ph, r, width=1920, height=1080)
    dcams[name] = cam_dict(cam)
    for bi, (y0, y1) in enumerate(bands):
        for mode in (0, 1):
            o = dt.render(cam, mode, y0, y1, stats=True)
            store["oct%d_%s_%d_id" % (mode, name, bi)] = o["id"]
            store["oct%d_%s_%d_t" % (mode, name, bi)] = o["t"]
            store["oct%d_%s_%d_rgba" % (mode, name, bi)] = o["rgba"]
            store["oct%d_%s_%d_stats" % (mode, name, bi)] = o["stats"]
        o = mdt.render(cam, 1, 1e-3 * dt.voxel, y0, y1, stats=True)
        store["bvh1_%s_%d_id" % (name, bi)] = o["id"]
        store["bvh1_%s_%d_t" % (name, bi)] = o["t"]
        store["bvh1_%s_%d_rgba" % (name, bi)] = o["rgba"]
        store["bvh1_%s_%d_stats" % (name, bi)] = o["stats"]
store["bands"] = np.array(bands, np.int32)
np.savez_compressed(os.path.join(OUT, "golden_dt_rows.npz"), **store)
meta["dt_cams"] = dcams
json.dump(meta, open(os.path.join(OUT, "golden_meta.json"), "w"), indent=1, sort_keys=True)
print("golden fixtures written:", {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})
