#!/usr/bin/env python
"""Golden fixtures of the BASELINE.json configurations C3, C4 and C5 AT THEIR STATED SIZES, from the COMPILED REFERENCE
(oracle/_ref/libref.so: the reference's own translation units + the SURVEY.md 8c rule written with glm).

Run in the build container only (needs /root/reference; about 25 GB of RAM and 15 minutes on 8 cores for C4):
    python tests/golden/make_golden_fullsize.py [c3] [c5] [c4]
Output: tests/golden/golden_fullsize.json -- per case the camera constants, structure checksums and, for every frame, the
SHA-256 of the id / t / rgba planes of each of 18 horizontal bands (so a mismatch is localised) plus hit counts and id sums.

  c3  rto.city_block_grid(512, 1234, 32) -> createOctreeFromVoxelGrid (OctreeVoxel.cpp:765-778) -> both octree traversals at
      1920x1080: octreeRaySkip (VolumeRaycastRenderer.cpp:50-155) and the GLSL restatement (RayTracerBVH.cpp:239-327).
  c5  the DT mesh (sceneCache.bin -> octree -> MarchingCubesRenderer::render) through BVH::BVH / BVH::query + Moller-Trumbore +
      shadow rays at 3840x2160 for 4 of the 64 orbit cameras (k = 0, 16, 37, 53; phi_k = 360 k / 64, theta 35, r = 0.6 * 4250).
  c4  rto.city_block_grid(1024, 4321, 64) -> Adaptive Dual Contouring soup -> BVH::BVH (the reference's own build, BVH.cpp:19-71)
      -> BVH::query + Moller-Trumbore + shadow at 3840x2160, 2 cameras.  The reference's own mesher needs hours at this size
      (45 s for the 335 k triangles of the DT grid, most of it in its mutex-guarded edge cache), so the soup comes from
      rto_host_dc_mesh, which tests/test_dc_mesh.py pins bit for bit to AdaptiveDualContouringRenderer.cpp compiled in place on 15
      grids, and is cross-checked here at full size against the oracle port's literal sequential restatement of that mesher.
"""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bind as B
import ray_tracing_octrees_b200 as rto

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_fullsize.json")
NBANDS = 18
R = B.ref()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cam_dict(cam):
    return dict(camPos=[float(x) for x in cam.camPos], invView=[float(x) for x in cam.invView],
                tanHalfFov=float(cam.tanHalfFov), aspect=float(cam.aspect), width=cam.width, height=cam.height)


def frame_record(out, W, H):
    """Per-band hashes of the three planes of one full frame (row 0 = top)."""
    ids, t, rgba = out["id"].reshape(H, W), out["t"].reshape(H, W), out["rgba"].reshape(H, W, 4)
    bh = H // NBANDS
    bands = []
    for b in range(NBANDS):
        y0, y1 = b * bh, (H if b == NBANDS - 1 else (b + 1) * bh)
        bands.append(dict(y0=y0, y1=y1, id=sha(ids[y0:y1]), t=sha(t[y0:y1]), rgba=sha(rgba[y0:y1]), hits=int((ids[y0:y1] >= 0).sum())))
    hit = ids >= 0
    return dict(bands=bands, hits=int(hit.sum()), id_sum=int(ids[hit].astype(np.int64).sum()), shadowed=int((hit & (rgba[..., 0] == np.float32(0.1))).sum()),
                seconds=float(out["sec"]))


def make_c3():
    g = rto.city_block_grid(512, 1234, 32)
    t0 = time.time()
    oc = R.octree(g.dims, g.min, g.voxel_size, g.data)
    n = oc.build()
    rec = dict(grid=dict(dim=512, seed=1234, blocks=32, voxels_sha=sha(g.data), filled=int(g.data.sum())), nodes=int(n), flat_sha=sha(oc.flat()), octree_build_s=time.time() - t0, frames=[])
    W, H = 1920, 1080
    for theta, phi in ((35.0, 40.0), (35.0, 220.0)):
        cam, _ = R.camera(theta, phi, 0.9 * 512, width=W, height=H)
        fr = dict(theta=theta, phi=phi, radius=0.9 * 512, cam=cam_dict(cam))
        for mode, name in ((0, "modeA"), (1, "modeB")):
            fr[name] = frame_record(oc.render(cam, mode), W, H)
            print("c3", theta, phi, name, fr[name]["hits"], "%.1f s" % fr[name]["seconds"], flush=True)
        rec["frames"].append(fr)
    oc.free()
    return rec


def make_c5():
    dims, gmin, vox, data = B.load_scene_cache(os.path.join(ROOT, "tests", "golden", "dt_sceneCache.bin.gz"))
    oc = R.octree(dims, gmin, vox, data)
    oc.build()
    mesh = oc.mesh()
    mesh.build()
    W, H = 3840, 2160
    bias = float(np.float32(1e-3) * np.float32(vox))
    rec = dict(tris=int(mesh.n), tris_sha=sha(mesh.tris()), shadow_bias=bias, frames=[])
    for k in (0, 16, 37, 53):
        theta, phi = 35.0, 360.0 * k / 64
        cam, _ = R.camera(theta, phi, 0.6 * 4250, width=W, height=H)
        fr = dict(k=k, theta=theta, phi=phi, radius=0.6 * 4250, cam=cam_dict(cam), bvh_shadow=frame_record(mesh.render(cam, 1, bias), W, H))
        print("c5 k", k, fr["bvh_shadow"]["hits"], "%.1f s" % fr["bvh_shadow"]["seconds"], flush=True)
        rec["frames"].append(fr)
    mesh.free(); oc.free()
    return rec


def make_c4():
    g = rto.city_block_grid(1024, 4321, 64)
    t0 = time.time()
    nodes = rto.create_octree_from_voxel_grid(g)           # host builder, pinned to createOctreeFromVoxelGrid + setOctree in tests/test_host_builders.py
    t1 = time.time()
    tris = rto.dual_contouring_mesh(g, nodes)
    t2 = time.time()
    print("c4 octree %d nodes %.1f s, DC mesh %d tris %.1f s" % (len(nodes), t1 - t0, len(tris), t2 - t1), flush=True)
    rec = dict(grid=dict(dim=1024, seed=4321, blocks=64, voxels_sha=sha(g.data), filled=int(g.data.sum())), nodes=int(len(nodes)), flat_sha=sha(nodes),
               tris=int(len(tris)), tris_sha=sha(tris), frames=[])
    if os.environ.get("RTO_GOLDEN_SKIP_PORT_DC", "0") != "1":
        # cross-check of the soup at full size: the port's literal, sequential restatement of the reference's mesher (two maps, visit order)
        P = B.port()
        poc = P.octree(g.dims, g.min, g.voxel_size, g.data)
        poc.build()
        pm = poc.dc_mesh()
        same = pm.n == len(tris) and sha(pm.tris()) == rec["tris_sha"]
        print("c4 port DC mesh: %d tris, identical=%s, %.1f s" % (pm.n, same, time.time() - t2), flush=True)
        assert same, "rto_host_dc_mesh differs from the oracle port's sequential restatement at 1024^3"
        rec["dc_soup_equals_port_restatement"] = True
        pm.free(); poc.free()
    del nodes
    t3 = time.time()
    mesh = R.mesh(tris)
    sec = mesh.build()                                     # the reference's own BVH::BVH
    print("c4 reference BVH::BVH %.1f s" % sec, flush=True)
    rec["ref_bvh_build_s"] = float(sec)
    W, H = 3840, 2160
    bias = float(np.float32(1e-3))
    rec["shadow_bias"] = bias
    for theta, phi in ((35.0, 40.0), (35.0, 130.0)):
        cam, _ = R.camera(theta, phi, 0.9 * 1024, width=W, height=H)
        fr = dict(theta=theta, phi=phi, radius=0.9 * 1024, cam=cam_dict(cam), bvh_shadow=frame_record(mesh.render(cam, 1, bias), W, H))
        print("c4", theta, phi, fr["bvh_shadow"]["hits"], "%.1f s" % fr["bvh_shadow"]["seconds"], flush=True)
        rec["frames"].append(fr)
    mesh.free()
    return rec


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a in ("c3", "c4", "c5")] or ["c3", "c5", "c4"]
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    data["bands_per_frame"] = NBANDS
    data["made_by"] = "tests/golden/make_golden_fullsize.py with oracle/_ref/libref.so (the reference's translation units compiled in place)"
    for w in which:
        data[w] = dict(c3=make_c3, c4=make_c4, c5=make_c5)[w]()
        json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print("written", OUT, os.path.getsize(OUT), "bytes")
