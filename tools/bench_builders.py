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
        # mesh scene: host route (reference-shaped BVH + SAH topology, exact ids) vs device route (linear BVH)
        import torch
        W, H = 1920, 1080
        ext = float(max(g.dims) * g.voxel_size)
        cams = [rto.Camera.from_degrees(35, 40.0 + 45.0 * k, 0.6 * ext if name == "dt" else 0.9 * ext).consts(45.0, float(np.float32(W) / np.float32(H)), W, H)[0] for k in range(4)]
        bias = 1e-3 * g.voxel_size
        t0 = time.perf_counter(); hb = rto.HostBVH(tris_h); r["bvh_host_build_s"] = time.perf_counter() - t0
        t0 = time.perf_counter(); sh = rto.Scene.bvh(None, prebuilt=hb); r["bvh_host_layout_upload_s"] = time.perf_counter() - t0
        td2, sd = t(lambda: rto.Scene.bvh_device(tris_h), 2); r["bvh_device_from_host_tris_s"] = td2
        tg, sg = t(lambda: rto.Scene.bvh_from_grid(g), 2); r["mesh_scene_from_grid_on_device_s"] = tg
        r["mesh_scene_host_route_total_s"] = r["octree_host_s"] + r["mc_host_s"] + r["bvh_host_build_s"] + r["bvh_host_layout_upload_s"]
        F = len(cams)
        rgba = torch.empty((F, H, W, 4), dtype=torch.float32, device="cuda"); hid = torch.empty((F, H, W), dtype=torch.int32, device="cuda"); tt = torch.empty((F, H, W), dtype=torch.float32, device="cuda")
        ids = {}
        for label, sc in (("host", sh), ("device", sg)):
            ms = []
            for _ in range(4):
                sc.render_device(cams, rto.MODE_BVH, rto.FLAG_SHADOWS, bias, 0, H, rgba.data_ptr(), hid.data_ptr(), tt.data_ptr()); ms.append(sc.last_kernel_ms())
            ids[label] = hid.clone(); ids[label + "_t"] = tt.clone()
            hits = int((hid >= 0).sum().item())
            r["render_%s_tree_ms" % label] = float(np.median(ms[1:])); r["render_%s_tree_Mrays_s" % label] = (F * W * H + hits) / float(np.median(ms[1:])) / 1e3
        diff = ids["host"] != ids["device"]
        r["pixels"] = F * W * H; r["hit_ids_differing"] = int(diff.sum().item())
        if r["hit_ids_differing"]:
            ta, tb = ids["host_t"][diff].double(), ids["device_t"][diff].double()
            r["max_rel_t_gap_where_ids_differ"] = float(((ta - tb).abs() / ta.clamp_min(1e-30)).max().item())
            r["differing_with_device_tree_closer"] = int((tb < ta).sum().item()); r["differing_with_host_tree_closer"] = int((ta < tb).sum().item())
            r["differing_with_equal_t"] = int((ta == tb).sum().item())
        del sh, sd, sg, hb
        rep[name] = r
        print(name, json.dumps(r), flush=True)
    if a.out:
        json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
