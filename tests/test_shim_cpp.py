"""The C++ shim headers (reference class names on top of the C ABI) compile with a plain host compiler and run: CPU part
here (host builders through the shim classes), full frames on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "ray_tracing_octrees_b200", "csrc", "shim")


def _build(tmp_path):
    exe = str(tmp_path / "example_headless")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.join(ROOT, "ray_tracing_octrees_b200")
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-O1", os.path.join(SHIM, "example_headless.cpp"), "-L" + libdir, "-lrto",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_shim_compiles_and_host_part_runs(rto, tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "octree nodes 6025, triangles 7768, bvh nodes 8191" in out.stdout      # same counts as the oracle on sphere-32
    assert "dual contouring triangles 4113" in out.stdout                         # tests/golden/golden_dc.json
    # the reference's host-side types of the boundary: BVHNode pointer tree (every triangle in exactly one leaf), Frustum::testAABB
    # (the unit box straddles the frustum, a box behind the camera is outside), OctreeNode without extra members + index side table
    assert "bvh pointer tree: 4096 leaves holding 7768 triangles" in out.stdout
    assert "frustum: unit box 0, box behind the camera -1, flat index of the root 0, of its child 3 4" in out.stdout


def test_shim_octree_node_has_the_reference_layout(rto, tmp_path):
    """OctreeNode (OctreeVoxel.h:45-62): 4 ints, 3 bools, parent and 8 child pointers -- nothing else (LP64: 96 bytes)."""
    src = tmp_path / "layout.cpp"
    src.write_text('#include "OctreeVoxel.h"\n#include "BVH.h"\n#include <cstddef>\n'
                   'static_assert(sizeof(OctreeNode) == 16 + 8 + 8 + 64, "OctreeNode carries members the reference does not have");\n'
                   'static_assert(offsetof(OctreeNode, parent) == 24 && offsetof(OctreeNode, children) == 32, "OctreeNode layout");\n'
                   'static_assert(sizeof(AABB) == 24 && sizeof(BVHNode) == 24 + 8 + 8 + sizeof(std::vector<const Triangle*>), "AABB / BVHNode layout");\n'
                   'int main() { return 0; }\n')
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-Wno-invalid-offsetof", "-I" + SHIM, "-fsyntax-only", str(src)])


def test_shim_frame_planes_without_a_device(rto, tmp_path):
    """Framebuffer's planes come from rto_host_alloc_pinned; on a machine without a CUDA device that call says RTO_ERR_NO_DEVICE and the
    allocator hands out ordinary memory instead (host-only use of the builders): planes can be made, grown, copied and freed."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a CUDA device")
    src = tmp_path / "planes.cpp"
    src.write_text('#include "RayTracerBVH.h"\n#include <cstdio>\n'
                   'int main() {\n'
                   '  void* p = nullptr; int rc = rto_host_alloc_pinned(64, &p);\n'
                   '  std::printf("pinned rc %d ptr %d\\n", rc, p != nullptr);\n'
                   '  Framebuffer fb; fb.rgba.resize(4 * 640 * 480, 0.5f); fb.hitId.resize(640 * 480, -1); fb.t.resize(640 * 480, 1e30f);\n'
                   '  fb.rgba.resize(4 * 1920 * 1080, 0.25f);\n'
                   '  Framebuffer copy = fb; double s = 0; for (float v : copy.rgba) s += v;\n'
                   '  std::printf("sum %.1f ids %zu\\n", s, copy.hitId.size());\n'
                   '  rto_host_free_pinned(nullptr);\n'
                   '  return 0; }\n')
    exe = str(tmp_path / "planes")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.join(ROOT, "ray_tracing_octrees_b200")
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-O1", "-I" + SHIM, "-I" + os.path.join(ROOT, "include"), str(src), "-L" + libdir, "-lrto", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "pinned rc 2 ptr 0" in out.stdout                                      # RTO_ERR_NO_DEVICE, nothing allocated
    want = 0.5 * 4 * 640 * 480 + 0.25 * (4 * 1920 * 1080 - 4 * 640 * 480)
    assert "sum %.1f ids 307200" % want in out.stdout


@pytest.mark.gpu
def test_shim_renders_on_gpu(rto, tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "octree frame:" in out.stdout and "mesh frame:" in out.stdout and "BVH::query candidates:" in out.stdout
    # the sphere scene spans one unit: with the reference's culling margin of 150 every node stays, so the culled frame equals the plain one
    import re
    plain = int(re.search(r"octree frame: (\d+) of", out.stdout).group(1))
    m = re.search(r"culled frame: (\d+) of (\d+) nodes visible, (\d+) pixels hit", out.stdout)
    assert m and m.group(1) == m.group(2) == "6025" and int(m.group(3)) == plain
    # the frames the C++ program got through the shim classes == the frames of the Python binding (which the parity tests pin to the
    # oracle), plane by plane: FNV-1a over hit ids, t and rgba
    import numpy as np
    assert rto.lib().rto_init(0) == 0

    def fnv(*arrays):
        h = 1469598103934665603
        for a in arrays:
            for b in np.ascontiguousarray(a).tobytes():
                h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return "%016x" % h
    grid = rto.generate_test_volume(32)
    nodes = rto.create_octree_from_voxel_grid(grid)
    f32 = lambda v: float(np.float32(v))
    cam, _ = rto.Camera(f32(0.5235988), f32(0.6981317), f32(1.2)).consts(45.0, f32(np.float32(160.0) / np.float32(120.0)), 160, 120)
    oc = rto.Scene.octree(nodes, grid.min, grid.voxel_size).render(cam, rto.MODE_OCTREE_GLSL)
    assert "octree frame hash " + fnv(oc["id"], oc["t"], oc["rgba"]) in out.stdout
    assert "scene cache: frame hash " + fnv(oc["id"], oc["t"], oc["rgba"]) in out.stdout      # saveSceneCache -> loadSceneCache -> same frame
    tris = rto.marching_cubes_mesh(grid, nodes)
    sc = rto.Scene.bvh(tris)
    bias = f32(np.float32(1e-3) * np.float32(grid.voxel_size))
    me = sc.render(cam, rto.MODE_BVH, rto.FLAG_SHADOWS, bias)
    assert "mesh frame hash " + fnv(me["id"], me["t"], me["rgba"]) in out.stdout
    assert "batch of 3 frames: first frame hash " + fnv(me["id"], me["t"], me["rgba"]) in out.stdout and "identical" in out.stdout
    # the single-ray octreeRaySkip wrapper == the batched call
    ro = np.array(cam.camPos, np.float32)
    nrd = -ro
    ln = np.sqrt(np.float32(np.float32(nrd[0] * nrd[0] + nrd[1] * nrd[1]) + nrd[2] * nrd[2]))
    rd = (nrd / ln).astype(np.float32)
    t, _ = rto.Scene.octree(nodes, grid.min, grid.voxel_size).trace_rays(ro[None, :], rd[None, :], rto.MODE_OCTREE_SKIP)
    m = re.search(r"octreeRaySkip towards the centre: ([0-9.e+-]+)", out.stdout)
    assert m and m.group(1) == "%.9g" % t[0] and t[0] < 1e29
