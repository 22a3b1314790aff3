"""Python mirror of the reference's interface for the ray-casting hot path, on top of the C ABI (librto.so).

Names follow the reference (abodthedude25/Ray_Tracing_Octrees):
  VoxelGrid                         OctreeVoxel.h:28-42
  create_octree_from_voxel_grid     createOctreeFromVoxelGrid (OctreeVoxel.cpp:765-778) + RayTracerBVH::setOctree's
                                    BFS numbering (RayTracerBVH.cpp:443-490) -> (n, 15) int32 GPUNodes array
  marching_cubes_mesh               MarchingCubesRenderer::render(root, grid, 0,0,0, root->size) (Renderer.cpp:14-36)
  BVH                               BVH.h:44-63 (build on the host, query on the GPU)
  Camera                            Camera.h:5-44 (orbit camera; angles in radians)
  RayTracerBVH                      RayTracerBVH.h:28-80: set_octree / render_scene_compute, plus the mesh path
                                    (set_mesh) that BASELINE.json's north_star adds.
Everything that traces a ray runs in CUDA; nothing here computes a pixel on the CPU.
"""
import ctypes as C
import gzip
import os

import numpy as np

from . import _lib as L
from ._lib import (FLAG_NO_PRUNE, FLAG_SHADOWS, FLAG_SORT_RAYS, MEM_DEVICE, MEM_HOST, MODE_BVH, MODE_OCTREE_GLSL, MODE_OCTREE_SKIP,
                   RtoCamera, RtoError, RtoFrame, check, lib)

MISS_T = np.float32(1e30)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _take(ptr, count, dtype, shape):
    """Copy a malloc'ed C array into numpy and free it."""
    if not ptr or count == 0:
        return np.zeros((0,) + tuple(shape[1:]), dtype)
    n = int(np.prod(shape))
    buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy().reshape(shape)
    lib().rto_host_free(ptr)
    return buf


class VoxelGrid:
    """VoxelGrid (OctreeVoxel.h:28-42): x-fastest uint8 occupancy, world min corner, voxel size."""

    def __init__(self, dims, gmin, voxel_size, data):
        self.dims = tuple(int(d) for d in dims)
        self.min = np.asarray(gmin, np.float32).copy()
        self.voxel_size = float(np.float32(voxel_size))
        self.data = np.ascontiguousarray(data, np.uint8).ravel()
        if self.data.size != self.dims[0] * self.dims[1] * self.dims[2]:
            raise ValueError("voxel data size does not match dims")

    @staticmethod
    def load(path):
        """loadVoxelGrid (CacheUtils.cpp:33-59); also accepts a gzip-compressed sceneCache.bin."""
        if path.endswith(".gz"):
            raw = gzip.open(path, "rb").read()
            dims = np.frombuffer(raw, np.int32, 3, 0)
            mv = np.frombuffer(raw, np.float32, 4, 12)
            n = int(np.frombuffer(raw, np.uint64, 1, 28)[0])
            data = np.frombuffer(raw, np.uint8, n, 36)
            if n != int(dims[0]) * int(dims[1]) * int(dims[2]):
                raise IOError("inconsistent voxel cache " + path)
            return VoxelGrid(dims, mv[:3], mv[3], data)
        dims = np.zeros(3, np.int32)
        mv = np.zeros(4, np.float32)
        ptr = C.c_void_p()
        check(lib().rto_host_grid_load(path.encode(), _p(dims), _p(mv), C.byref(ptr)))
        n = int(dims[0]) * int(dims[1]) * int(dims[2])
        data = _take(ptr, n, np.uint8, (n,))
        return VoxelGrid(dims, mv[:3], mv[3], data)

    def save(self, path):
        dims = np.asarray(self.dims, np.int32)
        mv = np.concatenate([self.min, [np.float32(self.voxel_size)]]).astype(np.float32)
        check(lib().rto_host_grid_save(path.encode(), _p(dims), _p(mv), _p(self.data)))


def load_csv_data_into_voxel_grid(verts_csv, faces_csv, voxel_size=5.0, device=False):
    """loadCSVDataIntoVoxelGrid (BuildingLoader.cpp:153-290): DT vertex/face CSV files -> VoxelGrid (None when the input is empty,
    where the reference returns an empty grid).  device=True rasterises the faces on the GPU (same grid)."""
    dims = np.zeros(3, np.int32)
    mv = np.zeros(4, np.float32)
    ptr = C.c_void_p()
    fn = lib().rto_device_csv_voxelize if device else lib().rto_host_csv_voxelize
    check(fn(os.fsencode(verts_csv), os.fsencode(faces_csv), float(voxel_size), _p(dims), _p(mv), C.byref(ptr)))
    n = int(dims[0]) * int(dims[1]) * int(dims[2])
    if n == 0:
        return None
    return VoxelGrid(dims, mv[:3], mv[3], _take(ptr, n, np.uint8, (n,)))


def create_octree_from_voxel_grid(grid):
    """-> (n, 15) int32 array of GPUNodes in the reference's BFS numbering (row index == node / leaf id)."""
    ptr = C.c_void_p()
    n = C.c_size_t()
    check(lib().rto_host_octree_build(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], C.byref(ptr), C.byref(n)))
    return _take(ptr, n.value, np.int32, (n.value, 15))


def view_proj(view16, fov_deg, aspect, z_near=0.01, z_far=5000.0):
    """proj * view of renderSceneComputeWithCulling (RayTracerBVH.cpp:733-734) as 16 floats, column-major."""
    v = np.ascontiguousarray(view16, np.float32).ravel()
    out = np.zeros(16, np.float32)
    check(lib().rto_host_view_proj(_p(v), float(fov_deg), float(aspect), float(z_near), float(z_far), _p(out)))
    return out


def frustum_cull(nodes, grid, view_proj16, margin=150.0, device=False):
    """The node array renderSceneComputeWithCulling uploads (RayTracerBVH.cpp:724-813) -> (culled (m, 15) int32, new_to_old (m,) int32)."""
    nodes = np.ascontiguousarray(nodes, np.int32).reshape(-1, 15)
    vp = np.ascontiguousarray(view_proj16, np.float32).ravel()
    ptr, back = C.c_void_p(), C.c_void_p()
    n = C.c_size_t()
    fn = lib().rto_device_frustum_cull if device else lib().rto_host_frustum_cull
    check(fn(_p(nodes), len(nodes), _p(grid.min), float(grid.voxel_size), _p(vp), float(margin), C.byref(ptr), C.byref(n), C.byref(back)))
    return _take(ptr, n.value, np.int32, (n.value, 15)), _take(back, n.value, np.int32, (n.value,))


def skip_probe_rays(view16, cam_pos, aspect):
    """The 49 probe rays of VolumeRaycastRenderer.cpp:1618-1629 -> (origins (49, 3), dirs (49, 3))."""
    v = np.ascontiguousarray(view16, np.float32).ravel()
    p = np.ascontiguousarray(cam_pos, np.float32).ravel()
    o, d = np.zeros((49, 3), np.float32), np.zeros((49, 3), np.float32)
    check(lib().rto_host_skip_probe_rays(_p(v), _p(p), float(aspect), _p(o), _p(d)))
    return o, d


def skip_distance_from_probes(t, last=0.0):
    t = np.ascontiguousarray(t, np.float32).ravel()
    return float(lib().rto_host_skip_distance_from_probes(_p(t), len(t), float(last)))


def create_octree_on_device(grid):
    """Same array as create_octree_from_voxel_grid, built on the GPU (rto_device_octree_build)."""
    ptr = C.c_void_p()
    n = C.c_size_t()
    check(lib().rto_device_octree_build(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], C.byref(ptr), C.byref(n)))
    return _take(ptr, n.value, np.int32, (n.value, 15))


def marching_cubes_mesh_on_device(grid):
    """Same triangle soup as marching_cubes_mesh, extracted on the GPU (rto_device_mc_mesh)."""
    ptr = C.c_void_p()
    n = C.c_size_t()
    check(lib().rto_device_mc_mesh(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], _p(grid.min), grid.voxel_size,
                                   C.byref(ptr), C.byref(n)))
    return _take(ptr, n.value, np.float32, (n.value, 9))


def marching_cubes_mesh(grid, nodes):
    """-> (m, 9) float32 triangle soup (v0, v1, v2) in the reference's emission order."""
    nodes = np.ascontiguousarray(nodes, np.int32)
    ptr = C.c_void_p()
    n = C.c_size_t()
    check(lib().rto_host_mc_mesh(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], _p(grid.min), grid.voxel_size,
                                 _p(nodes), len(nodes), C.byref(ptr), C.byref(n)))
    return _take(ptr, n.value, np.float32, (n.value, 9))


def dual_contouring_mesh(grid, nodes, view_proj=None, margin=50.0, algo="default"):
    """-> (m, 9) float32: the Adaptive Dual Contouring triangle soup in the reference's emission order (renderOctree over
    AdaptiveDualContouringRenderer::createTriangles, main.cpp:95-208); view_proj (16 floats, see view_proj()) culls like renderOctree."""
    nodes = np.ascontiguousarray(nodes, np.int32)
    vp = None if view_proj is None else np.ascontiguousarray(view_proj, np.float32).ravel()
    ptr = C.c_void_p()
    n = C.c_size_t()
    fn = {"default": lib().rto_host_dc_mesh, "replay": lib().rto_host_dc_mesh_replay, "device": lib().rto_device_dc_mesh}[algo]
    check(fn(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], _p(grid.min), grid.voxel_size,
             _p(nodes), len(nodes), _p(vp), float(margin), C.byref(ptr), C.byref(n)))
    return _take(ptr, n.value, np.float32, (n.value, 9))


def dual_contouring_mesh_with_normals(grid, nodes, view_proj=None, margin=50.0):
    """-> ((m, 9) triangles, (m, 3) normals): rto_host_dc_mesh_normals, the MCTriangles of the reference's mesher."""
    nodes = np.ascontiguousarray(nodes, np.int32)
    vp = None if view_proj is None else np.ascontiguousarray(view_proj, np.float32).ravel()
    ptr, nptr = C.c_void_p(), C.c_void_p()
    n = C.c_size_t()
    check(lib().rto_host_dc_mesh_normals(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], _p(grid.min), grid.voxel_size,
                                         _p(nodes), len(nodes), _p(vp), float(margin), C.byref(ptr), C.byref(nptr), C.byref(n)))
    return _take(ptr, n.value, np.float32, (n.value, 9)), _take(nptr, n.value, np.float32, (n.value, 3))


def save_triangle_cache(path, tris, normals=None):
    """saveTriangleCache (main.cpp:27-45): size_t count + count x MCTriangle (72 bytes)."""
    tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
    nm = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    check(lib().rto_host_tricache_save(os.fsencode(path), _p(tris), _p(nm), len(tris)))


def load_triangle_cache(path):
    """loadTriangleCache (main.cpp:48-67) -> ((m, 9) triangles, (m, 9) stored normals)."""
    ptr, nptr = C.c_void_p(), C.c_void_p()
    n = C.c_size_t()
    check(lib().rto_host_tricache_load(os.fsencode(path), C.byref(ptr), C.byref(nptr), C.byref(n)))
    return _take(ptr, n.value, np.float32, (n.value, 9)), _take(nptr, n.value, np.float32, (n.value, 9))


class Camera:
    """Orbit camera (Camera.h:5-44): theta / phi in radians, radius, target."""

    def __init__(self, theta, phi, radius, target=(0.0, 0.0, 0.0)):
        self.theta, self.phi, self.radius = float(theta), float(phi), float(radius)
        self.target = np.asarray(target, np.float32).copy()

    @staticmethod
    def from_degrees(theta_deg, phi_deg, radius, target=(0.0, 0.0, 0.0)):
        return Camera(float(np.deg2rad(np.float32(theta_deg))), float(np.deg2rad(np.float32(phi_deg))), radius, target)

    def consts(self, fov_deg, aspect, width, height):
        """getView/getPos + the per-frame constants of generateRay -> (RtoCamera, view matrix as 16 floats)."""
        cam = RtoCamera()
        view = np.zeros(16, np.float32)
        check(lib().rto_host_camera_orbit(self.theta, self.phi, self.radius, _p(self.target), fov_deg, aspect,
                                          width, height, C.byref(cam), _p(view)))
        return cam, view


class HostBVH:
    """BVH::BVH(triangles) on the host (BVH.cpp:19-71).  Keeps the triangle array alive like the caller must."""

    def __init__(self, tris):
        self.tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        self.h = C.c_void_p()
        check(lib().rto_host_bvh_build(_p(self.tris), len(self.tris), C.byref(self.h)))

    @property
    def num_nodes(self):
        return lib().rto_host_bvh_num_nodes(self.h)

    def export(self):
        n = self.num_nodes
        boxes = np.zeros((n, 6), np.float32)
        meta = np.zeros((n, 4), np.int32)
        check(lib().rto_host_bvh_export(self.h, _p(boxes), _p(meta), n))
        return boxes, meta

    def close(self):
        if self.h:
            lib().rto_host_bvh_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """A device-resident scene (RtoScene*)."""

    def __init__(self, handle, keep=()):
        self.h = handle
        self._keep = keep

    @staticmethod
    def octree(nodes, grid_min, voxel_size):
        nodes = np.ascontiguousarray(nodes, np.int32).reshape(-1, 15)
        gm = np.asarray(grid_min, np.float32)
        h = C.c_void_p()
        check(lib().rto_scene_create_octree(_p(nodes), len(nodes), _p(gm), float(voxel_size), C.byref(h)))
        return Scene(h)

    @staticmethod
    def octree_from_grid(grid):
        """Voxel grid -> octree scene, built and linearised on the GPU (rto_scene_create_octree_from_grid)."""
        h = C.c_void_p()
        check(lib().rto_scene_create_octree_from_grid(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], _p(grid.min),
                                                      float(grid.voxel_size), C.byref(h)))
        return Scene(h)

    def octree_layout(self):
        """Diagnostic: (desc, up, inner) arrays of a compact octree scene as numpy arrays."""
        n = self.info()["nodes"]
        desc = np.zeros(n + 8, np.uint32)
        up = np.zeros((n + 7) // 8 + 1, np.int32)
        ninner = max((n - 1) // 8, 1)
        inner = np.zeros((ninner, 4), np.int32)
        k = C.c_size_t()
        check(lib().rto_scene_octree_layout_read(self.h, _p(desc), _p(up), _p(inner), C.byref(k)))
        return desc, up, inner

    @staticmethod
    def bvh(tris, prebuilt=None):
        tris = prebuilt.tris if prebuilt is not None else np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        h = C.c_void_p()
        check(lib().rto_scene_create_bvh(_p(tris), len(tris), prebuilt.h if prebuilt is not None else None, C.byref(h)))
        return Scene(h, keep=(tris,))

    @staticmethod
    def bvh_device(tris):
        """Triangles -> linear BVH built on the GPU (rto_scene_create_bvh_device): the fast, not reference-shaped route."""
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        h = C.c_void_p()
        check(lib().rto_scene_create_bvh_device(_p(tris), len(tris), C.byref(h)))
        return Scene(h)

    @staticmethod
    def bvh_from_grid(grid):
        """Voxel grid -> octree -> MC mesh -> linear BVH, all on the GPU (rto_scene_create_bvh_from_grid)."""
        h = C.c_void_p()
        check(lib().rto_scene_create_bvh_from_grid(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], _p(grid.min),
                                                   float(grid.voxel_size), C.byref(h)))
        return Scene(h)

    @staticmethod
    def bvh_from_grid_dc(grid, view_proj=None, margin=50.0):
        """Voxel grid -> octree -> Dual-Contouring mesh -> linear BVH, all on the GPU (rto_scene_create_bvh_from_grid_dc)."""
        vp = None if view_proj is None else np.ascontiguousarray(view_proj, np.float32).ravel()
        h = C.c_void_p()
        check(lib().rto_scene_create_bvh_from_grid_dc(_p(grid.data), grid.dims[0], grid.dims[1], grid.dims[2], _p(grid.min),
                                                      float(grid.voxel_size), _p(vp), float(margin), C.byref(h)))
        return Scene(h)

    def skip_distance(self, view16, cam_pos, aspect, last=0.0):
        """VolumeRaycastRenderer's skip-distance estimate (VolumeRaycastRenderer.cpp:1598-1664) -> (distance, 49 probe results)."""
        v = np.ascontiguousarray(view16, np.float32).ravel()
        p = np.ascontiguousarray(cam_pos, np.float32).ravel()
        out = C.c_float()
        probes = np.zeros(49, np.float32)
        check(lib().rto_octree_skip_distance(self.h, _p(v), _p(p), float(aspect), float(last), C.byref(out), _p(probes)))
        return out.value, probes

    def save(self, path):
        """The device layout of this scene as one file (rto_scene_save)."""
        check(lib().rto_scene_save(self.h, os.fsencode(path)))

    @staticmethod
    def load(path):
        """A scene from a file written by save(): one read, one upload, no builder runs (rto_scene_load)."""
        h = C.c_void_p()
        check(lib().rto_scene_load(os.fsencode(path), C.byref(h)))
        return Scene(h)

    def info(self):
        kind, compact = C.c_int(), C.c_int()
        prims, nodes, nbytes = C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(lib().rto_scene_info(self.h, C.byref(kind), C.byref(prims), C.byref(nodes), C.byref(nbytes), C.byref(compact)))
        return dict(kind=kind.value, prims=prims.value, nodes=nodes.value, device_bytes=nbytes.value, compact=compact.value)

    @property
    def stream(self):
        return lib().rto_scene_stream(self.h)

    def sync(self):
        check(lib().rto_scene_sync(self.h))

    def last_kernel_ms(self):
        ms = C.c_float()
        check(lib().rto_scene_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        return int(lib().rto_scene_launch_count(self.h))

    def render(self, cam, mode, flags=0, shadow_bias=0.0, y0=0, y1=None, want=("rgba", "id", "t")):
        """Host-memory render of rows [y0, y1): returns dict(rgba (n,4) f32, id (n,) i32, t (n,) f32)."""
        y1 = cam.height if y1 is None else y1
        n = max(0, y1 - y0) * max(0, cam.width)          # invalid ranges are reported by the library (RTO_ERR_INVALID)
        out = dict(rgba=np.empty((n, 4), np.float32) if "rgba" in want else None,
                   id=np.empty(n, np.int32) if "id" in want else None,
                   t=np.empty(n, np.float32) if "t" in want else None)
        fr = RtoFrame(_p(out["rgba"]), _p(out["id"]), _p(out["t"]), MEM_HOST)
        check(lib().rto_render(self.h, C.byref(cam), mode, flags, shadow_bias, y0, y1, C.byref(fr)))
        return out

    def render_device(self, cams, mode, flags=0, shadow_bias=0.0, y0=0, y1=None, rgba_ptr=None, id_ptr=None, t_ptr=None):
        """Enqueue a render of one camera or a list of cameras writing to device pointers (asynchronous)."""
        if isinstance(cams, RtoCamera):
            cams = [cams]
        arr = cams if isinstance(cams, C.Array) else (RtoCamera * len(cams))(*cams)
        y1 = arr[0].height if y1 is None else y1
        fr = RtoFrame(rgba_ptr, id_ptr, t_ptr, MEM_DEVICE)
        check(lib().rto_render_batch(self.h, arr, len(arr), mode, flags, shadow_bias, y0, y1, C.byref(fr)))

    def render_host_ptrs(self, cams, mode, flags, shadow_bias, y0, y1, rgba_ptr, id_ptr, t_ptr):
        """Synchronous render into caller-owned HOST pointers (e.g. pinned torch tensors)."""
        if isinstance(cams, RtoCamera):
            cams = [cams]
        arr = cams if isinstance(cams, C.Array) else (RtoCamera * len(cams))(*cams)
        fr = RtoFrame(rgba_ptr, id_ptr, t_ptr, MEM_HOST)
        check(lib().rto_render_batch(self.h, arr, len(arr), mode, flags, shadow_bias, y0, y1, C.byref(fr)))

    def render_codes(self, cams, flags, shadow_bias, codes_ptr, first_frame=0, y0=0, y1=None, stream=None):
        """rto_render_codes: trace rows [y0, y1) of the cameras and write one 32-bit hit code per pixel (tile order) into the code
        buffer at device address codes_ptr (local, peer or IPC-mapped memory).  Asynchronous."""
        if isinstance(cams, RtoCamera):
            cams = [cams]
        arr = cams if isinstance(cams, C.Array) else (RtoCamera * len(cams))(*cams)
        y1 = arr[0].height if y1 is None else y1
        check(lib().rto_render_codes(self.h, arr, len(arr), flags, shadow_bias, y0, y1, C.c_void_p(codes_ptr), first_frame, C.c_void_p(stream or 0)))

    def resolve_codes(self, cams, codes_ptr, first_frame=0, y0=0, y1=None, rgba_ptr=None, id_ptr=None, t_ptr=None, memory=MEM_DEVICE, stream=None):
        """rto_resolve_codes: rebuild the planes of rows [y0, y1) of the cameras from hit codes (bit-identical to a direct render)."""
        if isinstance(cams, RtoCamera):
            cams = [cams]
        arr = cams if isinstance(cams, C.Array) else (RtoCamera * len(cams))(*cams)
        y1 = arr[0].height if y1 is None else y1
        fr = RtoFrame(rgba_ptr, id_ptr, t_ptr, memory)
        check(lib().rto_resolve_codes(self.h, arr, len(arr), y0, y1, C.c_void_p(codes_ptr), first_frame, C.byref(fr), C.c_void_p(stream or 0)))

    def render_via_codes(self, cam, flags=0, shadow_bias=0.0):
        """One frame through the compact path (codes on the device, planes rebuilt from them) -> host arrays like render()."""
        words = codes_frame_words(cam.width, cam.height)
        buf = ExchangeBuffer(words * 4)
        try:
            self.render_codes(cam, flags, shadow_bias, buf.ptr)
            n = cam.width * cam.height
            out = dict(rgba=np.empty((n, 4), np.float32), id=np.empty(n, np.int32), t=np.empty(n, np.float32))
            self.resolve_codes(cam, buf.ptr, rgba_ptr=_p(out["rgba"]), id_ptr=_p(out["id"]), t_ptr=_p(out["t"]), memory=MEM_HOST)
        finally:
            self.sync()
            buf.close()
        return out

    def stats(self, cam, mode, flags=0, shadow_bias=0.0, y0=0, y1=None):
        y1 = cam.height if y1 is None else y1
        st = np.zeros(5, np.uint64)
        check(lib().rto_render_stats(self.h, C.byref(cam), mode, flags, shadow_bias, y0, y1, _p(st)))
        return st

    def trace_rays(self, origins, dirs, mode, flags=0, tmin=0.0, tmax=1e30):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        t = np.empty(len(o), np.float32)
        ids = np.empty(len(o), np.int32)
        check(lib().rto_trace_rays(self.h, mode, flags, _p(o), _p(d), len(o), tmin, tmax, _p(t), _p(ids), MEM_HOST))
        return t, ids

    def trace_rays_device(self, o_ptr, d_ptr, n, mode, flags, t_ptr, id_ptr, tmin=0.0, tmax=1e30):
        """rto_trace_rays on device-resident ray lists and outputs (raw device addresses); asynchronous on the scene's stream."""
        check(lib().rto_trace_rays(self.h, mode, flags, C.c_void_p(o_ptr), C.c_void_p(d_ptr), n, tmin, tmax, C.c_void_p(t_ptr), C.c_void_p(id_ptr), MEM_DEVICE))

    def query(self, origins, dirs):
        """BVH::query for many rays -> (offsets int64 (n+1,), triangle ids int32) in the reference's candidate order."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        off = np.zeros(len(o) + 1, np.int64)
        total = C.c_size_t()
        check(lib().rto_bvh_query(self.h, _p(o), _p(d), len(o), _p(off), None, 0, C.byref(total)))
        ids = np.zeros(total.value, np.int32)
        if total.value:
            check(lib().rto_bvh_query(self.h, _p(o), _p(d), len(o), _p(off), _p(ids), total.value, C.byref(total)))
        return off, ids

    def close(self):
        if self.h:
            lib().rto_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def codes_frame_words(width, height):
    """32-bit words of one frame's hit-code plane (tile order, padded to whole 16 x 8 blocks)."""
    return int(lib().rto_codes_frame_words(width, height))


class ExchangeBuffer:
    """Device memory other GPUs of the box write hit codes into (rto_exchange_*): allocate on the receiving device and hand
    `handle` (64 bytes) to the other processes, or map another process's buffer with ExchangeBuffer.open(handle)."""

    def __init__(self, nbytes=0, _mapped=None):
        self.ptr, self.handle, self.mapped = None, None, _mapped is not None
        if _mapped is not None:
            self.ptr = _mapped
            return
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        check(lib().rto_exchange_alloc(nbytes, C.byref(p), h))
        self.ptr, self.handle, self.nbytes = p.value, bytes(h), nbytes

    @staticmethod
    def open(handle):
        h = (C.c_ubyte * 64).from_buffer_copy(bytes(handle))
        p = C.c_void_p()
        check(lib().rto_exchange_open(h, C.byref(p)))
        return ExchangeBuffer(_mapped=p.value)

    def close(self):
        if self.ptr:
            check((lib().rto_exchange_close if self.mapped else lib().rto_exchange_free)(C.c_void_p(self.ptr)))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """Several GPUs of one box behind one handle (rto_group_*): scene replicated, rows of a batch dealt to the devices, hit codes
    written into the first device's memory over NVLink and expanded there."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*devices)
        self.h = C.c_void_p()
        check(lib().rto_group_create(devs, len(devices), C.byref(self.h)))
        self.devices = list(devices)
        self._keep = ()

    def set_mesh(self, tris, prebuilt=None):
        tris = prebuilt.tris if prebuilt is not None else np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        check(lib().rto_group_scene_bvh(self.h, _p(tris), len(tris), prebuilt.h if prebuilt is not None else None))
        self._keep = (tris, prebuilt)

    def set_balancing(self, enabled=True, weights=None, chunks=0):
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        check(lib().rto_group_set_balancing(self.h, int(enabled), _p(w), chunks))

    def render(self, cams, flags=0, shadow_bias=0.0):
        """Host-memory render of whole frames -> dict(rgba (F, n, 4), id (F, n), t (F, n))."""
        if isinstance(cams, RtoCamera):
            cams = [cams]
        arr = (RtoCamera * len(cams))(*cams)
        n = cams[0].width * cams[0].height
        out = dict(rgba=np.empty((len(cams), n, 4), np.float32), id=np.empty((len(cams), n), np.int32), t=np.empty((len(cams), n), np.float32))
        fr = RtoFrame(_p(out["rgba"]), _p(out["id"]), _p(out["t"]), MEM_HOST)
        check(lib().rto_group_render_batch(self.h, arr, len(cams), flags, shadow_bias, C.byref(fr)))
        return out

    def render_device(self, cams, flags, shadow_bias, rgba_ptr, id_ptr, t_ptr):
        """Enqueue a batch writing planes in the first device's memory (asynchronous; sync() or order after `stream`)."""
        arr = cams if isinstance(cams, C.Array) else (RtoCamera * len(cams))(*cams)
        fr = RtoFrame(rgba_ptr, id_ptr, t_ptr, MEM_DEVICE)
        check(lib().rto_group_render_batch(self.h, arr, len(arr), flags, shadow_bias, C.byref(fr)))

    def last_ms(self):
        ms = np.zeros(len(self.devices), np.float32)
        check(lib().rto_group_last_ms(self.h, _p(ms)))
        return ms

    @property
    def stream(self):
        return lib().rto_group_stream(self.h)

    def sync(self):
        check(lib().rto_group_sync(self.h))

    def close(self):
        if self.h:
            lib().rto_group_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BVH:
    """BVH(triangles) + query(origin, direction) (BVH.h:44-63): tree built on the host, queries run on the GPU."""

    def __init__(self, tris):
        self.host = HostBVH(tris)
        self.scene = Scene.bvh(None, prebuilt=self.host)

    def query(self, origin, direction):
        off, ids = self.scene.query(np.asarray(origin, np.float32).reshape(1, 3), np.asarray(direction, np.float32).reshape(1, 3))
        return ids

    def query_batch(self, origins, dirs):
        return self.scene.query(origins, dirs)


class RayTracerBVH:
    """RayTracerBVH (RayTracerBVH.h:28-80) without OpenGL: scenes live in HBM, frames come back as arrays."""

    def __init__(self):
        self.scene = None
        self.mode = MODE_OCTREE_GLSL
        self.shadow_bias = 0.0

    def set_octree(self, nodes, grid):
        """setOctree(root, grid): `nodes` is the flattened octree (create_octree_from_voxel_grid)."""
        if self.scene is not None:
            self.scene.close()
        self.scene = None if nodes is None or len(nodes) == 0 else Scene.octree(nodes, grid.min, grid.voxel_size)
        self.mode = MODE_OCTREE_GLSL

    def set_mesh(self, tris, scene_scale=1.0):
        """Extension: ray-cast the triangle soup through the reference-shaped BVH (north_star's mesh path)."""
        if self.scene is not None:
            self.scene.close()
        self.scene = Scene.bvh(tris)
        self.mode = MODE_BVH
        self.shadow_bias = float(np.float32(1e-3) * np.float32(scene_scale))

    def ensure_compute_initialized(self):
        check(lib().rto_init(0))

    def render_scene_compute(self, camera, width, height, aspect, fov_deg, mode=None, flags=0):
        """renderSceneCompute(camera, w, h, aspect, fovDeg) -> dict(rgba, id, t) as (h, w, ...) arrays.
        Like the reference (RayTracerBVH.cpp:624-627) an empty scene renders nothing: returns None."""
        if self.scene is None:
            return None
        cam, _ = camera.consts(fov_deg, aspect, width, height)
        out = self.scene.render(cam, self.mode if mode is None else mode, flags, self.shadow_bias)
        return dict(rgba=out["rgba"].reshape(height, width, 4), id=out["id"].reshape(height, width), t=out["t"].reshape(height, width))


def generate_test_volume(dim):
    """Multi-shell sphere of generateTestVolume (main.cpp:337-372) as a VoxelGrid set up like main.cpp:1050-1070:
    FILLED iff rInner <= |p - c| <= rOuter, grid min -0.5, voxel 1/dim."""
    c = np.float32(0.5) * np.float32(dim - 1)
    r_out = np.float32(0.4) * np.float32(dim)
    r_in = np.float32(0.2) * np.float32(dim)
    ax = np.arange(dim, dtype=np.float32) - c
    dz, dy, dx = np.meshgrid(ax, ax, ax, indexing="ij")
    dist = np.sqrt(dx * dx + dy * dy + dz * dz)
    filled = ~((dist < r_in) | (dist > r_out))
    return VoxelGrid((dim, dim, dim), (-0.5, -0.5, -0.5), np.float32(1.0) / np.float32(dim), filled.astype(np.uint8).ravel())


def city_block_grid(dim, seed, blocks, max_height=None):
    """Synthetic city-block voxel grid: `blocks` x `blocks` lots on the x-z ground plane,
    a building on a lot with p = 0.7, footprint inset U{1..4} voxels, height U{8..max_height} voxels (y up), voxel 1.0,
    grid min = -dim/2.  RNG: std::mt19937(seed), three raw 32-bit draws r0, r1, r2 per lot in (z-lot, x-lot) order:
    present = r0 < floor(0.7 * 2^32), inset = 1 + r1 % 4, height = 8 + r2 % (max_height - 7).  (numpy's RandomState(seed) is the
    same generator with the same seeding, so a C++ caller reproduces the grid with <random> alone.)"""
    lot = dim // blocks
    max_height = (3 * dim) // 4 if max_height is None else max_height
    raw = np.random.RandomState(seed).randint(0, 2 ** 32, size=3 * blocks * blocks, dtype=np.uint32).astype(np.uint64).reshape(blocks, blocks, 3)
    vol = np.zeros((dim, dim, dim), np.uint8)            # [z, y, x]
    for bz in range(blocks):
        for bx in range(blocks):
            r0, r1, r2 = (int(v) for v in raw[bz, bx])
            present = r0 < int(0.7 * 2 ** 32)
            inset = 1 + r1 % 4
            height = 8 + r2 % (max_height - 7)
            if not present or lot - 2 * inset <= 0:
                continue
            x0, z0 = bx * lot + inset, bz * lot + inset
            vol[z0:z0 + lot - 2 * inset, 0:height, x0:x0 + lot - 2 * inset] = 1
    h = np.float32(dim) * np.float32(0.5)
    return VoxelGrid((dim, dim, dim), (-h, -h, -h), 1.0, vol.ravel())
