"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

  ref()  -> oracle/_ref/libref.so     the reference's own translation units compiled in place
                                       (oracle/Makefile `ref`) + the extension rules of SURVEY.md 8c
                                       written with real glm types (oracle/ref_harness.cpp)
  port() -> oracle/_build/liboracle.so the stand-alone restatement (oracle/oracle_port.cpp)

Both export the same C API under the prefixes `ref_` / `orc_`, so tests can run either.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libref.so")
PORT_SO = os.path.join(_HERE, "_build", "liboracle.so")


class CamConsts(C.Structure):
    """Mirrors RefCamConsts (ref_harness.cpp), Cam (oracle_port.cpp) and RtoCamera (include/rto_c.h)."""
    _fields_ = [("camPos", C.c_float * 3), ("invView", C.c_float * 16), ("tanHalfFov", C.c_float),
                ("aspect", C.c_float), ("width", C.c_int), ("height", C.c_int)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_port():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def build_ref():
    """Only possible where /root/reference exists (this container); the GPU box uses the prebuilt file."""
    if os.path.isdir("/root/reference/453-skeleton"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


class Backend:
    def __init__(self, path, prefix, kind):
        self.kind = kind           # "reference" | "port"  (bench.py cpu_baseline.kind)
        L = C.CDLL(path)
        vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        sig = {
            "grid_create": (vp, [i32, i32, i32, f32, f32, f32, f32, vp]),
            "grid_load": (vp, [C.c_char_p]),
            "grid_from_csv": (vp, [C.c_char_p, C.c_char_p, f32]),
            "grid_info": (None, [vp, vp, vp]),
            "grid_data": (None, [vp, vp]),
            "octree_build": (i32, [vp]),
            "octree_flat": (None, [vp, vp]),
            "octree_free": (None, [vp]),
            "camera_consts": (None, [f32, f32, f32, vp, f32, f32, i32, i32, C.POINTER(CamConsts), vp]),
            "mesh_from_octree": (vp, [vp]),
            "mesh_from_tris": (vp, [vp, sz]),
            "mesh_count": (sz, [vp]),
            "mesh_tris": (None, [vp, vp]),
            "bvh_build": (C.c_double, [vp]),
            "mesh_free": (None, [vp]),
            "bvh_export": (sz, [vp, vp, vp, sz]),
            "bvh_query": (sz, [vp, vp, vp, sz, vp, vp, sz]),
            "render_bvh": (C.c_double, [vp, C.POINTER(CamConsts), C.c_uint, f32, i32, i32, vp, vp, vp, vp, i32]),
            "render_octree": (C.c_double, [vp, C.POINTER(CamConsts), i32, i32, i32, vp, vp, vp, vp, i32]),
            "octree_rayskip": (None, [vp, vp, vp, sz, f32, f32, vp, vp]),
            "num_threads": (i32, []),
            "cull_nodes": (sz, [vp, f32, f32, f32, vp, f32, f32, i32, i32, vp]),
            "skip_distance": (f32, [vp, f32, f32, f32, vp, f32, f32, vp, vp, vp]),
        }
        sig["dc_mesh_from_octree"] = (vp, [vp, vp, f32])
        if hasattr(L, prefix + "dc_mesh_full"):          # the compiled reference only: MCTriangles with their normals
            sig["dc_mesh_full"] = (sz, [vp, vp, f32, vp, sz])
        for name, (res, args) in sig.items():
            fn = getattr(L, prefix + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    # -- convenience constructors -----------------------------------------------------------------
    def octree(self, dims=None, gmin=None, voxel=None, data=None, path=None, csv=None):
        return Octree(self, dims, gmin, voxel, data, path, csv)

    def mesh(self, tris):
        return Mesh(self, tris=tris)

    def camera(self, theta_deg, phi_deg, radius, target=(0, 0, 0), fov_deg=45.0, width=1024, height=768, aspect=None):
        """Camera(theta,phi,r) + inverse(view) + tan(fov/2) -> CamConsts.  Angles given in degrees are
        converted to radians in float32 (np.deg2rad on float32), exactly what callers must pass on."""
        cam = CamConsts()
        view = np.zeros(16, np.float32)
        tgt = np.asarray(target, np.float32)
        aspect = float(np.float32(width) / np.float32(height)) if aspect is None else aspect
        self.camera_consts(float(np.deg2rad(np.float32(theta_deg))), float(np.deg2rad(np.float32(phi_deg))), radius,
                           _p(tgt), fov_deg, aspect, width, height, C.byref(cam), _p(view))
        return cam, view


_ref = _port = None


def ref_available():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = Backend(REF_SO, "ref_", "reference")
    return _ref


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build_port()
        _port = Backend(PORT_SO, "orc_", "port")
    return _port


def best():
    """The strongest checker available: the compiled reference if present, else the port."""
    return ref() if ref_available() else port()

class Octree:
    def __init__(self, L, dims=None, gmin=None, voxel=None, data=None, path=None, csv=None):
        self.L = L
        if csv is not None:                      # (verts.csv, faces.csv, voxelSize): loadCSVDataIntoVoxelGrid, BuildingLoader.cpp:153-290
            self.h = L.grid_from_csv(os.fsencode(csv[0]), os.fsencode(csv[1]), float(csv[2]))
        elif path is not None:
            self.h = L.grid_load(path.encode())
            if not self.h:
                raise IOError(path)
        else:
            data = np.ascontiguousarray(data, dtype=np.uint8)
            assert data.size == dims[0] * dims[1] * dims[2]
            self.h = L.grid_create(dims[0], dims[1], dims[2], gmin[0], gmin[1], gmin[2], voxel, _p(data))
        d = np.zeros(3, np.int32)
        m = np.zeros(4, np.float32)
        L.grid_info(self.h, _p(d), _p(m))
        self.dims, self.gmin, self.voxel = tuple(int(x) for x in d), m[:3].copy(), float(m[3])
        self.num_nodes = 0

    def grid_data(self):
        out = np.zeros(self.dims[0] * self.dims[1] * self.dims[2], np.uint8)
        self.L.grid_data(self.h, _p(out))
        return out

    def build(self):
        n = self.L.octree_build(self.h)
        if n < 0:
            raise RuntimeError("BFS order check against RayTracerBVH::setOctree failed: %d" % n)
        self.num_nodes = n
        return n

    def cull(self, theta_deg, phi_deg, radius, fov_deg=45.0, aspect=16.0 / 9.0, width=1920, height=1080, target=(0, 0, 0)):
        """The node array RayTracerBVH::renderSceneComputeWithCulling(camera, w, h, aspect, fov, true) uploads (RayTracerBVH.cpp:724-813)."""
        tgt = np.asarray(target, np.float32)
        th, ph = float(np.deg2rad(np.float32(theta_deg))), float(np.deg2rad(np.float32(phi_deg)))
        n = self.L.cull_nodes(self.h, th, ph, radius, _p(tgt), fov_deg, aspect, width, height, None)
        out = np.zeros((n, 15), np.int32)
        if n:
            self.L.cull_nodes(self.h, th, ph, radius, _p(tgt), fov_deg, aspect, width, height, _p(out))
        return out

    def skip_distance(self, theta_deg, phi_deg, radius, aspect, last=0.0, target=(0, 0, 0)):
        """VolumeRaycastRenderer.cpp:1598-1664 -> (skip distance, 49 probe t, origins (49,3), dirs (49,3))."""
        tgt = np.asarray(target, np.float32)
        t, o, d = np.zeros(49, np.float32), np.zeros((49, 3), np.float32), np.zeros((49, 3), np.float32)
        v = self.L.skip_distance(self.h, float(np.deg2rad(np.float32(theta_deg))), float(np.deg2rad(np.float32(phi_deg))), radius, _p(tgt), aspect, last, _p(t), _p(o), _p(d))
        return float(v), t, o, d

    def flat(self):
        out = np.zeros((self.num_nodes, 15), np.int32)
        self.L.octree_flat(self.h, _p(out))
        return out

    def render(self, cam, mode, y0=0, y1=None, stats=False, threads=0, want=True):
        y1 = cam.height if y1 is None else y1
        n = (y1 - y0) * cam.width
        rgba = np.zeros((n, 4), np.float32) if want else None
        ids = np.zeros(n, np.int32) if want else None
        t = np.zeros(n, np.float32) if want else None
        st = np.zeros(2, np.uint64) if stats else None
        sec = self.L.render_octree(self.h, C.byref(cam), mode, y0, y1, _p(rgba), _p(ids), _p(t), _p(st), threads)
        return dict(rgba=rgba, id=ids, t=t, sec=sec, stats=st)

    def rayskip(self, o, d, tmin=0.0, tmax=1e30):
        o = np.ascontiguousarray(o, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        out = np.zeros(len(o), np.float32)
        ids = np.zeros(len(o), np.int32)
        self.L.octree_rayskip(self.h, _p(o), _p(d), len(o), tmin, tmax, _p(out), _p(ids))
        return out, ids

    def mesh(self):
        return Mesh(self.L, handle=self.L.mesh_from_octree(self.h))

    def dc_mesh(self, view_proj=None, margin=50.0):
        """renderOctree (main.cpp:95-208) over AdaptiveDualContouringRenderer::render: the DC triangle soup; view_proj None = no culling."""
        vp = None if view_proj is None else np.ascontiguousarray(view_proj, np.float32).ravel()
        return Mesh(self.L, handle=self.L.dc_mesh_from_octree(self.h, _p(vp), float(margin)))

    def dc_mesh_full(self, view_proj=None, margin=50.0):
        """The MCTriangles of the Dual-Contouring run, (m, 18): 3 vertices + 3 normals (compiled reference only)."""
        vp = None if view_proj is None else np.ascontiguousarray(view_proj, np.float32).ravel()
        out = np.zeros((1 << 16, 18), np.float32)
        n = self.L.dc_mesh_full(self.h, _p(vp), float(margin), _p(out), len(out))
        if n > len(out):
            out = np.zeros((n, 18), np.float32)
            self.L.dc_mesh_full(self.h, _p(vp), float(margin), _p(out), n)
        return out[:n].copy()

    def free(self):
        if self.h:
            self.L.octree_free(self.h)
            self.h = None


class Mesh:
    def __init__(self, L, tris=None, handle=None):
        self.L = L
        if handle is None:
            tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
            handle = L.mesh_from_tris(_p(tris), len(tris))
        self.h = handle
        self.n = L.mesh_count(self.h)
        self.build_sec = None

    def tris(self):
        out = np.zeros((self.n, 9), np.float32)
        self.L.mesh_tris(self.h, _p(out))
        return out

    def build(self):
        self.build_sec = self.L.bvh_build(self.h)
        return self.build_sec

    def export(self):
        n = self.L.bvh_export(self.h, None, None, 0)
        boxes = np.zeros((n, 6), np.float32)
        meta = np.zeros((n, 4), np.int32)
        self.L.bvh_export(self.h, _p(boxes), _p(meta), n)
        return boxes, meta

    def query(self, o, d):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        off = np.zeros(len(o) + 1, np.int64)
        total = self.L.bvh_query(self.h, _p(o), _p(d), len(o), _p(off), None, 0)
        ids = np.zeros(total, np.int32)
        self.L.bvh_query(self.h, _p(o), _p(d), len(o), _p(off), _p(ids), total)
        return off, ids

    def render(self, cam, flags=0, bias=0.0, y0=0, y1=None, stats=False, threads=0, want=True):
        y1 = cam.height if y1 is None else y1
        n = (y1 - y0) * cam.width
        rgba = np.zeros((n, 4), np.float32) if want else None
        ids = np.zeros(n, np.int32) if want else None
        t = np.zeros(n, np.float32) if want else None
        st = np.zeros(5, np.uint64) if stats else None
        sec = self.L.render_bvh(self.h, C.byref(cam), flags, bias, y0, y1, _p(rgba), _p(ids), _p(t), _p(st), threads)
        return dict(rgba=rgba, id=ids, t=t, sec=sec, stats=st)

    def free(self):
        if self.h:
            self.L.mesh_free(self.h)
            self.h = None



def load_scene_cache(path):
    """Parse a (optionally gzip-compressed) sceneCache.bin (CacheUtils.cpp:5-59) with numpy ->
    (dims, gmin, voxel, data).  Avoids the reference loader's std::cout chatter in bench output."""
    import gzip
    raw = gzip.open(path, "rb").read() if path.endswith(".gz") else open(path, "rb").read()
    dims = tuple(int(x) for x in np.frombuffer(raw, np.int32, 3, 0))
    mv = np.frombuffer(raw, np.float32, 4, 12)
    n = int(np.frombuffer(raw, np.uint64, 1, 28)[0])
    assert n == dims[0] * dims[1] * dims[2]
    return dims, mv[:3].copy(), float(mv[3]), np.frombuffer(raw, np.uint8, n, 36).copy()


def sphere_grid(dim):
    """generateTestVolume closed form (main.cpp:337-372) + grid setup of main.cpp:1050-1070."""
    c = np.float32(0.5) * np.float32(dim - 1)
    r_out = np.float32(0.4) * np.float32(dim)
    r_in = np.float32(0.2) * np.float32(dim)
    ax = (np.arange(dim, dtype=np.float32) - c)
    dz, dy, dx = np.meshgrid(ax, ax, ax, indexing="ij")
    dist = np.sqrt((dx * dx + dy * dy + dz * dz).astype(np.float32)).astype(np.float32)
    filled = ~((dist < r_in) | (dist > r_out))
    return (dim, dim, dim), (-0.5, -0.5, -0.5), float(np.float32(1.0) / np.float32(dim)), filled.astype(np.uint8).ravel()
