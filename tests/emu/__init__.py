"""CPU emulation of the product's traversal code (test infrastructure; see emu.cu)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ray_tracing_octrees_b200", "csrc")
# RTO_EMU_DEFINES="-DNAME=1 ...": the emulator of a kernel variant (its own .so), to check a variant on the CPU before it costs GPU time
EXTRA = os.environ.get("RTO_EMU_DEFINES", "").split()
SO = os.path.join(HERE, "_build", "libemu%s.so" % ("_" + "".join(c if c.isalnum() else "_" for c in "".join(EXTRA)) if EXTRA else ""))
SRCS = [os.path.join(HERE, "emu.cu"), os.path.join(CSRC, "host_builders.cpp"), os.path.join(CSRC, "host_layouts.cpp")]
DEPS = SRCS + [os.path.join(CSRC, f) for f in ("rto_kernels.cuh", "rto_internal.h", "rto_math.h")]


def build():
    if os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in DEPS):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call(["nvcc", "-ccbin", ccbin, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-fmad=false",
                           "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", *EXTRA, "-o", SO, *SRCS, "-lpthread"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.emu_octree_create.restype = vp
        L.emu_octree_create.argtypes = [vp, sz, vp, f32]
        L.emu_octree_free.argtypes = [vp]
        L.emu_render_octree.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
        L.emu_trace_octree.argtypes = [vp, i32, i32, vp, vp, sz, f32, f32, vp, vp]
        L.emu_bvh_create.restype = vp
        L.emu_bvh_create.argtypes = [vp, sz]
        L.emu_bvh_free.argtypes = [vp]
        L.emu_sah_chunk.argtypes = [vp, i32, i32, i32]
        L.emu_sah_chunk.restype = C.c_double
        L.emu_bvh_wide_nodes.argtypes = [vp]
        L.emu_bvh_wide_nodes.restype = i32
        L.emu_render_bvh.argtypes = [vp, vp, C.c_uint, f32, i32, i32, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Octree:
    def __init__(self, nodes, gmin, voxel):
        self.nodes = np.ascontiguousarray(nodes, np.int32)
        gm = np.asarray(gmin, np.float32)
        self.h = lib().emu_octree_create(_p(self.nodes), len(self.nodes), _p(gm), float(voxel))
        assert self.h

    def render(self, cam, mode, count=False, y0=0, y1=None):
        y1 = cam.height if y1 is None else y1
        n = (y1 - y0) * cam.width
        out = dict(rgba=np.empty((n, 4), np.float32), id=np.empty(n, np.int32), t=np.empty(n, np.float32))
        v = np.zeros(1, np.uint64)
        out["pixel_visits"] = np.zeros(n, np.uint32)
        lib().emu_render_octree(self.h, C.byref(cam), mode, int(count), y0, y1, _p(out["rgba"]), _p(out["id"]), _p(out["t"]), _p(v), _p(out["pixel_visits"]))
        out["visits"] = int(v[0])
        return out

    def trace(self, o, d, mode, count=False, tmin=0.0, tmax=1e30):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        t = np.empty(len(o), np.float32)
        ids = np.empty(len(o), np.int32)
        lib().emu_trace_octree(self.h, mode, int(count), _p(o), _p(d), len(o), tmin, tmax, _p(t), _p(ids))
        return t, ids


class Bvh:
    def __init__(self, tris):
        self.tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        self.h = lib().emu_bvh_create(_p(self.tris), len(self.tris))
        assert self.h

    WIDE = 0x100            # render(): walk the 4-wide quantised form (present when the Bvh was created with RTO_BVH_WIDE=1)

    @property
    def wide_nodes(self):
        return lib().emu_bvh_wide_nodes(self.h)

    def render(self, cam, flags=0, bias=0.0, y0=0, y1=None):
        y1 = cam.height if y1 is None else y1
        n = (y1 - y0) * cam.width
        out = dict(rgba=np.empty((n, 4), np.float32), id=np.empty(n, np.int32), t=np.empty(n, np.float32))
        lib().emu_render_bvh(self.h, C.byref(cam), flags, bias, y0, y1, _p(out["rgba"]), _p(out["id"]), _p(out["t"]))
        return out


def sah_chunk(leaf_boxes, first=0, root_at_end=False):
    """rto_sahchunk.h on the CPU over (m, 6) leaf boxes -> sum of the half areas of the rebuilt subtree's boxes (negative: broken tree)."""
    b = np.ascontiguousarray(leaf_boxes, np.float32).reshape(-1, 6)
    return float(lib().emu_sah_chunk(_p(b), len(b), first, 1 if root_at_end else 0))
