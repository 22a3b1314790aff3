"""ctypes loader for librto.so (the C ABI declared in include/rto_c.h).

The library is built in-tree by ray_tracing_octrees_b200.build.  Loading fails loudly when it is missing:
there is no Python or CPU fallback for the ray kernels.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RTO_LIB_VARIANT=<name> loads librto_<name>.so, an experimental build made by `python -m ray_tracing_octrees_b200.build --variant
# <name> -D...` for A/B kernel timing (tools/profile_case.py); the product is always librto.so
LIB_PATH = os.path.join(HERE, "librto_%s.so" % os.environ["RTO_LIB_VARIANT"] if os.environ.get("RTO_LIB_VARIANT") else "librto.so")

RTO_OK = 0
MODE_BVH, MODE_OCTREE_SKIP, MODE_OCTREE_GLSL = 0, 1, 2
FLAG_SHADOWS, FLAG_NO_PRUNE, FLAG_SORT_RAYS = 1, 2, 4
MEM_HOST, MEM_DEVICE = 0, 1


class RtoCamera(C.Structure):
    _fields_ = [("camPos", C.c_float * 3), ("invView", C.c_float * 16), ("tanHalfFov", C.c_float),
                ("aspect", C.c_float), ("width", C.c_int32), ("height", C.c_int32)]


class RtoFrame(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("hitId", C.c_void_p), ("t", C.c_void_p), ("memory", C.c_int32)]


class RtoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("librto error %d: %s" % (code, msg))
        self.code = code


# every symbol include/rto_c.h declares: name -> (restype, argtypes)
_vp, _i, _u32, _u64, _f, _sz = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float, C.c_size_t
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "rto_version": (C.c_char_p, []),
    "rto_last_error": (C.c_char_p, []),
    "rto_init": (_i, [_i]),
    "rto_device_info": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "rto_host_octree_build": (_i, [_vp, _i, _i, _i, _pp, C.POINTER(_sz)]),
    "rto_host_mc_mesh": (_i, [_vp, _i, _i, _i, _vp, _f, _vp, _sz, _pp, C.POINTER(_sz)]),
    "rto_host_dc_mesh": (_i, [_vp, _i, _i, _i, _vp, _f, _vp, _sz, _vp, _f, _pp, C.POINTER(_sz)]),
    "rto_host_dc_mesh_replay": (_i, [_vp, _i, _i, _i, _vp, _f, _vp, _sz, _vp, _f, _pp, C.POINTER(_sz)]),
    "rto_host_dc_mesh_normals": (_i, [_vp, _i, _i, _i, _vp, _f, _vp, _sz, _vp, _f, _pp, _pp, C.POINTER(_sz)]),
    "rto_host_tricache_save": (_i, [C.c_char_p, _vp, _vp, _sz]),
    "rto_host_tricache_load": (_i, [C.c_char_p, _pp, _pp, C.POINTER(_sz)]),
    "rto_device_dc_mesh": (_i, [_vp, _i, _i, _i, _vp, _f, _vp, _sz, _vp, _f, _pp, C.POINTER(_sz)]),
    "rto_host_bvh_build": (_i, [_vp, _sz, _pp]),
    "rto_host_bvh_free": (None, [_vp]),
    "rto_host_bvh_num_nodes": (_sz, [_vp]),
    "rto_host_bvh_export": (_i, [_vp, _vp, _vp, _sz]),
    "rto_host_camera_orbit": (_i, [_f, _f, _f, _vp, _f, _f, _i, _i, C.POINTER(RtoCamera), _vp]),
    "rto_host_view_proj": (_i, [_vp, _f, _f, _f, _f, _vp]),
    "rto_host_frustum_cull": (_i, [_vp, _sz, _vp, _f, _vp, _f, _pp, C.POINTER(_sz), _pp]),
    "rto_device_frustum_cull": (_i, [_vp, _sz, _vp, _f, _vp, _f, _pp, C.POINTER(_sz), _pp]),
    "rto_host_grid_load": (_i, [C.c_char_p, _vp, _vp, _pp]),
    "rto_host_grid_save": (_i, [C.c_char_p, _vp, _vp, _vp]),
    "rto_host_csv_voxelize": (_i, [C.c_char_p, C.c_char_p, _f, _vp, _vp, _pp]),
    "rto_device_csv_voxelize": (_i, [C.c_char_p, C.c_char_p, _f, _vp, _vp, _pp]),
    "rto_host_free": (None, [_vp]),
    "rto_scene_create_octree": (_i, [_vp, _sz, _vp, _f, _pp]),
    "rto_scene_create_bvh": (_i, [_vp, _sz, _vp, _pp]),
    "rto_device_octree_build": (_i, [_vp, _i, _i, _i, _pp, C.POINTER(_sz)]),
    "rto_scene_create_octree_from_grid": (_i, [_vp, _i, _i, _i, _vp, _f, _pp]),
    "rto_device_mc_mesh": (_i, [_vp, _i, _i, _i, _vp, _f, _pp, C.POINTER(_sz)]),
    "rto_scene_create_bvh_device": (_i, [_vp, _sz, _pp]),
    "rto_scene_create_bvh_from_grid": (_i, [_vp, _i, _i, _i, _vp, _f, _pp]),
    "rto_scene_create_bvh_from_grid_dc": (_i, [_vp, _i, _i, _i, _vp, _f, _vp, _f, _pp]),
    "rto_scene_octree_layout_read": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_sz)]),
    "rto_scene_save": (_i, [_vp, C.c_char_p]),
    "rto_scene_load": (_i, [C.c_char_p, _vp]),
    "rto_scene_destroy": (None, [_vp]),
    "rto_scene_info": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "rto_scene_stream": (_vp, [_vp]),
    "rto_render": (_i, [_vp, C.POINTER(RtoCamera), _i, _u32, _f, _i, _i, C.POINTER(RtoFrame)]),
    "rto_render_batch": (_i, [_vp, _vp, _i, _i, _u32, _f, _i, _i, C.POINTER(RtoFrame)]),
    "rto_trace_rays": (_i, [_vp, _i, _u32, _vp, _vp, _sz, _f, _f, _vp, _vp, _i]),
    "rto_device_sort_pairs": (_i, [_vp, _vp, _sz]),
    "rto_host_alloc_pinned": (_i, [_sz, C.POINTER(_vp)]),
    "rto_host_free_pinned": (None, [_vp]),
    "rto_bvh_query": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, C.POINTER(_sz)]),
    "rto_render_stats": (_i, [_vp, C.POINTER(RtoCamera), _i, _u32, _f, _i, _i, _vp]),
    "rto_octree_skip_distance": (_i, [_vp, _vp, _vp, _f, _f, C.POINTER(_f), _vp]),
    "rto_host_skip_probe_rays": (_i, [_vp, _vp, _f, _vp, _vp]),
    "rto_host_skip_distance_from_probes": (_f, [_vp, _i, _f]),
    "rto_scene_sync": (_i, [_vp]),
    "rto_scene_last_kernel_ms": (_i, [_vp, C.POINTER(_f)]),
    "rto_scene_launch_count": (_u64, [_vp]),
    "rto_codes_frame_words": (_sz, [_i, _i]),
    "rto_render_codes": (_i, [_vp, _vp, _i, _u32, _f, _i, _i, _vp, _sz, _vp]),
    "rto_resolve_codes": (_i, [_vp, _vp, _i, _i, _i, _vp, _sz, C.POINTER(RtoFrame), _vp]),
    "rto_exchange_alloc": (_i, [_sz, _pp, _vp]),
    "rto_exchange_free": (_i, [_vp]),
    "rto_exchange_open": (_i, [_vp, _pp]),
    "rto_exchange_close": (_i, [_vp]),
    "rto_group_create": (_i, [_vp, _i, _pp]),
    "rto_group_destroy": (None, [_vp]),
    "rto_group_size": (_i, [_vp]),
    "rto_group_scene_bvh": (_i, [_vp, _vp, _sz, _vp]),
    "rto_group_scene": (_vp, [_vp, _i]),
    "rto_group_render_batch": (_i, [_vp, _vp, _i, _u32, _f, C.POINTER(RtoFrame)]),
    "rto_group_set_balancing": (_i, [_vp, _i, _vp, _i]),
    "rto_group_last_ms": (_i, [_vp, _vp]),
    "rto_group_stream": (_vp, [_vp]),
    "rto_group_sync": (_i, [_vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("librto.so is not built (%s). Run `python -m ray_tracing_octrees_b200.build`; "
                              "there is no fallback implementation." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("RTO_LIB_VARIANT") and not hasattr(L, name):
                continue                    # experimental builds of older sources may lack newer entry points
            fn = getattr(L, name)           # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != RTO_OK:
        raise RtoError(rc, lib().rto_last_error().decode())
