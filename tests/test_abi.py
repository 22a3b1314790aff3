"""CPU only: the C-ABI library loads, exports every symbol include/rto_c.h declares, and refuses to trace
without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rto_c.h")).read()
    return sorted(set(re.findall(r"RTO_API[^;(]*?\b(rto_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(rto):
    from ray_tracing_octrees_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "librto.so does not export " + name
    assert sorted(_lib.SIGNATURES) == declared, "python binding and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (rto_[a-z0-9_]+)", out)))
    assert exported == declared, "exported symbols differ from the header: %s" % (set(exported) ^ set(declared))


def test_library_is_sm100a_with_lineinfo(rto):
    from ray_tracing_octrees_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts(rto):
    assert ctypes.sizeof(rto.RtoCamera) == 92
    assert ctypes.sizeof(rto.RtoFrame) == 32
    from oracle import bind
    assert ctypes.sizeof(bind.CamConsts) == ctypes.sizeof(rto.RtoCamera)


def test_version_and_error_strings(rto):
    assert b"sm_100a" in rto.lib().rto_version()


def test_no_cpu_fallback_without_gpu(rto):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present; the refusal path is for GPU-less hosts")
    tris = np.random.default_rng(0).normal(0, 1, (8, 9)).astype(np.float32)
    with pytest.raises(rto.RtoError) as e:
        rto.Scene.bvh(tris)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    g = rto.generate_test_volume(8)
    nodes = rto.create_octree_from_voxel_grid(g)
    with pytest.raises(rto.RtoError) as e:
        rto.Scene.octree(nodes, g.min, g.voxel_size)
    assert e.value.code == 2
    with pytest.raises(rto.RtoError) as e:
        rto.Scene.load("/nonexistent/scene.rtoscene")      # a cached scene is a DEVICE layout: nothing to do with it without one
    assert e.value.code == 2


def test_missing_library_fails_loudly(rto, monkeypatch):
    from ray_tracing_octrees_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librto.so")
    with pytest.raises(ImportError):
        _lib.lib()
