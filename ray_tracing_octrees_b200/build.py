"""Build librto.so (the C-ABI product library) in-tree with nvcc for sm_100a.

    python -m ray_tracing_octrees_b200.build [--force]

Flags that matter for parity with the CPU oracle: -fmad=false (no fused multiply-add), IEEE division and
square root, no flush-to-zero; host code gets -ffp-contract=off.  -lineinfo keeps ncu's source view usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librto.so")
SOURCES = ["rto_device.cu", "rto_group.cu", "rto_build.cu", "rto_dc.cu", "rto_blob.cu", "host_builders.cpp", "host_layouts.cpp", "host_dc.cpp"]
DEPS = SOURCES + ["rto_kernels.cuh", "rto_scene.cuh", "rto_devtypes.h", "rto_voxelize.h", "rto_frustum.h", "rto_dc.h", "rto_sahchunk.h", "rto_sort.cuh", "rto_internal.h", "rto_nvtx.h", "rto_math.h", "mc_tables.h", "../../include/rto_c.h"]


def nvcc_cmd(extra=(), out=None):
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    return ["nvcc", "-ccbin", ccbin, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
            "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-O2", "-shared",
            *extra, "-o", out or OUT, *[os.path.join(CSRC, s) for s in SOURCES], "-lpthread", "-ldl"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = nvcc_cmd(["-Xptxas", "-v"] if verbose else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building librto.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


def build_variant(name, defines):
    """Experimental build librto_<name>.so with extra -D flags (A/B kernel timing; never loaded unless RTO_LIB_VARIANT=<name>)."""
    out = os.path.join(HERE, "librto_%s.so" % name)
    r = subprocess.run(nvcc_cmd(list(defines), out), capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building " + out)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
