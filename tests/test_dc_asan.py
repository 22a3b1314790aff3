"""The Dual-Contouring builders under AddressSanitizer + UBSan (tests/dc_asan/driver.cpp): host_dc.cpp and the per-cell code of
rto_dc.h -- the code the CUDA kernels of rto_dc.cu share -- run on exactly-sized heap buffers, so an out-of-range voxel, node or
record index fails here.  (compute-sanitizer is not available on the GPU pool.)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ray_tracing_octrees_b200", "csrc")


def test_dc_builders_are_clean_under_asan(tmp_path):
    exe = str(tmp_path / "dc_asan")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer",
           os.path.join(ROOT, "tests", "dc_asan", "driver.cpp"), os.path.join(CSRC, "host_dc.cpp"), os.path.join(CSRC, "host_builders.cpp"),
           os.path.join(CSRC, "host_layouts.cpp"), "-lpthread", "-o", exe]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all ok" in out.stdout and "MISMATCH" not in out.stdout
