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
