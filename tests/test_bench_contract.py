"""bench.py's reference arm (`--impl reference`: the reference's own CPU code, oracle/_ref/libref.so, or the port where that is absent) runs
without a GPU and prints the one JSON line the driver reads; the keys and the strings that have to agree with the CUDA arm's line are checked
here (the CUDA arm itself needs a GPU: the driver runs it, `profiles/r02_bench_n1_v6.json` is its last line)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["unit"] == "Mrays/s" and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # metric and config are the CUDA arm's, string for string: both come from the same constants of bench.py
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert d["metric"] in src and d["config"]["workload"].split(" (")[0] in src
    assert re.search(r"1920x1080", d["config"]["image"]) and "shadow" in d["config"]["rays"]
    # the line a GPU run of the other arm left behind names the same metric and the same workload
    last = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1_v6.json")).read().strip().splitlines()[-1])
    assert last["metric"] == d["metric"] and last["config"] == d["config"] and last["unit"] == d["unit"]
    for key in ("roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert key in last
    assert 0 < last["roofline"]["frac"] <= 1.0 and last["roofline"]["bound"] != "hbm"
