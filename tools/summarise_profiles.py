#!/usr/bin/env python
"""Turn the ncu captures a GPU pass left in gpurun_out/ (scratch) into the tracked summaries under profiles/.
    python tools/summarise_profiles.py [--round r01] [--tag v5]
For every gpurun_out/prof_<name>.ncu-rep: profiles/<round>_<name>_<tag>.txt with the counters the north star asks for
(achieved L2/HBM throughput, FP32/ALU pipe utilisation, issue utilisation, warp-execution efficiency), a histogram of issue slots
by active lanes from the source page, and the per-launch DRAM traffic; profiles/dram_traffic.json is rewritten from the BVH capture."""
import argparse, collections, csv, glob, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_bytes.sum.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * mult.get(unit, 1)


def summarise(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = ["Kernel Name  %s" % vals[hdr.index("Kernel Name")], "Grid %s  Block %s" % (vals[hdr.index("Grid Size")], vals[hdr.index("Block Size")])]
    got = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append("%-92s %s %s" % (w, vals[i], units[i]))
            got[w] = (vals[i], units[i])
    src = ncu_csv(rep, "source")
    h = src[1]
    iI, iT = h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    buckets, tot, thr = collections.Counter(), 0, 0
    for r in src[2:]:
        try:
            ie, te = int(r[iI]), int(r[iT])
        except (ValueError, IndexError):
            continue
        if ie:
            buckets[int((te / ie) // 4) * 4] += ie
            tot += ie
            thr += te
    lines.append("")
    lines.append("issue slots by active lanes (source page, %d warp instructions, %.2f lanes on average = warp-execution efficiency %.1f %%):" % (tot, thr / max(tot, 1), 100.0 * thr / max(tot, 1) / 32))
    for k in sorted(buckets):
        lines.append("  lanes %2d-%2d: %5.1f %%" % (k, min(k + 3, 32), 100.0 * buckets[k] / tot))
    traffic = None
    if "dram__bytes_read.sum" in got and "dram__bytes_write.sum" in got:
        traffic = to_bytes(*got["dram__bytes_read.sum"]) + to_bytes(*got["dram__bytes_write.sum"])
        lines.append("")
        lines.append("DRAM traffic per launch: %.1f MB (read + write)" % (traffic / 1e6))
    def num(key):
        return float(got[key][0].replace(",", "")) if key in got else None
    counters = {"kernel": vals[hdr.index("Kernel Name")], "grid": vals[hdr.index("Grid Size")], "ncu_ms": num("gpu__time_duration.sum"),
                "dram_bytes_per_launch": int(traffic) if traffic else None, "warp_instructions": num("smsp__inst_executed.sum"),
                "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"), "lanes_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
                "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"),
                "l1_throughput_pct": num("l1tex__throughput.avg.pct_of_peak_sustained_active"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
                "l2_throughput_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"), "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "long_scoreboard_warps_per_issue": num("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")}
    return "\n".join(lines) + "\n", traffic, counters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--tag", default="v5")
    ap.add_argument("--src", default=os.path.join(ROOT, "gpurun_out"))
    a = ap.parse_args()
    for rep in sorted(glob.glob(os.path.join(a.src, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        if name in ("bvh_pt", "bvh_warp", "bvh_r1", "oct_A_fast", "oct_B"):
            continue                                         # captures of variants that were removed; summarised by hand in README.md
        text, traffic, counters = summarise(rep)
        out = os.path.join(ROOT, "profiles", "%s_%s_%s_full.txt" % (a.round, name, a.tag))
        open(out, "w").write(text)
        print("wrote", out)
        if name == "bvh":
            counters["source"] = "profiles/" + os.path.basename(out) + " (ncu --set full --clock-control none of one k_render_bvh<shadows,pruned> launch under bench.py)"
            json.dump(counters, open(os.path.join(ROOT, "profiles", "%s_bvh_counters.json" % a.round), "w"), indent=1)
        if name == "bvh" and traffic:
            json.dump({"k_render_bvh_bytes_per_launch": int(traffic), "source": os.path.basename(out),
                       "what": "dram__bytes_read.sum + dram__bytes_write.sum of one k_render_bvh<shadows,pruned> launch (8 x 1080p frames), ncu --set full"},
                      open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
