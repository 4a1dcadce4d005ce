"""Turns an ncu launch list of scripts/kbench.py (metrics: dram__bytes_read.sum, dram__bytes_write.sum,
smsp__inst_executed_pipe_alu.sum, smsp__inst_executed.sum, gpu__time_duration.sum) into the per-pair / per-cell
figures bench.py reports as roofline.traffic and roofline.alu_instr_per_cell (run here, CPU only):
    python scripts/ncu_counters.py <launches.csv> <kbench.json> <out.json>"""
import csv, json, sys

def main(csv_path, kbench_path, out_path):
    kb = json.loads(open(kbench_path).read().strip().splitlines()[-1])
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(r[iid], {"k": r[ik]})
        d[r[im]] = float(r[iv].replace(",", ""))
    # the timed DP launches: every bsw_ DP kernel that is not a COUNT variant (COUNT = third template argument)
    def is_dp(name):
        return any(x in name for x in ("bsw_short_kernel", "bsw_win_kernel", "bsw_long_kernel", "bsw_duo2_kernel"))
    def is_count(name):
        args = name[name.index("<") + 1:name.index(">")].replace("(bool)", "").split(",") if "<" in name else []
        return "duo2" not in name and len(args) >= 3 and args[2].strip() == "1"
    dp = [d for d in launches.values() if is_dp(d["k"]) and not is_count(d["k"])]
    runs = kb["runs"]
    tot = lambda m: sum(d.get(m, 0.0) for d in dp) / runs
    by_kernel = {}
    for d in dp:
        nm = d["k"].split("(")[0]
        by_kernel[nm] = by_kernel.get(nm, 0.0) + d.get("gpu__time_duration.sum", 0.0) / runs
    top = max(by_kernel, key=by_kernel.get)
    out = {
        "kernel": f"{top} ({100 * by_kernel[top] / sum(by_kernel.values()):.0f} % of the DP kernels' time)",
        "workload": kb["workload"], "pairs": kb["pairs"], "cells": kb["cells"], "dp_launches_per_step": len(dp) // runs,
        "dram_bytes_per_pair": (tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")) / kb["pairs"],
        "dram_bytes_per_step": tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum"),
        "alu_instr_per_cell": tot("smsp__inst_executed_pipe_alu.sum") * 32 / kb["cells"],
        "instr_per_cell": tot("smsp__inst_executed.sum") * 32 / kb["cells"],
        "dp_kernel_ms_per_step_under_ncu": tot("gpu__time_duration.sum") / 1e6,
        "source": f"ncu launch list {csv_path} (dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed_pipe_alu.sum x 32 "
                  f"per visited cell; all DP launches of {runs} resident passes of scripts/kbench.py, averaged)",
    }
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main(*sys.argv[1:4])
