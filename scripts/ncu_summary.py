"""Turns an .ncu-rep into a compact per-launch CSV of the metrics DESIGN.md cites (run here, CPU only)."""
import csv, io, subprocess, sys

KEEP = """launch__registers_per_thread launch__shared_mem_per_block_dynamic launch__occupancy_limit_shared_mem
gpu__time_duration.sum sm__warps_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.avg.per_cycle_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio dram__bytes_read.sum dram__bytes_write.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
smsp__inst_executed_op_shared_ld.sum smsp__inst_executed_op_shared_st.sum sm__cycles_elapsed.max sm__cycles_active.avg""".split()

def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx] + [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{k}" for k in range(len(data))])
        for name in ["Kernel Name", "Grid Size", "Block Size"] + cols:
            w.writerow([name, units[idx[name]]] + [r[idx[name]] for r in data])
    print("wrote", out, len(data), "launches")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
