#!/usr/bin/env python3
"""Summarise an .ncu-rep here (no GPU needed): headline metrics, warp-stall ratios and the hottest SASS lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_hot_lines]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__t_set_accesses.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_set_accesses_pipe_lsu.sum", "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:82s} {vals[i]:>22s} {units[i]}")
print()
st = [(h, float(vals[i])) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and vals[i]]
if not st:
    st = [(h, float(vals[i])) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_warp_active.pct") and vals[i]]
for h, v in sorted(st, key=lambda kv: -kv[1])[:10]:
    print(f"{h:82s} {v:10.3f}")
print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
ix = {h: i for i, h in enumerate(h2)}
data = rows[2:]
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
tott = sum(int(r[ix["Thread Instructions Executed"]]) for r in data)
samples = sum(int(r[ix["# Samples"]]) for r in data)
print(f"warp instructions {tot}, thread instructions {tott} ({tott / max(tot, 1):.2f} per instruction), samples {samples}")
print("hottest SASS by stall samples:")
stall_cols = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:nhot]:
    top = sorted(((int(r[ix[c]]), c) for c in stall_cols), reverse=True)[:2]
    print(f"{r[ix['Address']][-5:]} {r[ix['Source']][:62]:62s} smp {int(r[ix['# Samples']]):7d} exec {int(r[ix['Instructions Executed']]):10d} thr {r[ix['Avg. Threads Executed']]:>4s} " +
          " ".join(f"{c[6:]}={n}" for n, c in top))
