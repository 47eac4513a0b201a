#!/usr/bin/env python3
"""Turns an .ncu-rep of the trace kernel (one launch, `ncu --set full --clock-control none`) into the small JSON that
bench.py quotes in its `roofline` object, STAMPED with the sha256 of the kernel sources it was captured on: bench.py drops
the quote the day csrc/ changes.  No GPU needed.

usage: python tools/ncu_to_json.py gpurun_out/prof.ncu-rep profiles/r2_trace_ncu.json "<kernel name>" "<what was run>"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_sources_sha  # noqa: E402


def main():
    rep, out, kernel, what = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    pick = [r for r in rows[2:] if kernel.split("<")[0] in ",".join(r)] or rows[2:]
    vals = pick[0]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name, scale_unit=True):
        v = float(vals[col[name]].replace(",", ""))
        u = units[col[name]]
        if scale_unit:
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}.get(u, 1.0)
        return v

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    j = {"kernel": kernel, "what": what, "kernel_sources_sha256": kernel_sources_sha(),
         "traffic_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
         "fp32_pipe_active_pct": get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
         "alu_pipe_active_pct": get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", False),
         "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
         "branch_targets_uniform_pct": get("smsp__sass_average_branch_targets_threads_uniform.pct", False),
         "achieved_warps_per_sm": get("sm__warps_active.avg.pct_of_peak_sustained_active", False) * 64 / 100,
         "threads_per_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio", False),
         "registers_per_thread": get("launch__registers_per_thread", False),
         "gpu_time_ms_under_ncu": get("gpu__time_duration.sum"),
         "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}"}
    json.dump(j, open(out, "w"), indent=1)
    print(json.dumps(j, indent=1))


if __name__ == "__main__":
    main()
