#!/usr/bin/env python
"""Reads gpurun_out/prof_<tag>.ncu-rep and launches_<tag>.csv (produced by profiles/run_ncu.sh on the GPU box) and
writes the committed summaries profiles/<tag>_kernels.md, profiles/<tag>_launches.md and profiles/traffic.json."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def raw_page(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return hdr, rows[1], rows[2:]


def main():
    lines = [f"# ncu --set full summary ({tag}) — `bash profiles/run_ncu.sh {tag}` under gpurun, 1x B200", ""]
    traffic = {}
    if os.path.exists(rep):
        hdr, units, rows = raw_page(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows:
            name = r[idx["Kernel Name"]] if "Kernel Name" in idx else r[4]
            short = name.split("<")[0].replace("void ", "")
            lines.append(f"## {name[:160]}")
            lines.append("")
            lines.append("| metric | value |")
            lines.append("|---|---|")
            for k in KEYS:
                if k in idx:
                    lines.append(f"| {k} | {r[idx[k]]} {units[idx[k]]} |")
            lines.append("")
            try:
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_read.sum"]]]
                wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_write.sum"]]]
                traffic.setdefault(short, []).append(rd + wr)
            except Exception:
                pass
        # units row lives in rows[1] of the csv; record it for the reader
        with open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("\n".join(lines))
    lpath = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if os.path.exists(lpath):
        txt = open(lpath).read()
        start = txt.find('"ID"')
        rows = list(csv.DictReader(io.StringIO(txt[start:])))
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows:
            try:
                v = float(r["Metric Value"].replace(",", ""))
            except Exception:
                continue
            unit = r.get("Metric Unit", "ns")
            if unit in ("usecond", "us"):
                v *= 1e3
            elif unit in ("msecond", "ms"):
                v *= 1e6
            nm = r["Kernel Name"].split("(")[0][:90]
            agg[nm][0] += 1
            agg[nm][1] += v
        tot = sum(v[1] for v in agg.values())
        out = [f"# launch list ({tag}): every kernel of `bench.py --steps 2 --warmup 3 --no-graph` (whole process, incl. set-up) "
               f"with its device time under ncu (cold-cache, serialised — compare SHARES)", "",
               "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for nm, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            out.append(f"| {nm} | {cnt} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |")
        with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as f:
            f.write("\n".join(out) + "\n")
        print("\n".join(out[:30]))
    if traffic:
        tj = {k: sum(v) / len(v) for k, v in traffic.items()}
        tj["_note"] = f"bytes: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full ({tag}), cfg5 full size"
        json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
