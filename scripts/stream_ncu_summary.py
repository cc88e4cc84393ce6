"""Condense `ncu --set full` raw CSV pages of the streaming kernels (gpurun_out/sf_*_raw.csv) into profiles/<tag>_streams_ncu.txt:
the counters that say what each kernel is bound by. usage: python scripts/stream_ncu_summary.py <tag> <raw.csv> [<raw.csv> ...]"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
out = ["# ncu --set full --clock-control none of the streaming kernels at 20 M points (scripts/probe_streams.py); one launch each", ""]
seen = set()
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path, errors="ignore")))
    if len(rows) < 3:
        continue
    hdr = rows[0]
    units = rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        out.append(f"## {name}")
        for k in KEYS:
            if k in d:
                out.append(f"  {k:72s} {d[k]:>16s} {u.get(k, '')}")
        st = []
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    st.append((float(v), k.split("issue_stalled_")[1].split("_per_")[0]))
                except ValueError:
                    pass
        out.append("  warp stall reasons (stalled warps per issued instruction): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:5]))
        out.append("")
open(os.path.join(ROOT, "profiles", f"{sys.argv[1]}_streams_ncu.txt"), "w").write("\n".join(out))
print("\n".join(out))
