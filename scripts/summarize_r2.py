"""Condense the round-2 ncu captures (gpurun_out/r2_<cfg>_raw.csv / _source.csv, written on the GPU box by
scripts/capture_r2.sh) into the text summaries under profiles/: key counters, stall reasons, and the executed
warp-instructions / stall samples per CUDA source line (SASS rows zipped, in order, with `nvdisasm -g` of the cubin).

usage: python scripts/summarize_r2.py"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = {  # capture -> (object file, kernel name fragment, description)
    "c4": ("register_batch.o", "reg_batch_kernelILi0ELi256", "reg_batch_kernel<P2P_SVD,256>, 2,048 C4 pairs (one launch)"),
    "c1": ("register_batch.o", "reg_batch_kernelILi0ELi256", "reg_batch_kernel<P2P_SVD,256>, ONE C1 pair = one icp4r_register call"),
    "c2batch": ("register_map.o", "reg_iter_kernelILi2ELi5ELi0ELi1", "reg_iter_kernel<P2PLANE_KNN,5,ITER,proof flavour>, 16 C2 scans, iteration ~12"),
    "c2single": ("register_map.o", "reg_iter_kernelILi2ELi5ELi0ELi0", "reg_iter_kernel<P2PLANE_KNN,5,ITER,lean flavour>, one C2 scan, iteration ~10"),
    "c5": ("register_map.o", "reg_iter_kernelILi2ELi5ELi0ELi0", "reg_iter_kernel<P2PLANE_KNN,5,ITER,lean flavour>, C5 (16,384-pt scan, 20 M-pt map), iteration ~10"),
    "c3": ("register_map.o", "reg_iter_kernelILi2ELi5ELi0ELi0", "reg_iter_kernel<P2PLANE_KNN,5,ITER,lean flavour>, C3 last frame vs the final 6.9 M-pt map"),
}
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def lines_of(cubin, frag):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = None
    for i, l in enumerate(dis):
        if re.match(r"\s*\.section\s+\.text\.", l) and frag in l:
            start = i
            break
    if start is None:
        return None
    cur, ins = ("?", 0), []
    for l in dis[start + 1:]:
        if re.match(r"\s*\.section\s", l) and ins:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
            ins.append(cur)
    return ins


def main():
    src_cache = {}
    for cfg, (obj, frag, what) in CFG.items():
        raw = os.path.join(ROOT, "gpurun_out", f"r2_{cfg}_raw.csv")
        srcp = os.path.join(ROOT, "gpurun_out", f"r2_{cfg}_source.csv")
        if not os.path.exists(raw):
            continue
        rows = list(csv.reader(open(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        out = [f"# {what}", f"# ncu --set full --clock-control none --import-source on (scripts/capture_r2.sh); kernel: {vals[hdr.index('Kernel Name')]}", ""]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"{k:70s} {vals[i]:>16s} {units[i]}")
        out.append("")
        out.append("warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active, descending):")
        st = [(float(vals[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
              for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        for v, h in sorted(st, reverse=True)[:8]:
            out.append(f"    {h:28s} {v:6.2f}")
        if os.path.exists(srcp):
            with tempfile.TemporaryDirectory() as td:
                subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "icp-4dradar_b200", "csrc", obj)], cwd=td, capture_output=True)
                cub = [f for f in os.listdir(td) if f.endswith(".cubin")]
                ins = lines_of(os.path.join(td, cub[0]), frag) if cub else None
            srows = list(csv.reader(open(srcp)))
            sh = srows[1]
            ix = {h: i for i, h in enumerate(sh)}
            seen, uniq = set(), []
            for r in srows[2:]:
                if len(r) != len(sh) or r[ix["Address"]] in seen:
                    continue
                seen.add(r[ix["Address"]])
                uniq.append(r)
            if ins and len(ins) == len(uniq):
                agg = collections.defaultdict(lambda: [0, 0])
                for r, loc in zip(uniq, ins):
                    agg[loc][0] += int(r[ix["Instructions Executed"]] or 0)
                    agg[loc][1] += int(r[ix["# Samples"]] or 0)
                ti = sum(a[0] for a in agg.values()) or 1
                ts = sum(a[1] for a in agg.values()) or 1
                out.append("")
                out.append(f"executed warp-instructions {ti}, stall samples {ts}; top source lines by samples (share of samples | share of instructions):")
                for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
                    out.append(f"    {100 * a[1] / ts:5.1f} % | {100 * a[0] / ti:5.1f} %   {loc[0]}:{loc[1]}")
            else:
                out.append("")
                out.append(f"(source attribution skipped: {len(uniq)} SASS rows in the report vs {len(ins) if ins else 0} in the current build)")
        open(os.path.join(ROOT, "profiles", f"r2_{cfg}_ncu.txt"), "w").write("\n".join(out) + "\n")
        print("wrote", f"profiles/r2_{cfg}_ncu.txt")


if __name__ == "__main__":
    main()
