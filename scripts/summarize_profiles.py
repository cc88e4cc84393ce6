"""Turn the raw ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.

usage: python scripts/summarize_profiles.py <round-tag> <launches.csv> <report.ncu-rep> <kernel-substring> <cubin-object-name>
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep, ksub, cubin_name = sys.argv[1:6]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list ----------------------------------------------------------------------------------------------
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    agg[name].append(float(row["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write(f"# source: {os.path.basename(launches)}; total {tot/1e3:.1f} us over {sum(len(v) for v in agg.values())} launches\n")
    f.write(f"{'total_us':>10} {'launches':>8} {'avg_us':>9} {'share':>7}  kernel\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{sum(v)/1e3:10.1f} {len(v):8d} {sum(v)/len(v)/1e3:9.2f} {100*sum(v)/tot:6.1f}%  {k}\n")

# ---- full-set capture of the dominant kernel --------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shared_ld.sum"]
with open(os.path.join(out_dir, f"{tag}_{ksub}_ncu.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on ; report {os.path.basename(rep)} (cold-cache replays)\n")
    for d in rows[2:]:
        for w in want:
            for i, h in enumerate(hdr):
                if h == w:
                    f.write(f"{w:72s} {d[i][:110]} {units[i]}\n")
        # stall reasons > 2 %
        f.write("warp stall reasons (pc-sampling counts, share of all samples, > 2 %):\n")
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(d[i])
                except ValueError:
                    pass
        tot_s = sum(st.values()) or 1.0
        for k2, v in sorted(st.items(), key=lambda kv: -kv[1]):
            if v / tot_s > 0.02:
                f.write(f"    {k2:28s} {100 * v / tot_s:5.1f} %\n")
        f.write("\n")
    # per-source-line attribution of the first captured launch
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::1"], capture_output=True, text=True).stdout
    tmp = os.path.join("/tmp", f"{tag}_src.csv")
    open(tmp, "w").write(src)
    xdir = "/tmp/xelf_sum"
    subprocess.run(f"rm -rf {xdir} && mkdir -p {xdir} && cd {xdir} && cuobjdump -xelf all {ROOT}/icp-4dradar_b200/libicp4r_cuda.so", shell=True,
                   capture_output=True)
    kname = rows[2][hdr.index("Kernel Name")]
    m = re.match(r"void (\w+)<([^>]*)>", kname.replace("icp4r::", ""))
    mangled_hint = ksub
    if m:
        args = [a.strip() for a in m.group(2).split(",")]
        mangled_hint = f"{m.group(1)}I" + "".join(f"Li{a}E" for a in args) + "E"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), tmp, os.path.join(xdir, cubin_name), mangled_hint, "30"],
                       capture_output=True, text=True)
    f.write("executed warp-instructions and stall samples by source line (first captured launch):\n")
    f.write(r.stdout + r.stderr[-500:])
print("written to", out_dir)
