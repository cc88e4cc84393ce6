"""Condense an ncu launch list of scripts/probe_streams.py (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per
launch) into (a) profiles/<tag>_streams_launches.txt — the kernels of ONE build / voxel-grid / Add_Points call with their time, DRAM
bytes and achieved DRAM rate — and (b) the stream_* entries of profiles/<tag>_ncu_facts.json that bench.py's `streams` record reads.
usage: python scripts/stream_facts.py <tag> <launches.csv>"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, path = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path, errors="ignore")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {n: i for i, n in enumerate(hdr)}
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = L.setdefault(int(r[ix["ID"]]), {"k": r[ix["Kernel Name"]].split("(")[0].replace("icp4r::", "").replace("<unnamed>::", ""), "grid": r[ix["Grid Size"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * mul.get(r[ix["Metric Unit"]], 1)
seq = list(L.values())


def last_run(first, last):
    """the complete launch sequence first ... last that moved the most bytes (the probe also builds smaller maps)"""
    best, best_b = [], -1.0
    ends = [i for i, d in enumerate(seq) if d["k"] == last]
    for e in ends:
        for b in range(e, -1, -1):
            if seq[b]["k"] == first:
                run = seq[b:e + 1]
                tot = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in run)
                if tot >= best_b:
                    best, best_b = run, tot
                break
    return best


out_txt = [f"# one call each, from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python scripts/probe_streams.py`",
           "# (per-launch times under ncu are serialised and cold-cache; bench values are never taken under a profiler)", ""]
facts = {}
for key, title, first, last in (("stream_build", "map_build of the 20 M-point dense map", "bbox_init", "coarse_count_kernel"),
                                ("stream_voxel_grid", "icp4r_voxel_grid, 20 M points, 0.5 m leaves", "vg_minmax_init", "vg_leaf_kernel"),
                                ("stream_add_points", "Add_Points(3,000 points, false) into a 3 M-point map", "inc_info_init", "inc_cell_kernel")):
    run = last_run(first, last)
    if not run:
        continue
    out_txt.append(f"## {title}")
    tot_t = tot_b = 0.0
    ks = collections.OrderedDict()
    for d in run:
        t = d.get("gpu__time_duration.sum", 0.0)
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        tot_t += t
        tot_b += b
        ks[d["k"]] = ks.get(d["k"], 0.0) + t
        out_txt.append(f"  {d['k']:28s} grid {d['grid']:>14s} {t:9.1f} us   read {d.get('dram__bytes_read.sum', 0) / 1e6:9.1f} MB  written {d.get('dram__bytes_write.sum', 0) / 1e6:9.1f} MB"
                       f"  {b / max(t, 1e-9) / 1e3:8.1f} GB/s")
    out_txt.append(f"  total {tot_t:9.1f} us, {tot_b / 1e6:.1f} MB of DRAM traffic")
    out_txt.append("")
    facts[key] = {"dram_bytes": tot_b, "kernel_time_us": tot_t, "kernels_us": {k: round(v, 1) for k, v in ks.items()}}
open(os.path.join(ROOT, "profiles", f"{tag}_streams_launches.txt"), "w").write("\n".join(out_txt))
fp = os.path.join(ROOT, "profiles", f"{tag}_ncu_facts.json")
allf = json.load(open(fp)) if os.path.exists(fp) else {}
allf.update(facts)
json.dump(allf, open(fp, "w"), indent=1, sort_keys=True)
print("\n".join(out_txt))
