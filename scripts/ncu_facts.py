"""Turn ncu --set full reports (.ncu-rep) into profiles/r2_ncu_facts.json: per kernel the per-launch DRAM traffic and the
issue statistics bench.py quotes in `roofline.traffic` / `roofline.ncu`.

usage: python scripts/ncu_facts.py key=report.ncu-rep[:launch_index] ...        (merges into the existing JSON)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("NCU_FACTS_OUT", os.path.join(ROOT, "profiles", "r2_ncu_facts.json"))
WANT = {
    "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write", "gpu__time_duration.sum": "ncu_duration_us",
    "smsp__inst_executed.sum": "warp_instructions", "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def facts(rep, launch=0):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2 + launch]
    out = {"report": os.path.basename(rep), "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            x = float(v.replace(",", ""))
            if h.startswith("dram__bytes"):
                x *= UNIT.get(u, 1.0)
            if h == "gpu__time_duration.sum":
                x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            out[WANT[h]] = x
    return out


def main():
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for a in sys.argv[1:]:
        key, rep = a.split("=", 1)
        launch = 0
        if ":" in rep:
            rep, li = rsplit = rep.rsplit(":", 1)
            launch = int(li)
        data[key] = facts(rep, launch)
        print(key, json.dumps(data[key]))
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
