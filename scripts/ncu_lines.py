"""Dev aid: attribute an ncu SASS-page CSV (per-instruction executed counts / stall samples) to CUDA source lines
by zipping it, in order, with `nvdisasm -g` output of the same kernel from the built cubin.

usage: python scripts/ncu_lines.py <ncu_sass.csv> <cubin> <kernel-name-substring> [top]
"""
import csv, re, subprocess, sys, collections

sass_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
# ncu emits each row twice in this CSV flavour: de-duplicate on address
seen, uniq = set(), []
for r in data:
    a = r[ix["Address"]]
    if a in seen:
        continue
    seen.add(a)
    uniq.append(r)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the function section
start = None
for i, l in enumerate(dis):
    if l.startswith(".text.") and kname in l:
        start = i
        break
    if re.match(r"\s*\.section\s+\.text\.", l) and kname in l:
        start = i
        break
assert start is not None, "kernel not found in cubin"
cur = ("?", 0)
ins = []
for l in dis[start + 1:]:
    if re.match(r"\s*\.section\s", l) and ins:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((cur, m.group(2)))
print(f"ncu rows {len(uniq)}  disasm instrs {len(ins)}")
n = min(len(uniq), len(ins))
agg = collections.defaultdict(lambda: [0, 0, 0])
def I(r, h):
    try:
        return int(r[ix[h]])
    except Exception:
        return 0
for r, (loc, txt) in zip(uniq[:n], ins[:n]):
    a = agg[loc]
    a[0] += I(r, "Instructions Executed")
    a[1] += I(r, "# Samples")
    a[2] += 1
tot = sum(a[0] for a in agg.values()) or 1
smp = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instructions {tot}, samples {smp}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]:9d} {100*a[0]/tot:5.1f}%  samples {a[1]:5d} {100*a[1]/smp:5.1f}%  sass {a[2]:4d}  {loc[0]}:{loc[1]}")
# optional 5th argument file:line — list that line's SASS with execution counts
if len(sys.argv) > 5:
    f, ln = sys.argv[5].rsplit(":", 1)
    print(f"--- SASS attributed to {f}:{ln}")
    for r, (loc, txt) in zip(uniq[:n], ins[:n]):
        if loc == (f, int(ln)):
            print(f"{I(r, 'Instructions Executed'):9d}  smp {I(r, '# Samples'):4d}  {txt}")
