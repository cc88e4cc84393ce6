#!/bin/bash
# Round-2 ncu captures of the dominant kernel of every bench config (run on the GPU box through gpurun; every target is
# first run plain). The .ncu-rep files are too large to travel back together: they are condensed on the box into
# gpurun_out/r2_ncu_facts.json (scripts/ncu_facts.py), per-report raw / source CSV pages, and only the two smallest
# reports are kept.
set -e
cd "$(dirname "$0")/.."
O=gpurun_out
for w in c1 c2single c2batch c4 c5 c3; do python scripts/probe_ncu_target.py $w > $O/plain_$w.log 2>&1; done
N="ncu --set full --clock-control none --import-source on"
$N -k regex:reg_batch -s 2 -c 1 -o $O/r2_c1 -f python scripts/probe_ncu_target.py c1 > $O/ncu_c1.log 2>&1
$N -k regex:reg_iter_kernel -s 52 -c 1 -o $O/r2_c2single -f python scripts/probe_ncu_target.py c2single > $O/ncu_c2single.log 2>&1
$N -k regex:reg_iter_kernel -s 33 -c 1 -o $O/r2_c2batch -f python scripts/probe_ncu_target.py c2batch > $O/ncu_c2batch.log 2>&1
$N -k regex:reg_batch -s 1 -c 1 -o $O/r2_c4 -f python scripts/probe_ncu_target.py c4 > $O/ncu_c4.log 2>&1
$N -k regex:reg_iter_kernel -s 52 -c 1 -o $O/r2_c5 -f python scripts/probe_ncu_target.py c5 > $O/ncu_c5.log 2>&1
$N -k regex:reg_iter_kernel -s 42031 -c 1 -o $O/r2_c3 -f python scripts/probe_ncu_target.py c3 > $O/ncu_c3.log 2>&1 || true
export NCU_FACTS_OUT=$O/r2_ncu_facts.json
args=""
for k in c1:reg_batch_kernel_c1 c2single:reg_iter_kernel_c2_single c2batch:reg_iter_kernel_c2_batch16 c4:reg_batch_kernel c5:reg_iter_kernel_c5 c3:reg_iter_kernel_c3; do
  f=${k%%:*}; key=${k##*:}
  if [ -f $O/r2_$f.ncu-rep ]; then
    args="$args $key=$O/r2_$f.ncu-rep"
    ncu -i $O/r2_$f.ncu-rep --page raw --csv > $O/r2_${f}_raw.csv 2>/dev/null
    ncu -i $O/r2_$f.ncu-rep --page source --csv > $O/r2_${f}_source.csv 2>/dev/null
  fi
done
python scripts/ncu_facts.py $args
rm -f $O/r2_c2single.ncu-rep $O/r2_c2batch.ncu-rep $O/r2_c5.ncu-rep $O/r2_c3.ncu-rep
ls -la $O | head -40
