#!/bin/bash
# One full-set ncu capture of one kernel of one workload:  tools/ncu_one.sh <out-prefix> <workload> <kernel-regex> [skip]
out=$1; wl=$2; k=$3; skip=${4:-6}
mkdir -p "$(dirname "$out")"
ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -o "$out" -f \
    python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-extra --e2e-steps 1 > "$out.log" 2>&1
python profiles/ncu_summary.py "$out.ncu-rep" > "$out.txt" 2>&1
