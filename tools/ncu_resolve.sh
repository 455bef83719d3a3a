#!/bin/bash
# Full-set ncu capture of one steady-state k_resolve launch:  tools/ncu_resolve.sh <out-prefix> [bench args...]
out=$1; shift
mkdir -p "$(dirname "$out")"
ncu --set full --clock-control none --import-source on -k regex:"k_resolve" -s 3 -c 1 -o "$out" -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --e2e-steps 0 "$@" > "$out.log" 2>&1
python profiles/ncu_summary.py "$out.ncu-rep" > "$out.txt" 2>&1
