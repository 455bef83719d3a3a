#!/bin/bash
# End-to-end throughput of the default workload against the pipeline chunk size of the zero-copy path.
mkdir -p gpurun_out
for rep in 1 2; do
for lg in 16 17 18 19; do
    BNG_ZC_CHUNK_LOG2=$lg timeout -s KILL 100 python bench.py --steps 5 --no-cpu --e2e-steps 5 2> gpurun_out/chunk_$lg.err |
        python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('chunk 2^$lg rep $rep: e2e', j['e2e']['value'], 'header-split', (j.get('e2e_header_split') or {}).get('value'))"
done
done | tee gpurun_out/e2e_chunk_sweep.txt
