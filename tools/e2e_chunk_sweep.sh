#!/bin/bash
# End-to-end throughput of the default workload against the pipeline chunk size of the zero-copy path and the
# grid of its gather / scatter kernels.
#   gpurun -- 'bash tools/e2e_chunk_sweep.sh'
mkdir -p gpurun_out/r02
for bps in ${BPS:-1 2 8}; do
for lg in ${LGS:-16 17 18 19}; do
    BNG_ZC_BLOCKS_PER_SM=$bps BNG_ZC_CHUNK_LOG2=$lg timeout -s KILL 100 python bench.py --steps 3 --no-cpu --no-extra --e2e-steps 8 2> gpurun_out/r02/chunk_$lg.err |
        python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('blocks/SM $bps chunk 2^$lg: e2e', j['e2e']['value'], 'header-split', (j.get('e2e_header_split') or {}).get('value'))"
done
done | tee gpurun_out/r02/e2e_chunk_sweep.txt
