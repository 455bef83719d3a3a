#!/bin/bash
# Subscriber-table sparsity vs classify time (second probes are what most warps wait for).
mkdir -p gpurun_out
for h in 2 4 8 16 32; do
  for w in pipeline_imix antispoof_64 qos_64; do
    BNG_SUBS_HEADROOM=$h timeout -s KILL 100 python bench.py --workload $w --steps 10 --no-cpu --e2e-steps 1 2> gpurun_out/hr_$h.err |
        python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('headroom $h', j['config']['workload'], j['value'], j['ms_per_step'], j['roofline']['kernels_ms'])"
  done
done | tee gpurun_out/headroom_sweep.txt
