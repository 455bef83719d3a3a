#!/bin/bash
# Throughput and per-batch latency of the default workload against the batch size.
#   gpurun --timeout 600 -- 'bash tools/batch_sweep.sh'
mkdir -p gpurun_out/r02
for lg in 12 14 16 18 20 22; do
    timeout -s KILL 120 python bench.py --frames $((1 << lg)) --steps 30 --no-cpu --no-extra --e2e-steps 2 2> gpurun_out/r02/sweep_$lg.err |
        python -c "import sys,json; j=json.loads(sys.stdin.readline()); print(json.dumps({'frames': j['config']['frames_per_gpu_per_step'], 'Mpps': j['value'], 'ms_per_batch': j['ms_per_step'], 'e2e_Mpps': j['e2e']['value'], 'kernels_ms': j['roofline']['kernels_ms'], 'launches_per_batch': j['gpu_launches'] / j['steps']}))"
done | tee gpurun_out/r02/batch_sweep.jsonl
