#!/bin/bash
# Captures the profiling artefacts of one round on the B200 box (run under gpurun from the repo root):
#   gpurun --timeout 1800 -- 'bash profiles/capture.sh r01'
# Outputs land in gpurun_out/ (scratch); the summaries worth judging are then copied into profiles/.
set -u
R=${1:-r01}
mkdir -p gpurun_out
# 1. launch list: every kernel of a short default bench run with its device time (cold-cache, serialised:
#    compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s 60 -c 120 --csv \
    --log-file gpurun_out/${R}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 \
    > gpurun_out/${R}_launches.log 2>&1
# 2. full-set capture of the dominant kernel (classify) and of the ordered phase (one instance each)
ncu --set full --clock-control none --import-source on -k regex:"k_pipe_classify" -s 5 -c 1 -o gpurun_out/${R}_classify -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${R}_classify.log 2>&1
ncu --set full --clock-control none -k regex:"k_resolve|k_rs_scatter|k_heads" -s 40 -c 4 -o gpurun_out/${R}_group_resolve -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${R}_group_resolve.log 2>&1
# 3. the other programs' kernels
ncu --set full --clock-control none -k regex:"k_dhcp_fastpath" -s 4 -c 1 -o gpurun_out/${R}_dhcp -f \
    python bench.py --workload dhcp --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${R}_dhcp.log 2>&1
ncu --set full --clock-control none -k regex:"k_antispoof" -s 4 -c 1 -o gpurun_out/${R}_antispoof -f \
    python bench.py --workload antispoof_64 --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${R}_antispoof.log 2>&1
# 4. clocks during a plain (unprofiled) run, next to the number itself
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 200 > gpurun_out/${R}_clocks.csv &
SMI=$!
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
kill $SMI
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
for w in pipeline_64 antispoof_64 nat_steady_64 nat_cold_64 nat_ingress_64 qos_64 qos_egress_64 dhcp; do
    python bench.py --workload $w --steps 10 > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err
done
echo done
