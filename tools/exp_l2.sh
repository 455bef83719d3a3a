#!/bin/bash
# A/B harness for classify-kernel experiments: runs the pipeline workloads against alternative builds of
# the library (bng_b200/libbng_b200_<tag>.so, selected with BNG_B200_LIB) and prints the per-kernel times.
#   gpurun --timeout 900 -- 'bash tools/exp_l2.sh exp1 exp2 ...'
set -u
mkdir -p gpurun_out
run() { # tag, env...
    local tag=$1; shift
    for w in ${WORKLOADS:-pipeline_64}; do
        env "$@" python bench.py --workload $w --steps 10 --no-cpu --e2e-steps 1 2> gpurun_out/exp_$tag.$w.err |
            python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('$tag', '$w', j['value'], j['ms_per_step'], j['roofline'].get('kernels_ms'))"
    done
}
run base X=1
run base_f8 BNG_FLOWS_PER_SUB=8
for t in "$@"; do
    run $t BNG_B200_LIB=$PWD/bng_b200/libbng_b200_$t.so
    run ${t}_f8 BNG_B200_LIB=$PWD/bng_b200/libbng_b200_$t.so BNG_FLOWS_PER_SUB=8
done
