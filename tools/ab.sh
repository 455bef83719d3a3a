#!/bin/bash
# A/B: the bench workloads against each library variant built by tools/build_variants.sh
#   tools/ab.sh <outdir> "<variants>" "<workloads>"
out=$1; variants=$2; wls=$3
mkdir -p $out
for v in $variants; do
  echo "=== variant $v"
  BNG_B200_LIB=$(pwd)/bng_b200/variants/libbng_$v.so tools/bench_all.sh $out/$v $wls 2>&1 | grep -v "^$"
done
