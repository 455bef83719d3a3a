#!/bin/bash
# A/B builds of libbng_b200.so with different compile-time knobs (selected at run time with BNG_B200_LIB=...):
#   tools/build_variants.sh name1 "-DFOO=1" name2 "-DBAR=2" ...   -> bng_b200/variants/libbng_<name>.so
cd "$(dirname "$0")/../bng_b200/csrc" || exit 1
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( make -s -j4 BUILD=build/v_$name OUT=$(pwd)/../variants/libbng_$name.so EXTRA="$flags" && echo "built $name ($flags)" ) &
done
wait
