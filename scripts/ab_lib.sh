#!/bin/bash
# A/B build: a second copy of the library with extra -D flags, for same-box comparisons of a kernel variant.
#   bash scripts/ab_lib.sh <out.so> -DFLAG [-DFLAG2 ...]      then   FB200_LIB=<out.so> python scripts/bench_....py
set -e
OUT=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
for f in $ROOT/fortuna_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC "$@" \
    -I $ROOT/include -c $f -o $TMP/$(basename $f .cu).o &
done
wait
nvcc -shared -o $OUT $TMP/*.o -lcudart -lcuda
rm -rf $TMP
echo $OUT
