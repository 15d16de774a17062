#!/bin/bash
# usage: tools/build_variant.sh <name> [nvcc flags...]  -> bear_b200/_variants/libbear_<name>.so (for A/B runs with BEAR_B200_LIB)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p bear_b200/_variants
S=bear_b200/csrc
SRCS=$(python -c "from bear_b200 import build; print(' '.join('$S/' + s for s in build.SOURCES))")
pids=""; objs=""
for f in $SRCS; do
  o=/tmp/variant_${name}_$(basename $f).o
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --std=c++17 -Xcompiler -fPIC -I include -I $S "$@" -c $f -o $o &
  pids="$pids $!"; objs="$objs $o"
done
for p in $pids; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o bear_b200/_variants/libbear_$name.so $objs
rm -f $objs
echo built bear_b200/_variants/libbear_$name.so
