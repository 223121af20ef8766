#!/bin/bash
# dev: quick bench of the in-tree library and of every variants/libptb200_<name>.so named on the command line
export PTB200_NO_FACADE=1
for v in default "$@"; do
  if [ $v = default ]; then unset PTB200_LIB; else export PTB200_LIB=$PWD/variants/libptb200_$v.so; fi
  echo "== variant $v"; scripts/quick_bench.sh 2>&1 | cut -c1-150
done
