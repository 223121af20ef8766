#!/bin/bash
export PTB200_NO_FACADE=1
for v in default mb8 mb10 mb12; do
  if [ $v = default ]; then unset PTB200_LIB; else export PTB200_LIB=$PWD/variants/libptb200_$v.so; fi
  echo "== variant $v"; scripts/quick_bench.sh 2>&1 | cut -c1-150
done
