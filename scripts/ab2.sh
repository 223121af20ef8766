#!/bin/bash
export PTB200_NO_FACADE=1
for v in A B cur; do
  if [ $v = cur ]; then unset PTB200_LIB; else export PTB200_LIB=$PWD/variants/libptb200_$v.so; fi
  python bench.py --workload c1 --spp 64 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', '%.1f Msamples/s ms/step %.2f' % (d['value'], d['ms_per_step']))
"; done
