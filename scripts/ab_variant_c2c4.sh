#!/bin/bash
# dev: A/B of variants/libptb200_<name>.so against the in-tree library on C2 and C4 (two passes each), then the
# hit-id parity tests with the variant loaded.  usage (on the GPU box): scripts/ab_variant_c2c4.sh hit8
V=${1:?variant name}
export PTB200_NO_FACADE=1
one() { timeout 200 python bench.py --workload $1 --spp $2 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$1', '%.1f Msamples/s | trace %.4f ms/launch | nodes %.2f' % (d['value'], r['launch_ms'], r['nodes_per_ray']))
"; }
for v in default $V default $V; do
  if [ $v = default ]; then unset PTB200_LIB; else export PTB200_LIB=$PWD/variants/libptb200_$v.so; fi
  echo "== $v"; one c2 64; one c4 16
done
export PTB200_LIB=$PWD/variants/libptb200_$V.so
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "hit_ids or edge or soup" 2>&1 | tail -2
