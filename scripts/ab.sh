#!/bin/bash
# A/B: current library vs a variant, C1 and C3, python-built scenes
export PTB200_NO_FACADE=1
for v in "" "$PWD/variants/libptb200_mb8.so"; do
  export PTB200_LIB=$v; [ -z "$v" ] && unset PTB200_LIB
  echo "== lib ${v:-current}"
  for w in "c1 64" "c3 32" "c2 32"; do set -- $w
   python bench.py --workload $1 --spp $2 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; ks=r['kernel_share']
        print(d['config']['workload'][:2], '%.1f Msamples/s | ms/step %.2f launches %d | trace %.4f ms/launch | share trace %.2f shadow %.2f shade %.2f gen %.2f' % (d['value'], d['ms_per_step'], d['gpu_launches'], r['launch_ms'], ks['trace_closest'], ks['trace_shadow'], ks['shade'], ks['generate']))
"; done; done
