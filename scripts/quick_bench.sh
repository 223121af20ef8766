#!/bin/bash
# dev: short benches of the main workloads, one summary line each
for w in "c2 64" "c4 16" "c1 64" "c3 32"; do set -- $w
  timeout 300 python bench.py --workload $1 --spp $2 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; ks=r['kernel_share']
        print(d['config']['workload'][:2], '%.1f Msamples/s %.1f Mrays/s | trace %.4f ms/launch | share trace %.2f shadow %.2f shade %.2f gen %.2f | nodes %.2f leaves %.2f per ray | e2e %.1f | build %.1f ms sah %.1f->%.1f' % (d['value'], d['mrays_per_s'], r['launch_ms'], ks['trace_closest'], ks['trace_shadow'], ks['shade'], ks['generate'], r['nodes_per_ray'], r['leaves_per_ray'], d['e2e']['value'], d['bvh']['build_ms'], d['bvh']['sah_lbvh'], d['bvh']['sah_refined']))
"; done
