#!/bin/bash
# dev: k_trace_deep (two node steps per vote round on >= 64 Ki primitives) on/off, C4, then the GPU suite
export PTB200_NO_FACADE=1
one() { timeout 100 python bench.py --workload c4 --spp 16 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('c4 $1', '%.1f Msamples/s | trace %.4f ms/launch | nodes %.2f' % (d['value'], r['launch_ms'], r['nodes_per_ray']))
"; }
one deep; PTB200_DEEP_MIN=0 one off
timeout 100 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
