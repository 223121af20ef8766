#!/bin/bash
# dev: refill threshold x pool size at (near) full sizes
ex() { python -c '
import json,sys
tag=sys.argv[1]
for l in sys.stdin:
    if l.startswith("{\"metric\""):
        d=json.loads(l); r=d["roofline"]
        print("%-36s %8.1f Msamples/s %8.1f Mrays/s trace %.4f ms" % (tag, d["value"], d["mrays_per_s"], r["launch_ms"]), flush=True)
' "$1"; }
export PTB200_NO_FACADE=1
run() { PTB200_REFILL_MIN=$3 python bench.py --workload $1 --spp $2 --steps 3 --warmup 2 --no-cpu-baseline --pool $4 2>/dev/null | ex "$1 spp=$2 refill=$3 pool=$4"; }
for rf in 16 20 24 28; do run c2 256 $rf 8388608; run c4 128 $rf 8388608; run c1 64 $rf 8388608; run c3 128 $rf 8388608; done
