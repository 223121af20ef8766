#!/bin/bash
ex() { python -c '
import json,sys
tag=sys.argv[1]
for l in sys.stdin:
    if l.startswith("{\"metric\""):
        d=json.loads(l); r=d["roofline"]
        print("%-36s %8.1f Msamples/s %8.1f Mrays/s trace %.4f ms share %.2f" % (tag, d["value"], d["mrays_per_s"], r["launch_ms"], r["kernel_share"]["trace_closest"]), flush=True)
' "$1"; }
for wl in "c2 64" "c4 16" "c1 64"; do set -- $wl
  for lt in 1 4 8; do for rf in 8 16 24 32; do
    PTB200_LEAF_THRESH=$lt PTB200_REFILL_MIN=$rf python bench.py --workload $1 --spp $2 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | ex "$1 leaf=$lt refill=$rf"
  done; done
done
