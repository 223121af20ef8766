#!/bin/bash
# dev sweep of scheduler knobs; prints "<knobs> <workload> value mrays trace_launch_ms"
ex() { python -c '
import json,sys
tag=sys.argv[1]
for l in sys.stdin:
    if l.startswith("{\"metric\""):
        d=json.loads(l); r=d["roofline"]
        print("%-36s %8.1f Msamples/s %8.1f Mrays/s trace %.4f ms nodes/ray %.2f leaves/ray %.2f" % (tag, d["value"], d["mrays_per_s"], r["launch_ms"], r["nodes_per_ray"], r["leaves_per_ray"]), flush=True)
' "$1"; }
for wl in "c2 64" "c4 16"; do set -- $wl
  for lt in 1 4 8 16 32; do for rf in 1 6 16; do
    PTB200_LEAF_THRESH=$lt PTB200_REFILL_MIN=$rf python bench.py --workload $1 --spp $2 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | ex "$1 leaf=$lt refill=$rf"
  done; done
  for pool in 524288 2097152 4194304 8388608; do
    python bench.py --workload $1 --spp $2 --steps 2 --warmup 1 --no-cpu-baseline --pool $pool 2>/dev/null | ex "$1 pool=$pool"
  done
done
