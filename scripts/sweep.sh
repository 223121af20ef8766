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
export PTB200_NO_FACADE=1
for wl in "c2 64" "c4 16"; do set -- $wl
  for knobs in "8 12" "4 12" "12 12" "8 8" "8 16" "8 20" "6 16" "12 16"; do set -- $wl $knobs
    PTB200_LEAF_THRESH=$3 PTB200_REFILL_MIN=$4 python bench.py --workload $1 --spp $2 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | ex "$1 leaf=$3 refill=$4"
  done
  for pool in 2097152 4194304 8388608; do
    python bench.py --workload $1 --spp $2 --steps 3 --warmup 2 --no-cpu-baseline --pool $pool 2>/dev/null | ex "$1 pool=$pool"
  done
done
