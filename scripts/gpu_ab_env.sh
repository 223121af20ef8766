#!/bin/bash
# dev: A/B of environment switches on short runs.  usage: bash scripts/gpu_ab_env.sh tag "workload spp" name=ENV=VAL[,ENV=VAL] ...
TAG=$1; WL=$2; shift; shift
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}; [ "$envs" = "$spec" ] && envs="A=1"
  for w in $WL; do
    wl=${w%%:*}; spp=${w#*:}
    env $(echo $envs | tr ',' ' ') timeout 300 python bench.py --workload $wl --spp $spp --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG}_${name}_$wl.json 2> gpurun_out/${TAG}_${name}_$wl.err
    python - gpurun_out/${TAG}_${name}_$wl.json $name $wl <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']; inc=d['incoherent_secondary']
    print('%-14s %-6s value %7.1f e2e %7.1f incoh %s nodes/ray %.2f trace launch %.4f ms share %.3f build %.1f ms' % (sys.argv[2], sys.argv[3], d['value'], d['e2e']['value'], ('%.1f' % inc['mrays_per_s']) if inc else '-', r['nodes_per_ray'], r['launch_ms'], r['kernel_share']['trace_closest'], d['bvh']['build_ms']))
except Exception as e:
    print(sys.argv[2], sys.argv[3], 'FAILED', e)
PY
  done
done
