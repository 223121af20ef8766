#!/bin/bash
# dev: A/B of whole library builds (variants/libptb200_<name>.so, loaded through PTB200_LIB) on short C4 / C2 runs
# usage: bash scripts/gpu_ab.sh tag name1 name2 ...   ("default" = the in-tree library)
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v = default ]; then unset PTB200_LIB; else export PTB200_LIB=$PWD/variants/libptb200_$v.so; fi
  for w in "c4 64" "c2 128" "c1 64" "c3 64"; do set -- $w
    timeout 300 python bench.py --workload $1 --spp $2 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${TAG}_${v}_$1.json 2> gpurun_out/${TAG}_${v}_$1.err
    python - gpurun_out/${TAG}_${v}_$1.json $v $1 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']; inc=d['incoherent_secondary']
    print('%-10s %-3s value %7.1f e2e %7.1f incoh %s nodes/ray %.2f leaves/ray %.2f trace launch %.4f ms share %.3f build %.1f ms' % (sys.argv[2], sys.argv[3], d['value'], d['e2e']['value'], ('%.1f' % inc['mrays_per_s']) if inc else '-', r['nodes_per_ray'], r['leaves_per_ray'], r['launch_ms'], r['kernel_share']['trace_closest'], d['bvh']['build_ms']))
except Exception as e:
    print(sys.argv[2], sys.argv[3], 'FAILED', e)
PY
  done
done
