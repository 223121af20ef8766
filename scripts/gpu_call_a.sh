#!/bin/bash
# round 2, first GPU call: the whole GPU suite, the default bench line, and short A/Bs of the new switches
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r02a_env.txt; nproc >> $O/r02a_env.txt
timeout 1500 python -m pytest tests -m gpu -q -s -rfE --timeout 900 -p no:cacheprovider > $O/r02a_gputest.txt 2>&1
echo "pytest rc $?" >> $O/r02a_gputest.txt
tail -5 $O/r02a_gputest.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/r02a_bench_c4.json 2> $O/r02a_bench_c4.err
echo "bench rc $?"; head -c 600 $O/r02a_bench_c4.json
short() { # name, env..., -- args
  n=$1; shift
  env "$@" timeout 300 python bench.py --workload c4 --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > $O/r02a_ab_$n.json 2> $O/r02a_ab_$n.err
  python - $O/r02a_ab_$n.json $n <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'value %.1f e2e %.1f incoh %.1f nodes/ray %.2f launch_ms %.4f' % (d['value'], d['e2e']['value'], d['incoherent_secondary']['mrays_per_s'], d['roofline']['nodes_per_ray'], d['roofline']['launch_ms']))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
short base A=1
short nodeeprays PTB200_DEEP_RAYS=0
short nodrain PTB200_DRAIN_MAX=0
for spp in 32 256; do
  timeout 300 python bench.py --workload c2 --spp $spp --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > $O/r02a_c2_spp$spp.json 2> $O/r02a_c2_spp$spp.err
  PTB200_DRAIN_MAX=0 timeout 300 python bench.py --workload c2 --spp $spp --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > $O/r02a_c2_spp${spp}_nodrain.json 2> $O/r02a_c2_spp${spp}_nodrain.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02a_c2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value %.1f ms %.3f' % (d['value'], d['ms_per_step']))
    except Exception as e: print(f, 'FAILED', e)
PY
