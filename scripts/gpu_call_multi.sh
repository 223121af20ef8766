#!/bin/bash
# multi-GPU call: the ABI-level multi-GPU tests and the strong-scaling bench at N = number of visible GPUs
# usage (through gpurun --gpus N): bash scripts/gpu_call_multi.sh N [workload] [tag]
N=${1:-2}; W=${2:-c4}; TAG=${3:-r02}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/${TAG}_multi${N}_env.txt
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -rfEs --timeout 900 -p no:cacheprovider > $O/${TAG}_multi${N}_tests.txt 2>&1
echo "pytest rc $?" >> $O/${TAG}_multi${N}_tests.txt
tail -15 $O/${TAG}_multi${N}_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
  bench.py --gpus $N --steps 3 --warmup 3 --workload $W > $O/${TAG}_bench_${W}_${N}gpu.json 2> $O/${TAG}_bench_${W}_${N}gpu.err
echo "bench rc $?"; tail -3 $O/${TAG}_bench_${W}_${N}gpu.err
python - $O/${TAG}_bench_${W}_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('N=%d value %.1f ms/step %.2f reduce %.3f ms rank ms %s e2e %.1f scaling %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['reduce_ms_per_step'], d['rank_ms_per_step'], d['e2e']['value'], d['scaling']))
except Exception as e: print('FAILED', e)
PY
