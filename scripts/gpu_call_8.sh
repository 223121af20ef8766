#!/bin/bash
# 8-GPU call: multi-GPU ABI tests, the default strong-scaling bench (C4) and C5 (3840x2160 @ 4096 spp split over the ranks)
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r02_multi${N}_env.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -rfEs --timeout 600 -p no:cacheprovider > $O/r02_multi${N}_tests.txt 2>&1
tail -3 $O/r02_multi${N}_tests.txt
for W in c4 c5; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus $N --steps 3 --warmup 3 --workload $W > $O/r02_bench_${W}_${N}gpu.json 2> $O/r02_bench_${W}_${N}gpu.err
  echo "bench $W rc $?"
  python - $O/r02_bench_${W}_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('N=%d %s value %.1f ms/step %.2f reduce %.3f ms rank ms %s e2e %.1f scaling %s' % (d['n_gpus'], d['config']['workload'][:2], d['value'], d['ms_per_step'], d['reduce_ms_per_step'], d['rank_ms_per_step'], d['e2e']['value'], d['scaling']))
except Exception as e: print('FAILED', e)
PY
done
# the CLI over 8 devices: one process, pt_renderer_create_multi
( time optix-path-tracer_b200/csrc/ptb200_render --scene 6 --width 1920 --height 1080 --samples 512 --gpus $N --file $O/r02_cli_mesh_${N}gpu ) > $O/r02_cli_${N}gpu.txt 2>&1
tail -6 $O/r02_cli_${N}gpu.txt
( time optix-path-tracer_b200/csrc/ptb200_render --scene 6 --width 1920 --height 1080 --samples 512 --gpus 1 --file $O/r02_cli_mesh_1gpu ) > $O/r02_cli_1gpu.txt 2>&1
tail -6 $O/r02_cli_1gpu.txt
python - <<'PY'
import numpy as np
from PIL import Image
a=np.asarray(Image.open('gpurun_out/r02_cli_mesh_8gpu.png')).astype(int); b=np.asarray(Image.open('gpurun_out/r02_cli_mesh_1gpu.png')).astype(int)
print('CLI 8-GPU PNG vs 1-GPU PNG: max |diff| =', np.abs(a-b).max(), 'of 255; shape', a.shape)
PY
rm -f $O/r02_cli_mesh_1gpu.png
