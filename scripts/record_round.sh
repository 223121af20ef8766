#!/bin/bash
# records the round's single-GPU measurements into gpurun_out/ (copied to profiles/ afterwards)
# usage (GPU box): PTB200_COMMIT=<sha> bash scripts/record_round.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${R}_env.txt; nproc >> $O/${R}_env.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_c4_reference.json 2> $O/${R}_bench_c4_reference.err
python bench.py > $O/${R}_bench_c4.json 2> $O/${R}_bench_c4.err
python bench.py --workload c2 --no-other-configs > $O/${R}_bench_c2.json 2>/dev/null
for w in c1 c3; do python bench.py --workload $w --steps 3 --warmup 3 --no-other-configs > $O/${R}_bench_$w.json 2>/dev/null; done
python bench.py --workload c4_10m --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $O/${R}_bench_c4_10m.json 2>/dev/null
python bench.py --env-nee --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > $O/${R}_bench_c4_env_nee.json 2>/dev/null
# launch list of one timed step of the default command (C4 at 512 spp, 32 Mi-path pool: 58 iterations x 5 kernels + the
# drain = ~295 launches per step; the scene commit and three warm-up steps are skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -s 940 -c 296 --csv --log-file $O/${R}_launches_c4.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
python scripts/summarize_launches.py $O/${R}_launches_c4.csv > $O/${R}_launches_c4.md
# full captures of the dominant kernel (pool-full launch) of the current build -> ncu_traffic.json
python scripts/ncu_traffic.py capture c4 c2 > $O/${R}_ncu_traffic.log 2>&1
# the shading side of C4: Lambertian shade kernel and the miss kernel on a pool-full iteration
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_shade|k_miss" -s 4 -c 2 -f \
  -o $O/${R}_prof_c4_shade python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
python scripts/ncu_summary.py $O/${R}_prof_c4_shade.ncu-rep > $O/${R}_prof_c4_shade.md 2>/dev/null
python scripts/ncu_summary.py $O/ncu_c4.ncu-rep > $O/${R}_prof_c4_trace.md 2>/dev/null
python scripts/ncu_summary.py $O/ncu_c2.ncu-rep > $O/${R}_prof_c2_trace.md 2>/dev/null
ls -la $O | tail -30
for f in $O/${R}_bench_*.json; do echo $f; cut -c1-260 $f | tail -1; done
cat $O/${R}_launches_c4.md | head -14
cat $O/ncu_traffic.json | head -60
