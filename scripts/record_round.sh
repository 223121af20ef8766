#!/bin/bash
# records the round's measurements into gpurun_out/ (copied to profiles/ afterwards)
R=${1:-r01}
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_c2_reference.json 2>/dev/null
python bench.py > gpurun_out/${R}_bench_c2.json 2>/dev/null
for w in c1 c3 c4; do python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${R}_bench_$w.json 2>/dev/null; done
python bench.py --workload c4_10m --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_c4_10m.json 2>/dev/null
# launch list of the default bench command (first 800 launches after warm-up launches)
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 800 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline > /dev/null 2>&1
# full captures of the dominant kernel on large launches
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -o gpurun_out/${R}_prof_c2_trace python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -o gpurun_out/${R}_prof_c4_trace python bench.py --workload c4 --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_shade<\(int\)0, \(int\)0>|k_shadow" -s 6 -c 2 -o gpurun_out/${R}_prof_c1_shade python bench.py --workload c1 --steps 1 --warmup 1 --spp 32 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -12
for f in gpurun_out/${R}_bench_*.json; do echo $f; cut -c1-330 $f | tail -1; done
