#!/bin/bash
# last GPU call of a round: whole GPU suite, the default bench line, and fresh ncu counters of the closest-hit kernel
# of the CURRENT build (ncu_traffic.json), each stage under its own timeout
# usage (GPU box): PTB200_COMMIT=<sha> bash scripts/gpu_call_final.sh r02f
R=${1:-r02f}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${R}_env.txt; nproc >> $O/${R}_env.txt
timeout 200 python -m pytest tests -m gpu -q -s -rfE --timeout 150 -p no:cacheprovider > $O/${R}_gputest.txt 2>&1
echo "pytest rc $?" >> $O/${R}_gputest.txt
tail -3 $O/${R}_gputest.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py > $O/${R}_bench_c4.json 2> $O/${R}_bench_c4.err
echo "bench rc $?"; cut -c1-400 $O/${R}_bench_c4.json | tail -1
timeout 150 python scripts/ncu_traffic.py capture c4 c2 > $O/${R}_ncu_traffic.log 2>&1
python scripts/ncu_summary.py $O/ncu_c4.ncu-rep > $O/${R}_prof_c4_trace.md 2>/dev/null
python scripts/ncu_summary.py $O/ncu_c2.ncu-rep > $O/${R}_prof_c2_trace.md 2>/dev/null
rm -f $O/ncu_c2.ncu-rep   # the C4 report is kept (source page of the kernel the round worked on)
head -c 1500 $O/ncu_traffic.json
