#!/bin/bash
# dev: per-kernel time / DRAM traffic / instruction summary of one workload (ncu, serialised launches)
# usage: scripts/ncu_kernels.sh c2 16 [skip] [count]
W=${1:-c2}; SPP=${2:-16}; S=${3:-60}; C=${4:-160}
export PTB200_NO_FACADE=1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct,launch__grid_size --cache-control none --clock-control none -s $S -c $C --csv --log-file gpurun_out/nk_$W.csv python bench.py --workload $W --spp $SPP --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - gpurun_out/nk_$W.csv <<'PY'
import csv,sys,re,collections
rows=[l for l in open(sys.argv[1]) if l.startswith('"')]
agg=collections.defaultdict(lambda: collections.defaultdict(float)); cnt=collections.Counter()
for r in csv.DictReader(rows):
    name=re.sub(r"\(.*","",r["Kernel Name"]).replace("<unnamed>::","").replace("void ","")
    try: v=float(r["Metric Value"].replace(",",""))
    except ValueError: continue
    agg[name][r["Metric Name"]]+=v
    if r["Metric Name"]=="gpu__time_duration.sum": cnt[name]+=1
tot=sum(m["gpu__time_duration.sum"] for m in agg.values())
for k,m in sorted(agg.items(), key=lambda kv:-kv[1]["gpu__time_duration.sum"])[:12]:
    n=cnt[k]; t=m['gpu__time_duration.sum']
    bw=(m['dram__bytes_read.sum']+m['dram__bytes_write.sum'])/max(t,1)  # bytes/ns = GB/s
    print(f"{k[:26]:26s} n={n:3d} share {100*t/tot:5.1f}% avg {t/n/1e3:7.1f}us dram R {m['dram__bytes_read.sum']/n/1e6:7.1f} W {m['dram__bytes_write.sum']/n/1e6:7.1f} MB/launch = {bw:6.0f} GB/s L2hit {m['lts__t_sector_hit_rate.pct']/n:5.1f}% inst {m['smsp__inst_executed.sum']/n/1e6:6.2f}M issue {m['smsp__issue_active.avg.pct']/n:4.1f}% warps {m['sm__warps_active.avg.pct_of_peak_sustained_active']/n:4.1f}% regs {m['launch__registers_per_thread']/n:.0f} grid {m['launch__grid_size']/n:.0f}")
PY
