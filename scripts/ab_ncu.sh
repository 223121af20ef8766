#!/bin/bash
export PTB200_NO_FACADE=1
for v in cur head; do
  if [ $v = cur ]; then unset PTB200_LIB; else export PTB200_LIB=$PWD/variants/libptb200_$v.so; fi
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,launch__registers_per_thread,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active --cache-control none --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/ab_$v.csv python bench.py --workload c1 --spp 16 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  echo "== $v"; python - gpurun_out/ab_$v.csv <<'PY'
import csv,sys,re,collections
rows=[l for l in open(sys.argv[1]) if l.startswith('"')]
agg=collections.defaultdict(lambda: collections.defaultdict(float)); cnt=collections.Counter()
for r in csv.DictReader(rows):
    name=re.sub(r"\(.*","",r["Kernel Name"]).replace("<unnamed>::","")
    v=float(r["Metric Value"].replace(",",""))
    agg[name][r["Metric Name"]]+=v
    if r["Metric Name"]=="gpu__time_duration.sum": cnt[name]+=1
for k,m in sorted(agg.items(), key=lambda kv:-kv[1]["gpu__time_duration.sum"])[:6]:
    n=cnt[k]; print(f"{k[:28]:28s} n={n:3d} time {m['gpu__time_duration.sum']/1e3:8.1f}us dramR {m['dram__bytes_read.sum']/1e6:8.1f}MB dramW {m['dram__bytes_write.sum']/1e6:8.1f}MB L2hit {m['lts__t_sector_hit_rate.pct']/n:5.1f}% inst {m['smsp__inst_executed.sum']/1e6:8.1f}M regs {m['launch__registers_per_thread']/n:.0f} occlim {m['launch__occupancy_limit_registers']/n:.0f} warps {m['sm__warps_active.avg.pct_of_peak_sustained_active']/n:.1f}%")
PY
done
