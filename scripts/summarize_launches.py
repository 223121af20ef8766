"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, sys, re, collections
path = sys.argv[1]
rows = [l for l in open(path) if l.startswith('"')]
rd = csv.DictReader(rows)
agg = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "")
    t = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("us", "usecond"): t *= 1e3
    a = agg.setdefault(name, [0, 0.0, r["Block Size"], r["Grid Size"]])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total us | share | block | grid |\n|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]/1e3:.1f} | {100*a[1]/tot:.1f}% | {a[2]} | {a[3]} |")
print(f"\ntotal {tot/1e3:.1f} us over {sum(a[0] for a in agg.values())} launches")
