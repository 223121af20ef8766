"""Summarise a .ncu-rep (raw page) into markdown: one row per captured kernel launch."""
import csv, subprocess, sys, json
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; units = rows[1]
want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__inst_executed.sum", "warp insts")]
res = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].replace("<unnamed>::", "")[:60]}
    for k, label in want:
        if k in hdr:
            d[label] = r[hdr.index(k)] + " " + units[hdr.index(k)]
    res.append(d)
for d in res:
    print("### `%s`\n" % d.pop("kernel"))
    print("| metric | value |\n|---|---|")
    for k, v in d.items():
        print(f"| {k} | {v} |")
    print()
