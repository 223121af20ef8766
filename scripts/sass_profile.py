"""Summarise `ncu --page source --csv` output: opcode mix, per-block lane utilisation, stall samples."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
isrc, isamp, iinst, ithr = [hdr.index(x) for x in ["Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"]]
tot_inst = sum(int(r[iinst]) for r in data); tot_samp = sum(int(r[isamp]) for r in data); tot_thr = sum(int(r[ithr]) for r in data)
print("SASS instrs", len(data), "warp-inst executed", tot_inst, "avg threads/inst %.2f" % (tot_thr / tot_inst), "samples", tot_samp)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc]); o = m.group(2).split('.')[0] if m else '?'
    op[o] += int(r[iinst]); ops[o] += int(r[isamp])
print("opcode: share of executed warp-instructions | share of stall samples")
for o, c in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 26):
    print(f"  {o:10s} {100*c/tot_inst:5.1f}% | {100*ops[o]/tot_samp:5.1f}%")
B = 80
print(f"\nby {B}-instruction blocks: inst share, threads/inst, sample share, first instr")
for i in range(0, len(data), B):
    blk = data[i:i + B]; ins = sum(int(r[iinst]) for r in blk); th = sum(int(r[ithr]) for r in blk); sm = sum(int(r[isamp]) for r in blk)
    if ins: print(f"  [{i:4d}] inst {100*ins/tot_inst:5.1f}% thr/inst {th/ins:5.1f} samples {100*sm/tot_samp:5.1f}%  {blk[0][isrc].strip()[:60]}")
