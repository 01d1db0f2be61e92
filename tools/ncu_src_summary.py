#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: dynamic opcode mix, pipe estimate, lane efficiency
per opcode and the hottest SASS ranges.  Usage: ncu -i X.ncu-rep --page source --csv --print-source sass | python tools/ncu_src_summary.py [kernel_index]"""
import csv, sys, collections, re
rows = list(csv.reader(sys.stdin))
# split per kernel
kernels = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
ki = int(sys.argv[1]) if len(sys.argv) > 1 else 0
k = kernels[ki]; h = k["hdr"]
iS, iI, iT = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
ALU = {"IADD3","LOP3","SHF","PRMT","FMNMX","FMNMX3","FSETP","ISETP","SEL","FSEL","MOV","LEA","PLOP3","I2FP","IABS","FCHK","VOTE","POPC","FLO","BREV","IMNMX","VIADD","VIMNMX","VIMNMX3","UIADD3","R2P","P2R","CS2R","LEA.HI","IADD"}
FMA = {"FFMA","FMUL","FADD","IMAD","HFMA2","HADD2","HMUL2"}
tot_i = tot_t = 0; by = collections.Counter(); byt = collections.Counter()
for r in k["rows"]:
    src = r[iS].strip(); i = int(r[iI] or 0); t = int(r[iT] or 0)
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    by[op] += i; byt[op] += t; tot_i += i; tot_t += t
print(k["name"]); print("warp-instr %.3e thread-instr %.3e avg lanes %.2f" % (tot_i, tot_t, tot_t / max(tot_i, 1)))
a = sum(v for o, v in by.items() if o in ALU); f = sum(v for o, v in by.items() if o in FMA)
print("ALU-pipe share %.1f%%  FMA-pipe share %.1f%%  other %.1f%%" % (100 * a / tot_i, 100 * f / tot_i, 100 * (tot_i - a - f) / tot_i))
for op, v in by.most_common(28):
    print("  %-8s %5.1f%%  lanes %.1f" % (op, 100 * v / tot_i, byt[op] / max(v, 1)))
# hottest contiguous ranges: bucket by 64 instructions
print("ranges (64-instr buckets): start_idx share lanes")
n = len(k["rows"])
for s in range(0, n, 64):
    seg = k["rows"][s:s + 64]
    i = sum(int(r[iI] or 0) for r in seg); t = sum(int(r[iT] or 0) for r in seg)
    if i > 0.01 * tot_i: print("  %5d  %5.1f%%  %.1f   %s" % (s, 100 * i / tot_i, t / max(i, 1), seg[0][iS].strip()[:50]))
