#!/usr/bin/env python
"""Attribute the per-SASS-instruction counts of an ncu source page to CUDA source lines using nvdisasm -g line markers.
   python tools/ncu_by_line.py <sass_csv from ncu --page source --csv --print-source sass> <nvdisasm -g -c listing> <mangled kernel name> [top]
The build profiled and the build disassembled must be the same."""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows = list(csv.reader(open(src_csv)))
h = rows[1]; iS, iI, iT, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("L1 Wavefronts Shared")
body = [r for r in rows[2:] if len(r) == len(h)]
# walk the disassembly of the kernel: collect (line marker, instruction) in order
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur = ("?", 0); seq = []
for l in lines[start + 1:]:
    if l.startswith("//--------------------- ") or l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        inl = "inlined" in l
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((cur, m.group(2).strip()))
assert len(seq) == len(body), (len(seq), len(body))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (loc, ins), r in zip(seq, body):
    a = agg[loc]; a[0] += int(r[iI] or 0); a[1] += int(r[iT] or 0); a[2] += int(r[iW] or 0)
ti = sum(a[0] for a in agg.values()); tw = sum(a[2] for a in agg.values())
srcs = {}
def text(loc):
    f, n = loc
    if f not in srcs:
        try: srcs[f] = open("/root/repo/raytracing-course-2024_b200/csrc/" + f).read().split("\n")
        except Exception: srcs[f] = []
    return srcs[f][n - 1].strip()[:100] if 0 < n <= len(srcs[f]) else ""
print("warp-instr %.3e   shared wavefronts %.3e" % (ti, tw))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% instr  lanes %4.1f  %5.2f%% wf  %s:%d  %s" % (100 * a[0] / ti, a[1] / max(a[0], 1), 100 * a[2] / max(tw, 1), loc[0], loc[1], text(loc)))
