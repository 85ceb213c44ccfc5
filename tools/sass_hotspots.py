"""Attribute the executed warp-instructions of one kernel (ncu `--page source --csv` SASS view of a `--set full
--import-source on` report) to the source lines of its OUTERMOST inlining frame, via `nvdisasm -gi` line info of the shipped
cubin.  Usage:
    cuobjdump -xelf all libquadsim.so && nvdisasm -gi quadsim.sm_100a.cubin > all.sass
    ncu -i rep.ncu-rep --page source --csv > rep_source.csv
    python tools/sass_hotspots.py all.sass rep_source.csv <mangled kernel name> [outer file suffix]
Prints warp-instructions per outer source line (share of the kernel), top 60."""
import csv
import re
import sys
from collections import defaultdict

sass, rep, kern = sys.argv[1:4]
outer_file = sys.argv[4] if len(sys.argv) > 4 else "qs_rollout_tc.cuh"
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
ins = []          # (offset, outer line, innermost file:line)
block = []
block_open = False
pat = re.compile(r'File "([^"]+)", line (\d+)')
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section") or l.startswith(".section"):
        break
    if "//## File" in l:
        if not block_open:
            block = []
        block_open = True
        block += pat.findall(l)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m:
        block_open = False
        outer = [int(n) for f, n in block if f.endswith(outer_file)]
        inner = block[0] if block else ("?", "0")
        ins.append((int(m.group(1), 16), outer[-1] if outer else -1, f"{inner[0].split('/')[-1]}:{inner[1]}", m.group(2).split()[0]))
rows = list(csv.reader(open(rep)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = hdr.index(__import__("os").environ.get("HOTSPOT_COL", "Instructions Executed"))
data = [r for r in rows[hdr_i + 1:] if len(r) > ci and r[0]]
assert len(data) == len(ins), (len(data), len(ins))
by_outer = defaultdict(float)
by_inner = defaultdict(float)
tot = 0.0
for (off, outer, inner, op), r in zip(ins, data):
    n = float(r[ci].replace(",", "") or 0)
    tot += n
    by_outer[outer] += n
    by_inner[(outer, inner.split(":")[0])] += n
print(f"total warp-instructions {tot:.4g}")
src = open(sys.argv[5]).read().split("\n") if len(sys.argv) > 5 else None
for outer, n in sorted(by_outer.items(), key=lambda kv: -kv[1])[:60]:
    text = src[outer - 1].strip()[:90] if src and outer > 0 else ""
    files = sorted(((v, k[1]) for k, v in by_inner.items() if k[0] == outer), reverse=True)[:3]
    print(f"{outer:5d} {100 * n / tot:6.2f} %  {text}   [{', '.join(f'{f} {100 * v / tot:.1f}' for v, f in files)}]")
if len(sys.argv) > 6:          # breakdown of one outer line by innermost file:line
    want = int(sys.argv[6])
    inner_tot = defaultdict(float)
    ops = defaultdict(lambda: defaultdict(float))
    for (off, outer, inner, op), r in zip(ins, data):
        if outer == want:
            n = float(r[ci].replace(",", "") or 0)
            inner_tot[inner] += n
            ops[inner][op] += n
    print(f"--- outer line {want} by innermost line")
    for inner, n in sorted(inner_tot.items(), key=lambda kv: -kv[1])[:40]:
        f, ln = inner.rsplit(":", 1)
        text = src[int(ln) - 1].strip()[:80] if src and f == outer_file else ""
        top = ", ".join(f"{o} {100 * v / tot:.2f}" for o, v in sorted(ops[inner].items(), key=lambda kv: -kv[1])[:4])
        print(f"{inner:28s} {100 * n / tot:6.2f} %  {text}  [{top}]")
