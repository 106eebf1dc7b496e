#!/usr/bin/env python
"""Correlate ncu per-instruction samples (--page source --csv, SASS view) with CUDA source lines
through nvdisasm's line info.   usage: ncu_hot_lines.py <report.ncu-rep> <kernel substring> [top]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Address" in r and "# Samples" in r)
ai, ni, ii, si = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
inst = [(int(r[ai], 16), int(r[ni] or 0), int(r[ii] or 0), r[si].strip()) for r in rows
        if len(r) == len(hdr) and r[ai].startswith("0x")]
base = inst[0][0]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "lidar-processing_b200" / "liblidar_b200.so")], cwd=td,
                   capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
line_of = {}
cur, infn = None, False
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0])
for a, n, i, s in inst:
    k = line_of.get(a - base, ("?", 0))
    agg[k][0] += n
    agg[k][1] += i
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
src_cache = {}
print(f"samples {tot}  warp-instructions {toti}")
for (f, l), (n, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src_cache:
        p = next(iter(ROOT.rglob(f)), None) if f != "?" else None
        src_cache[f] = p.read_text().splitlines() if p else []
    text = src_cache[f][l - 1].strip()[:90] if 0 < l <= len(src_cache[f]) else ""
    print(f"{100*n/tot:5.1f}% smp {100*i/toti:5.1f}% inst  {f}:{l:<4d} {text}")
