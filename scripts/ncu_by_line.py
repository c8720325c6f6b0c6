#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction samples / executed instructions by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <kernel-regex> <launch-index> [top=40] [so=liblc3d.so]
Maps SASS offsets to file:line with `nvdisasm -g` on the cubin extracted from the built .so
(the library is compiled with -lineinfo).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, rx, idx = sys.argv[1], sys.argv[2], int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = os.path.abspath(sys.argv[5]) if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "lowcost3dreconstruction_b200", "csrc", "liblc3d.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:{idx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1]
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    ins.append((int(r[0], 16), r[col["Source"]].strip(), int(r[col["# Samples"]] or 0),
                int(r[col["Instructions Executed"]] or 0), int(r[col["Thread Instructions Executed"]] or 0)))
base = ins[0][0]
mangled = None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# pick the function whose demangled name matches the kernel name from the report
short = re.sub(r"\(.*", "", kname).split("::")[-1]
short = re.sub(r"<.*", "", short).replace("void ", "").strip()
targs = re.findall(r"<\(int\)(\d+)", kname)
bargs = re.findall(r"\(bool\)(\d)", kname)
cur_fn, line_of = None, {}
cur_line = ("?", 0)
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur_fn = m.group(1)
        continue
    if cur_fn is None or short not in cur_fn:
        continue
    if targs and f"ILi{targs[0]}E" not in cur_fn and "ILi" in cur_fn:
        continue
    if bargs and f"Lb{bargs[0]}E" not in cur_fn and "Lb" in cur_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur_line
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for addr, src, smp, inst, tinst in ins:
    key = line_of.get(addr - base, ("?", 0))
    a = agg[key]
    a[0] += smp
    a[1] += inst
    a[2] += tinst
    tot[0] += smp
    tot[1] += inst
    tot[2] += tinst
print(f"kernel: {kname[:100]}\n total samples {tot[0]}  warp-inst {tot[1]}  thread-inst {tot[2]}  avg active {tot[2]/max(tot[1],1):.1f}")
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        p = os.path.join(os.path.dirname(so), f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:80] if 0 < ln <= len(srcs[f]) else ""
    print(f"{a[0]/max(tot[0],1)*100:5.1f}% smp {a[1]/max(tot[1],1)*100:5.1f}% inst act {a[2]/max(a[1],1):4.1f}  {f}:{ln:<4d} {text}")
