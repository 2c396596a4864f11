#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep.

    python profiles/by_line.py <file.ncu-rep> <object.o> <kernel-substring> [top]

The report's SASS page (ncu --page source --csv) gives executed counts and stall samples per
instruction; nvdisasm -g on the cubin inside <object.o> gives each instruction's source line.
Both list the kernel's instructions in address order, so they are joined by position.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, kname = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout

# ---- nvdisasm: instruction -> (file, line, opcode) for the wanted function
lines_of = []
cur_fn, cur_loc = None, ("?", 0)
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur_fn:
        lines_of.append((cur_fn, cur_loc, m.group(2).strip()))

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the page holds one table per kernel launch: "Kernel Name",<name> then a header row
tables, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        tables.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
tab = [t for t in tables if kname in t["name"]][0]
hdr = tab["hdr"]
iex, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]

# pick the disassembled function with the same instruction count
mangled = defaultdict(list)
for fn, loc, op in lines_of:
    mangled[fn].append((loc, op))
cands = [fn for fn, v in mangled.items() if len(v) == len(tab["rows"])]
if not cands:
    sys.exit(f"no disassembled function with {len(tab['rows'])} instructions: "
             f"{ {k: len(v) for k, v in mangled.items()} }")
fn = cands[0]
agg = defaultdict(lambda: [0, 0, 0, 0, defaultdict(int)])  # inst, samples, fp64 inst, n sass
tot_i = tot_s = tot_f = 0
kind = defaultdict(int)
for (loc, op), row in zip(mangled[fn], tab["rows"]):
    n, s = int(row[iex]), int(row[isamp])
    opc = op.split()[0] if not op.startswith("@") else op.split()[1]
    a = agg[loc]
    a[0] += n; a[1] += s; a[3] += 1
    base = opc.split(".")[0]
    kind[base] += n
    if base in ("DFMA", "DADD", "DMUL", "DSETP", "MUFU"):
        a[2] += n; tot_f += n
    for c in stall_cols:
        v = int(row[c] or 0)
        if v:
            a[4][hdr[c]] += v
    tot_i += n; tot_s += s
print(f"kernel {tab['name']}\nfunction {fn}: {len(tab['rows'])} SASS instr, {tot_i} warp-instr executed, "
      f"{tot_f} of them FP64/MUFU, {tot_s} samples")
print("opcode mix:", ", ".join(f"{k} {v / tot_i:.1%}" for k, v in sorted(kind.items(), key=lambda kv: -kv[1])[:16]))
print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'fp64%':>6s} {'sass':>5s}  top stalls")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ", ".join(f"{k[6:]} {v}" for k, v in sorted(a[4].items(), key=lambda kv: -kv[1])[:3])
    print(f"{loc[0] + ':' + str(loc[1]):28s} {a[0] / tot_i:6.1%} {a[1] / max(tot_s, 1):6.1%} "
          f"{a[2] / max(a[0], 1):6.1%} {a[3]:5d}  {st}")
