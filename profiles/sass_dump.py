#!/usr/bin/env python
"""cuobjdump -sass of ONE kernel of an object file, encodings stripped, plus an opcode census
(what the listing is committed for: LDGSTS = cp.async prefetch, the DFMA / DMUL / DADD density,
MUFU seeds, and STL / LDL = local-memory spills).

    python profiles/sass_dump.py roskfpos_b200/csrc/kfpos_t6.o t6_replay_kernelILb0ELb0ELi8ELb0E profiles/r02_sass_t6_replay_m8.txt
"""
import collections
import re
import subprocess
import sys

obj, key, out = sys.argv[1:4]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
lines, on = [], False
for ln in txt.splitlines():
    if "Function :" in ln:
        on = key in ln
    if on:
        lines.append(ln)
ins = []
for ln in lines:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", ln)
    if m:
        ins.append((m.group(1), m.group(2).strip()))
ops = collections.Counter()
for _, t in ins:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    ops[t.split()[0].split(".")[0]] += 1
n = len(ins)
with open(out, "w") as f:
    f.write(f"# {obj}: {lines[0].strip()}\n# {n} SASS instructions (static); opcode census:\n")
    for op, c in ops.most_common():
        f.write(f"#   {op:10s} {c:6d}  {100.0 * c / n:5.1f} %\n")
    f.write(f"# LDGSTS (cp.async) {ops.get('LDGSTS', 0)}, STL {ops.get('STL', 0)}, LDL {ops.get('LDL', 0)}, "
            f"FP64 (DFMA+DMUL+DADD+DSETP) {sum(ops.get(k, 0) for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))}\n")
    for a, t in ins:
        f.write(f"{a}  {t}\n")
print(open(out).read().split("\n0000")[0])
