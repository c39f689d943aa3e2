#!/usr/bin/env python
"""profiles/sass_summary.txt: per kernel of libb200spice.so (sm_100a SASS from cuobjdump):
instruction count, registers / shared memory (cuobjdump -res-usage), and counts of the
mnemonics that show how each kernel moves data and does its FP64 work:
  UBLKCP / SYNCS  1-D bulk copies (TMA engine) and their mbarrier waits
  LDGSTS          cp.async (global -> shared without registers)
  LDG / STG       plain global loads / stores
  DFMA DMUL DADD  FP64 pipe;  MUFU  special-function unit (rcp / ex2 / lg2 / rsq seeds)
  HMMA/UTC*       tensor-core ops (expected: none -- no dense contraction on this path)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "xictools_b200", "libb200spice.so")
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]

usage = {}
cur = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", ln)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
        cur = None

arch = set(re.findall(r"arch = (sm_\w+)", sass))
kern = collections.OrderedDict()
name = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and name:
        op = m.group(1)
        kern[name]["_n"] += 1
        kern[name][op] += 1

cols = ["UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "DFMA", "DMUL", "DADD", "MUFU", "ATOMG", "BAR"]
out = [f"SASS summary of xictools_b200/libb200spice.so -- architectures in the fatbin: {sorted(arch)}",
       "tensor-core mnemonics (HMMA / IMMA / DMMA / UTCMMA / UTCHMMA / QGMMA): " +
       str(sum(c[k] for c in kern.values() for k in c if re.match(r"(H|I|D|Q)MMA|UTC.*MMA", k))) +
       "   bulk-tensor (UTMALDG): " + str(sum(c["UTMALDG"] for c in kern.values())),
       f"{'kernel':38s} {'instr':>7s} {'regs':>5s} {'smem':>6s} " + " ".join(f"{c:>6s}" for c in cols)]
for k, c in sorted(kern.items(), key=lambda kv: -kv[1]["_n"]):
    r, sm = usage.get(k, (0, 0))
    out.append(f"{demangle(k)[:38]:38s} {c['_n']:7d} {r:5d} {sm:6d} " + " ".join(f"{c[x]:6d}" for x in cols))
open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
