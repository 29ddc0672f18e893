#!/usr/bin/env python
"""Per-kernel SASS instruction counts of the built library (cuobjdump -sass; no GPU needed).

  python tools/sass_summary.py > profiles/sass_summary_rNN.txt

Columns: total instructions, then the mnemonics that prove what the kernels are made of —
UBLKCP (TMA bulk copy, cp.async.bulk), SYNCS (mbarrier), IMAD / IMAD.HI / IMAD.WIDE (integer multiply pipe),
DFMA (FP64 pipe), SHFL, LDS/STS (shared memory), LDG/STG (global), LDC/LDCU (constant bank), ACQBULK/… PDL
(griddepcontrol shows up as ACQBULK-free `NANOSLEEP`-less code: counted via the `PDL` column = GRIDDEPCTL... )."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ntt-gpu-qtesla_b200", "libqtesla_b200.so")
COLS = ["UBLKCP", "SYNCS", "IMAD", "IMAD.HI", "IMAD.WIDE", "IMAD.X", "IADD3", "LEA", "LOP3", "VIMNMX", "DFMA", "SHFL", "LDS", "STS",
        "LDG", "STG", "LDC", "LDCU", "BAR", "WARPSYNC", "PDL"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n

kernels = collections.OrderedDict()
cur = None
arch = set()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    cur["total"] += 1
    base = op.split(".")[0]
    if base == "IMAD":
        if op.startswith("IMAD.HI"):
            cur["IMAD.HI"] += 1
        elif op.startswith("IMAD.WIDE"):
            cur["IMAD.WIDE"] += 1
        elif op.startswith("IMAD.X"):
            cur["IMAD.X"] += 1
        elif op.startswith(("IMAD.MOV", "IMAD.IADD", "IMAD.SHL")):
            cur["IMAD(mov/iadd/shl)"] += 1
        else:
            cur["IMAD"] += 1
    elif base in ("ACQBULK", "GRIDDEPCTL") or op.startswith("GRIDDEP"):
        cur["PDL"] += 1
    elif base in COLS:
        cur[base] += 1

print(f"# {os.path.relpath(LIB, ROOT)}  arch = {', '.join(sorted(arch))}  ({len(kernels)} kernels)")
print("# counts are static SASS instructions per kernel (cuobjdump -sass), not executed instructions")
hdr = ["total"] + COLS + ["IMAD(mov/iadd/shl)"]
print("kernel | " + " | ".join(hdr))
tot = collections.Counter()
for name, c in kernels.items():
    tot.update(c)
    print(demangle(name)[:110] + " | " + " | ".join(str(c.get(h, 0)) for h in hdr))
print("ALL | " + " | ".join(str(tot.get(h, 0)) for h in hdr))
