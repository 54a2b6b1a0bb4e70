"""Per-kernel SASS evidence of libreach_sm100a.so: counts of the instructions that show the Blackwell-native path --
DMMA.8x8x4 (FP64 tensor pipe; tcgen05 has no f64 kind), UBLKCP (cp.async.bulk, TMA engine, both directions), UTMALDG
(cp.async.bulk.tensor, tensor-map TMA), SYNCS (mbarrier), LDGSTS (cp.async), plus registers per thread.
  python tools/sass_evidence.py > profiles/r02/sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "reachability.jl_b200" / "libreach_sm100a.so"
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+)", line)
    if m and cur:
        regs[cur] = int(m.group(1))
kernels = OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        kernels[cur]["total"] += 1
        for key in ("DMMA", "UBLKCP.S.G", "UBLKCP.G.S", "UTMALDG", "SYNCS", "LDGSTS", "DFMA", "SHFL", "BAR"):
            if op.startswith(key):
                kernels[cur][key] += 1
demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS evidence of {LIB.name} (cuobjdump -sass, sm_100a); one line per kernel")
print("# kernel | regs | instr | DMMA.8x8x4 | UBLKCP g->s | UBLKCP s->g | UTMALDG | SYNCS (mbarrier) | LDGSTS | DFMA | SHFL | BAR")
tot = Counter()
for (name, c), dn in zip(kernels.items(), demangle):
    short = re.sub(r"\(anonymous namespace\)::|<unnamed>::|reach::", "", dn).replace("void ", "")
    short = re.sub(r"\((?:bool|int|long|unsigned int)\)", "", short)     # template arguments print as (int)4, (bool)1
    short = re.sub(r"\(.*$", "", short)
    print(f"{short:60s} | {regs.get(name, '-')} | {c['total']} | {c['DMMA']} | {c['UBLKCP.S.G']} | {c['UBLKCP.G.S']} | {c['UTMALDG']} | "
          f"{c['SYNCS']} | {c['LDGSTS']} | {c['DFMA']} | {c['SHFL']} | {c['BAR']}")
    tot.update(c)
print(f"# total: {len(kernels)} kernels, {tot['DMMA']} DMMA, {tot['UBLKCP.S.G']} UBLKCP.S.G, {tot['UBLKCP.G.S']} UBLKCP.G.S, "
      f"{tot['UTMALDG']} UTMALDG, {tot['SYNCS']} SYNCS, {tot['LDGSTS']} LDGSTS")
