"""Opcode histogram per kernel of libsde_loss.so (cuobjdump -sass): instruction count, the opcodes that prove the
Blackwell paths (UTMALDG = TMA tensor loads, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed fp32, ACQBULK / PREEXIT =
programmatic dependent launch) and every atomic with its type (no ATOM*.F32 / RED*.F32 anywhere: the reductions are
fixed-order, the scatters integer).  usage: python tools/sass_histogram.py [lib.so] > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "simpledepthestimation_b200", "libsde_loss.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(sde::\w+(, sde::\w+)*\)$", "", cur)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
arch = re.findall(r"arch = (sm_\w+)", txt)
print(f"# {os.path.basename(lib)}: {len(kernels)} kernels, arch {sorted(set(arch))}")
KEY = ("UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "ACQBULK", "PREEXIT", "MEMBAR", "BAR")
float_atomics = []
for name, ops in kernels.items():
    total = sum(ops.values())
    base = collections.Counter()
    for op, n in ops.items():
        base[op.split(".")[0]] += n
    key = {k: sum(n for op, n in ops.items() if op.startswith(k)) for k in KEY}
    atom = {op: n for op, n in ops.items() if op.startswith(("ATOM", "RED"))}
    float_atomics += [(name, op) for op in atom if ".F32" in op or ".F64" in op or ".F16" in op]
    print(f"\n## {name}\ninstructions {total}; " + ", ".join(f"{k} {v}" for k, v in key.items() if v))
    print("top: " + ", ".join(f"{op} {n}" for op, n in base.most_common(14)))
    if atom:
        print("atomics: " + ", ".join(f"{op} {n}" for op, n in sorted(atom.items())))
print("\n# floating-point atomics: " + (", ".join(f"{k}:{o}" for k, o in float_atomics) if float_atomics else "none"))
