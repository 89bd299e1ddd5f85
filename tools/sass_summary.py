"""SASS / resource listing of the built libbpt.so (cuobjdump; no GPU needed): per kernel the registers, stack (= local memory for the
traversal stack and spills), shared memory, instruction count and the opcodes that matter for this path (local / global / shared
memory traffic, FP32 pipe, conversions on the XU pipe, votes). Usage: python tools/sass_summary.py > profiles/<name>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bisemutum-engine_b200", "csrc", "libbpt.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(d):
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    m = re.match(r"([\w:]+(?:<[^(]*>)?)\(", d)
    d = m.group(1) if m else d.split("(")[0]
    return re.sub(r"\(bool\)", "", d)


res = {}
txt = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt):
    res[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
ops = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); ops[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        ops[cur][m.group(1)] += 1
names = demangle(sorted(ops))
COLS = ["LDG", "STG", "LDL", "STL", "LDS", "STS", "ATOMG", "RED", "FFMA", "FMUL", "FADD", "FMNMX", "FMNMX3", "MUFU", "I2F", "I2FP", "F2I", "VOTE", "SHFL", "BAR", "CALL"]
print("# SASS / resource listing of libbpt.so (sm_100a; `cuobjdump --dump-resource-usage` + `cuobjdump -sass`, tools/sass_summary.py)\n")
print("No tcgen05 / TMA / MMA instruction is expected or present: nothing on this path is a contraction or a tile mover (DESIGN section 5).")
allops = collections.Counter()
for c in ops.values():
    allops.update(c)
tensor = {k: v for k, v in allops.items() if re.match(r"(UTC|TCGEN|HMMA|IMMA|QMMA|UTMA|UBLKCP|LDGSTS)", k)}
print("Tensor / bulk-copy opcodes in the whole library: %s\n" % (tensor or "none"))
print("| kernel | registers | stack B | shared B | instructions | " + " | ".join(COLS) + " |")
print("|---|---|---|---|---|" + "---|" * len(COLS))
for k in sorted(ops, key=lambda n: short(names[n])):
    r = res.get(k, (0, 0, 0, 0))
    c = ops[k]
    n = sum(v for o, v in c.items() if o != "NOP")
    print("| `%s` | %d | %d | %d | %d | " % (short(names[k]), r[0], r[1], r[2], n) + " | ".join(str(c.get(o, 0)) for o in COLS) + " |")
