"""Per-kernel SASS mnemonic counts of the shipped library (runs without a GPU):
    python profiles/tools/sass_evidence.py > profiles/r02_sass_evidence.txt
DMMA = FP64 tensor-core MMA, UTMALDG = TMA tensor load, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg (B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "gaussianprocesses.jl_b200", "lib", "libgpb200.so")
COLS = ["DMMA", "UTMALDG", "SYNCS", "USETMAXREG", "LDS", "STS", "LDL", "STL", "DFMA", "DMUL", "DADD", "MUFU", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", sass)
try:
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
except Exception:
    pass
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        for c in COLS:
            if op == c or op.startswith(c + "."):
                counts[cur][c] += 1
        counts[cur]["_all"] += 1


def short(n):
    d = demangle.get(n, n)
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    d = re.sub(r"\((?!int\))[^<>]*$", "", d)         # drop the argument list, keep template arguments
    return d[:58]


print("# cuobjdump -sass gaussianprocesses.jl_b200/lib/libgpb200.so  (nvcc -gencode arch=compute_100a,code=sm_100a), produced by")
print("# profiles/tools/sass_evidence.py; per kernel: instruction count and the mnemonics that show the data path")
print("%-58s %6s " % ("kernel", "instr") + " ".join("%7s" % c for c in COLS))
for n, c in counts.items():
    print("%-58s %6d " % (short(n), c["_all"]) + " ".join("%7d" % c[k] for k in COLS))
