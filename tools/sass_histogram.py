"""SASS opcode histogram of the kernels in libgvb200.so (cuobjdump -sass): which Blackwell instructions the
hand-written kernels compile to.  Usage: python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "genevector_b200", "libgvb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0] if not m.group(1).startswith(("UTC", "UTMA", "LDTM", "STTM", "SYNCS")) else m.group(1)] += 1
KEY = ("UTCIMMA", "UTCOMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCATOMSWS", "SYNCS")
print(f"{so}: {len(hist)} kernels\n")
print("== Blackwell tensor / TMA / tensor-memory / mbarrier instructions per kernel")
for k, h in hist.items():
    keys = {op: n for op, n in h.items() if op.startswith(KEY)}
    if keys:
        print(f"{k[:150]}\n    total {sum(h.values())} SASS instructions; " + ", ".join(f"{op} x{n}" for op, n in sorted(keys.items())))
sel = [k for k in hist if "gram_tiles_kernel<2, gv::MiEpi<10, true>" in k or "gram_tiles_kernel<(int)2, gv::MiEpi<(int)10, (bool)1>" in k]
for k in sel[:1]:
    print(f"\n== full opcode histogram of {k[:120]}")
    for op, n in hist[k].most_common():
        print(f"    {op:34s} {n}")
