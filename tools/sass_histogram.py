"""SASS opcode histogram of the shipped library (cuobjdump -sass): which memory / synchronisation instructions the
kernels are made of, and which kernels hold the TMA (UTMALDG) and mbarrier (SYNCS) instructions.
usage: python tools/sass_histogram.py [path to .so] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "malstroem_b200", "libmalstroem_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
hist = collections.Counter()
per_kernel = collections.defaultdict(collections.Counter)
kernel = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kernel = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        hist[op] += 1
        per_kernel[kernel][op] += 1
sel = ["LDS", "LDG", "LDCU", "ATOMS", "STS", "LDC", "STG", "SHFL", "REDG", "BAR", "VOTE", "REDUX", "ATOMG", "VOTEU",
       "CREDUX", "MEMBAR", "NANOSLEEP", "MATCH", "SYNCS", "UTMALDG", "LDGSTS", "UBLKCP"]
print("SASS opcode histogram of %s (cuobjdump -sass, sm_100a), selected opcodes:" % os.path.relpath(lib))
for op in sorted(sel, key=lambda o: -hist[o]):
    if hist[op]:
        print("%7d %s" % (hist[op], op))
print("\nUTMALDG (cp.async.bulk.tensor.2d) + SYNCS (mbarrier) per kernel:")
for k, c in per_kernel.items():
    if c["UTMALDG"]:
        print("  %s: UTMALDG %d, SYNCS %d" % (k, c["UTMALDG"], c["SYNCS"]))
print("\nfull histogram (top 40):")
for op, n in hist.most_common(40):
    print("%7d %s" % (n, op))
print("\nkernels: %d" % len(per_kernel))
