"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X`): per-kernel launches, total time, share."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = OrderedDict()
for r in rows[1:]:
    if len(r) <= iv or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("ms::", "")
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + v)
total = sum(t for _, t in tot.values())
print("total %.3f ms over %d launches (%d steps captured: %.3f ms per step)" % (total, sum(n for n, _ in tot.values()), steps, total / steps))
for name, (n, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print("%-44s %4d launches %9.3f ms  %5.1f%%" % (name, n, t, 100 * t / total))
