"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for row in r:
    if len(row) <= vi:
        continue
    name = row[ki].split("(")[0][:70]
    v = float(row[vi].replace(",", ""))
    v = v / 1e3 if row[ui] == "ns" else v * 1e3 if row[ui] == "ms" else v
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(f"total {T/1e3:.2f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v/1e3:9.2f} ms {100*v/T:5.1f}%  n={cnt[k]:5d}  avg {v/cnt[k]:9.1f} us  {k}")
