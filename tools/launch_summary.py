"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(r[ui], v / 1e6)
    name = r[ki].split("(")[0].replace("void ", "").replace("plb::", "")
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"| `{k}` | {cnt[k]} | {v:.2f} | {100 * v / s:.1f} % |")
print(f"| all | {sum(cnt.values())} | {s:.2f} | |")
