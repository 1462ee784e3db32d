"""Per-kernel totals of the LAST repetition in an ncu gpu__time_duration launch list (csv)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
data = rows[1:]
half = data[len(data) // 2:]
agg = collections.OrderedDict()
for r in half:
    v = float(r[iv].replace(",", ""))
    v = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
    a = agg.setdefault(r[ik].split("(")[0][:60], [0, 0.0, []])
    a[0] += 1; a[1] += v; a[2].append(round(v, 3))
tot = 0
for k, (n, t, l) in agg.items():
    tot += t
    print(f"{k:62s} {n:3d} {t:8.3f} ms  {l[:8]}")
print("sum", round(tot, 3))
