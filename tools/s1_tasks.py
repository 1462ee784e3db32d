"""Summarise the per-task lines of an S1_TRACE=2 stage-1 probe log (tools/s1_probe.py)."""
import collections
import re
import sys

rows = []
for l in open(sys.argv[1]):
    m = re.search(r"task\s+(\d+) cls\s+(\d+) npts\s+(\d+) nev\s+(\d+) nvox\s+(\d+) labels\s+(\d+) active\s+(\d+) windows\s+(\d+) \| kcyc build (\d+) transl (\d+) replay (\d+) wb (\d+)", l)
    if m:
        rows.append(tuple(map(int, m.groups())))
tot = collections.Counter()
for r in rows:
    tot["build"] += r[8]; tot["setup"] += r[9]; tot["replay"] += r[10]; tot["wb"] += r[11]
print(dict(tot))
rows.sort(key=lambda r: -(r[8] + r[9] + r[10] + r[11]))
for r in rows[:10]:
    print(r, "busiest/all %.2f" % (r[7] / max(r[6], 1)), "kcyc total", r[8] + r[9] + r[10])
for lo, hi in ((0, 2048), (2048, 8192), (8192, 16384), (16384, 10**9)):
    sel = [r for r in rows if lo < r[3] <= hi]
    if sel:
        print(f"nev({lo},{hi}]: {len(sel)} tasks, mean kcyc setup {sum(r[9] for r in sel) / len(sel):.0f} replay+wb {sum(r[10] + r[11] for r in sel) / len(sel):.0f}, "
              f"max total {max(r[8] + r[9] + r[10] + r[11] for r in sel)}")
