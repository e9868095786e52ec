"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.  python tools/launch_summary.py file.csv [split-kernel-substring]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    seq.append((name, v))
marks = [i for i, (n, _) in enumerate(seq) if len(sys.argv) > 2 and sys.argv[2] in n] + [len(seq)]
parts = [seq[:marks[0]]] + [seq[marks[i]:marks[i + 1]] for i in range(len(marks) - 1)]
for pi, part in enumerate(parts):
    if not part:
        continue
    agg = collections.OrderedDict()
    for n, v in part:
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += v
    tot = sum(t for _, t in agg.values())
    print("segment %d: %d launches, %.1f us" % (pi, len(part), tot))
    for n, (c, t) in agg.items():
        print("  %-44s %5d  %11.1f us  avg %9.1f  %5.1f %%" % (n[-44:], c, t, t / c, 100 * t / tot))
