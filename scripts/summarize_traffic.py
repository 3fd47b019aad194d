"""Summarise an ncu launch list that carries gpu__time_duration.sum and dram__bytes_{read,write}.sum:
per kernel name the launch count, summed time, summed DRAM bytes and the resulting GB/s.

    python scripts/summarize_traffic.py gpurun_out/launches.csv [skip] [take]
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
idi, ki, mi, vi, ui = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
SCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
         "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}
launch = collections.OrderedDict()
for row in r:
    if len(row) <= vi:
        continue
    d = launch.setdefault(int(row[idi]), {"name": row[ki]})
    d[row[mi]] = float(row[vi].replace(",", "")) * SCALE.get(row[ui], 1)
items = list(launch.values())
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
take = int(sys.argv[3]) if len(sys.argv) > 3 else len(items)
items = items[skip:skip + take]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in items:
    k = re.sub(r"^void ", "", re.sub(r"\(.*", "", d["name"]))
    a = agg[k]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in agg.values())
tb = sum(v[2] + v[3] for v in agg.values())
print("launches %d  time %.3f ms  dram %.2f GB  (%.0f GB/s overall)" % (len(items), tot * 1e3, tb / 1e9, tb / tot / 1e9))
print("%7s %10s %6s %9s %9s %8s  %s" % ("share", "ms", "n", "rd GB", "wr GB", "GB/s", "kernel"))
for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%6.2f%% %10.3f %6d %9.3f %9.3f %8.0f  %s" % (100 * t / tot, t * 1e3, n, rd / 1e9, wr / 1e9,
                                                      (rd + wr) / t / 1e9 if t else 0, k[:100]))
