"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for row in r:
    if len(row) <= vi: continue
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1)
    rows.append((row[ki], ns))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
take = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[skip:skip + take]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, ns in rows:
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"^void ", "", k)
    agg[k][0] += 1; agg[k][1] += ns
tot = sum(v[1] for v in agg.values())
print("launches %d total %.3f ms" % (len(rows), tot / 1e6))
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%6.2f%% %9.3f ms %5d  %s" % (100 * ns / tot, ns / 1e6, n, k[:110]))
