"""profiles/r02_conv_traffic.json from an ncu launch list of ONE bench run restricted to the tensor-core convolution
kernels (what `bench.py` reads for `roofline.traffic`):

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"igemm|wgrad" \
        --clock-control none --csv --log-file gpurun_out/conv_traffic.csv \
        python bench.py --steps 1 --warmup 1 --graph 0 --extras 0 --no-cpu-baseline
    python scripts/make_conv_traffic.py gpurun_out/conv_traffic.csv <steps in the run> <C-ABI conv calls per step> > profiles/r02_conv_traffic.json

The run executes the same step several times (warm-up, timed, end-to-end, eager roofline pass); the LAST step's
launches are used.  `dram_bytes_per_launch_avg` divides by the number of C-ABI convolution calls per step -- the unit
`bench.py` times (a stride-2 data gradient is one call and up to four kernel launches)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

path, n_steps, calls_per_step = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
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
assert len(items) % n_steps == 0, (len(items), n_steps)
per = len(items) // n_steps
items = items[-per:]
by = collections.OrderedDict()
for d in items:
    k = re.sub(r"^void ", "", re.sub(r"\(.*", "", d["name"])).replace("aadg::", "")
    a = by.setdefault(k, {"launches": 0, "dram_gb": 0.0, "ms": 0.0})
    a["launches"] += 1
    a["dram_gb"] += (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / 1e9
    a["ms"] += d.get("gpu__time_duration.sum", 0.0) * 1e3
total = sum(v["dram_gb"] for v in by.values()) * 1e9
import bench  # noqa: E402  (the workload string bench.py compares against)
sys.argv = [sys.argv[0]]
a = bench.parse_args()
print(json.dumps({
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:igemm|wgrad "
              "--clock-control none python bench.py --steps 1 --warmup 1 --graph 0 --extras 0 --no-cpu-baseline "
              "(last step of the run; cold-cache, serialised launches)",
    "workload": bench.workload_config(a, 1)["workload"],
    "conv_launches_per_step": calls_per_step,
    "kernel_launches_per_step": per,
    "dram_bytes_per_step": total,
    "dram_bytes_per_launch_avg": total / calls_per_step,
    "ncu_time_ms_per_step": sum(v["ms"] for v in by.values()),
    "by_kernel": by,
}, indent=1))
