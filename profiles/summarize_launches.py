"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) per kernel.
usage: python profiles/summarize_launches.py gpurun_out/r02_launches.csv > profiles/r02_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
per = defaultdict(lambda: defaultdict(dict))
for r in rows:
    per[r[0]][r[12]] = (r[4], float(r[14].replace(",", "")), r[13])
agg = defaultdict(lambda: [0, 0.0, 0.0])
scale = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}
bscale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for lid, m in per.items():
    name = next(iter(m.values()))[0]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name).replace("lisb::", "")
    t = m.get("gpu__time_duration.sum"); rd = m.get("dram__bytes_read.sum"); wr = m.get("dram__bytes_write.sum")
    if not t:
        continue
    a = agg[name]
    a[0] += 1
    a[1] += t[1] * scale.get(t[2], 1e-6)
    a[2] += (rd[1] * bscale.get(rd[2], 1.0) if rd else 0.0) + (wr[1] * bscale.get(wr[2], 1.0) if wr else 0.0)
print(f"# ncu launch list {sys.argv[1]}, aggregated per kernel")
print("# cold-cache, serialised launches: compare shares, not absolutes. time in ms, traffic = dram read+write")
print(f"{'kernel':70s} {'launches':>8s} {'avg ms':>9s} {'avg GB':>8s} {'GB/s':>8s} {'total ms':>9s}")
for name, (c, ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if c == 0 or ms == 0:
        continue
    print(f"{name[:70]:70s} {c:8d} {ms / c:9.3f} {by / c / 1e9:8.3f} {by / ms / 1e6:8.0f} {ms:9.2f}")
