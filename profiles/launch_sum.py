"""Sums an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name: python profiles/launch_sum.py file.csv [skip_launches]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
col = {n: i for i, n in enumerate(rows[hdr])}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 2:]:
    if len(r) <= col["Metric Value"]:
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    try:
        v = float(r[col["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[col["Metric Unit"]]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.1f} us {100 * v / total:5.1f} % {cnt[k]:5d} x  {k[:100]}")
print(f"{total:10.1f} us total")
