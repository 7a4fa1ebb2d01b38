"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) by kernel: launches, total
and mean device time, share of the listed time.  Usage: python scripts/launch_summary.py profiles/r2_launches_bench.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    name = re.sub(r"<.*", "", name) if name.startswith("at::") else name
    t = float(r[-1])
    unit = r[-2]
    t_us = t / 1e3 if unit in ("ns", "nsecond") else t * (1e3 if unit in ("ms", "msecond") else 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t_us
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot:.1f} us of device time (cold caches, serialised: read the SHARES, not the absolutes)")
print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'mean us':>9s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {t:10.1f} {t / n:9.2f} {100 * t / tot:6.1f}%")
