"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total,
average and share (cold-cache, serialised timings: compare SHARES, not absolutes)."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10 and r[0].isdigit()]
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[4])[:62]
    tot[name] += float(r[-1]) * 1e-6
    cnt[name] += 1
total = sum(tot.values())
print(f"# total {total:.3f} ms over {len(rows)} launches")
for name in sorted(tot, key=tot.get, reverse=True):
    print(f"{name:62s} n={cnt[name]:4d} total={tot[name]:9.3f} ms avg={tot[name] / cnt[name] * 1e3:9.1f} us "
          f"share={100 * tot[name] / total:5.1f}%")
