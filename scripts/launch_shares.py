"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) > mv:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        tot[r[kn][:60]] += v
        cnt[r[kn][:60]] += 1
s = sum(tot.values())
for k, v in tot.most_common(15):
    print(f"{k:62s} n={cnt[k]:4d} total={v / 1e3:9.1f}us share={100 * v / s:5.1f}% avg={v / cnt[k] / 1e3:8.2f}us")
