"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: launch_table.py file.csv..."""
import collections, csv, re, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r[4])
        agg[name][0] += 1
        agg[name][1] += float(r[-1].replace(",", ""))
    print(f, "total %.1f ms, %d launches" % (sum(v[1] for v in agg.values()) / 1e6, len(rows)))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
        print(f"  {v[1]/1e6:9.2f} ms {v[0]:5d} x {v[1]/v[0]/1e3:8.1f} us  {k[:100]}")
