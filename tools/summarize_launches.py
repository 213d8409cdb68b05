"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel: count, ms, share."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
agg = collections.OrderedDict(); cnt = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[-44:]
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    agg[name] = agg.get(name, 0) + v; cnt[name] += 1
tot = sum(agg.values())
print("%-46s %5s %12s %7s" % ("kernel", "n", "ms", "share"))
for k, v in sorted(agg.items(), key=lambda t: -t[1]):
    print("%-46s %5d %12.3f %6.1f%%" % (k, cnt[k], v / 1e6, 100 * v / tot))
print("%-46s %5s %12.3f" % ("total", "", tot / 1e6))
