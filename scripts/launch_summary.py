"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py launches.csv"""
import csv, sys, collections
tot = collections.OrderedDict()
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 14 and r[12] == "gpu__time_duration.sum":
        k = r[4][:90]
        c = tot.setdefault(k, [0, 0.0])
        c[0] += 1; c[1] += float(r[14].replace(",", "")) * (1e-6 if r[13] == "ns" else 1e-3 if r[13] in ("us", "usecond") else 1.0)
allms = sum(v[1] for v in tot.values())
print("%-92s %6s %10s %6s" % ("kernel", "count", "total ms", "share"))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-92s %6d %10.3f %5.1f%%" % (k, v[0], v[1], 100 * v[1] / allms))
