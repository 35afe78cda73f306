"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total us, share.
usage: python tools/launch_shares.py launches.csv [n_applies]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].replace("mrx::<unnamed>::", "").replace("unnamed>::", "")[:58]; v = float(r[vi].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("kernel,launches,total_us,share")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"\"{k}\",{a[0] / div:g},{a[1] / 1e3 / div:.1f},{a[1] / tot:.4f}")
