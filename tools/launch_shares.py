#!/usr/bin/env python3
"""ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X cmd`) -> the same CSV with
a header of per-kernel shares of the device time. usage: launch_shares.py launches.csv out.csv "command line that was profiled" """
import csv, sys, collections
src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(src, errors="replace") if not l.startswith("==") and not l.startswith("#")]
rows = list(csv.reader(lines))
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
    name = r[kn].split("(")[0][:70]
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
with open(dst, "w") as f:
    f.write("# shares of the device time of `%s` under ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares, not absolutes)\n" % cmd)
    for name, v in tot.most_common(14):
        f.write("# %5.1f %%  %4d launches  %9.3f ms  %s\n" % (100 * v / total, cnt[name], v, name))
    f.writelines(lines)
print("wrote", dst, "%d launches, %.1f ms" % (sum(cnt.values()), total))
