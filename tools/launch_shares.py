"""Kernel shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`).
usage: python tools/launch_shares.py <launches.csv> <iterations covered>"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
iters = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
head = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[head]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
us, cnt, order = collections.Counter(), collections.Counter(), []
for r in rows[head + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"^void\s+|^svcm::", "", r[ki].split("(")[0]).replace("svcm::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
    if name not in us:
        order.append(name)
    us[name] += v; cnt[name] += 1
total = sum(us.values())
for k in order:
    print("%-26s launches/iter %4.1f  us/iter %8.1f  share %5.1f%%" % (k, cnt[k] / iters, us[k] / iters, 100 * us[k] / total))
print("%-26s launches/iter %4.1f  us/iter %8.1f" % ("TOTAL", sum(cnt.values()) / iters, total / iters))
