"""Kernel-time shares from an ncu launch list (--metrics gpu__time_duration.sum --csv): usage launch_shares.py <csv> [title]"""
import collections
import csv
import sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
    tot[r[ki]] += v; cnt[r[ki]] += 1
allt = sum(tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "(%d launches)" % sum(cnt.values()))
for k, v in tot.most_common():
    print("%-62s n=%5d  sum=%11.1f us  mean=%8.1f us  share=%.3f" % (k[:60], cnt[k], v, v / cnt[k], v / allt))
lib = [k for k in tot if not (k.startswith(("hive::", "hive_", "mcts_")) or "chw_to_nhwc64" in k or "hash_eval" in k)]
print("kernels that are not this repo's:", lib if lib else "none")
