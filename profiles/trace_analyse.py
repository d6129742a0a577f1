"""Reads a per-CTA timeline written by profiles/trace_probe.py and prints, per kernel, how long a launch of one
slice takes under contention (first CTA start -> last CTA end), the dependency gaps inside a slice's chain and
how busy the SMs are.  usage: trace_analyse.py trace.npz [steps]"""
import sys
import numpy as np
t = np.load(sys.argv[1])["trace"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
names = ["analyse", "flood", "moves", "encode", "planes"]
t0 = t["t0"].min()
s = (t["t0"] - t0) / 1e3
e = (t["t1"] - t0) / 1e3
span = e.max()
print("records %d  span %.1f us  per step %.1f us" % (len(t), span, span / steps))
for k, nm in enumerate(names):
    m = t["kernel"] == k
    d = (e - s)[m]
    if not m.any():
        continue
    print("%-8s CTAs/step %7.0f  CTA time mean %6.2f p50 %6.2f p95 %6.2f us  CTA-us/step %9.0f" % (
        nm, m.sum() / steps, d.mean(), np.median(d), np.percentile(d, 95), d.sum() / steps))
# launches: group by (kernel, g_offset), split in time by gaps (a launch's CTAs start close together)
print("\nper-launch (one slice) duration under contention, and the gap to the next kernel of the chain:")
offs = np.unique(t["g_offset"])
launch = {}
for k in range(5):
    for o in offs:
        m = (t["kernel"] == k) & (t["g_offset"] == o)
        if not m.any():
            continue
        ss, ee = s[m], e[m]
        order = np.argsort(ss)
        ss, ee = ss[order], ee[order]
        # split into launches: `steps` launches per (kernel, slice); cluster by start time with k-means-free rule:
        cuts = np.nonzero(np.diff(ss) > 6.0)[0] + 1
        groups = np.split(np.arange(len(ss)), cuts)
        launch[(k, int(o))] = [(ss[g].min(), ee[g].max(), len(g)) for g in groups]
for k, nm in enumerate(names):
    durs = [b - a for (kk, o), L in launch.items() if kk == k for (a, b, n) in L]
    if not durs:
        continue
    cnt = [len(L) for (kk, o), L in launch.items() if kk == k]
    print("%-8s launches found per slice %s  duration mean %6.2f p50 %6.2f max %6.2f us" % (nm, sorted(set(cnt)), np.mean(durs), np.median(durs), np.max(durs)))
# chain gaps for slice 0..: end of kernel k launch j -> start of kernel k+1 launch j (only if counts match)
for k in range(3):
    gaps = []
    for o in offs:
        A, B = launch.get((k, int(o))), launch.get((k + 1, int(o)))
        if A and B and len(A) == len(B):
            gaps += [b[0] - a[1] for a, b in zip(A, B)]
    if gaps:
        print("gap %s -> %s: mean %5.2f p50 %5.2f us (n=%d)" % (names[k], names[k + 1], np.mean(gaps), np.median(gaps), len(gaps)))
# SM busy: fraction of the span in which an SM hosts at least one CTA, and mean resident CTAs
sms = np.unique(t["sm"])
grid = np.linspace(0, span, 4000)
busy = np.zeros(len(grid))
for k in range(5):
    m = t["kernel"] == k
    a = np.searchsorted(grid, s[m]); b = np.searchsorted(grid, e[m])
    occ = np.zeros(len(grid) + 1)
    np.add.at(occ, a, 1); np.add.at(occ, b, -1)
    occ = np.cumsum(occ)[:-1]
    print("%-8s mean resident CTAs on the GPU %7.1f" % (names[k], occ.mean()))
