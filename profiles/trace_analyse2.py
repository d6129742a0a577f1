"""Per-slice timeline of the two-kernel step from a per-CTA trace (profiles/trace_probe.py, -DHIVE_TRACE build):
launch durations (first CTA start -> last CTA end) of hive_step_kernel (id 0) and hive_planes_kernel (id 4), the gap
between consecutive step kernels of a slice and what the step kernel waited for.  usage: trace_analyse2.py trace.npz steps"""
import sys
import numpy as np
t = np.load(sys.argv[1])["trace"]
steps = int(sys.argv[2])
t0 = t["t0"].min()
s = (t["t0"] - t0) / 1e3
e = (t["t1"] - t0) / 1e3
print("records %d  span %.1f us  per step %.1f us" % (len(t), e.max(), e.max() / steps))
offs = np.unique(t["g_offset"])


def launches(k, o, n_expected):
    m = (t["kernel"] == k) & (t["g_offset"] == o)
    ss, ee = s[m], e[m]
    order = np.argsort(ss)
    ss, ee = ss[order], ee[order]
    per = len(ss) // n_expected
    return [(ss[i * per:(i + 1) * per].min(), ee[i * per:(i + 1) * per].max(), np.mean(ee[i * per:(i + 1) * per] - ss[i * per:(i + 1) * per]),
             np.percentile(ss[i * per:(i + 1) * per], 90) - ss[i * per:(i + 1) * per].min()) for i in range(n_expected)]


for o in offs[:2]:
    st, pl = launches(0, o, steps), launches(4, o, steps)
    print("slice at game %d: step launch [start, end] dur | mean CTA time | time until 90%% of its CTAs started || planes [start, end] dur" % o)
    for i in range(min(steps, 8)):
        gap = st[i][0] - st[i - 1][1] if i else 0.0
        print("  step %2d: [%7.1f %7.1f] %5.1f us | CTA %5.1f | ramp %5.1f | gap to previous step kernel %5.1f || planes [%7.1f %7.1f] %5.1f us" % (
            i, st[i][0], st[i][1], st[i][1] - st[i][0], st[i][2], st[i][3], gap, pl[i][0], pl[i][1], pl[i][1] - pl[i][0]))
durs = []; gaps = []; pdur = []; ctat = []; p_after = []
for o in offs:
    st, pl = launches(0, o, steps), launches(4, o, steps)
    durs += [b - a for a, b, _, _ in st]
    ctat += [c for _, _, c, _ in st]
    gaps += [st[i][0] - st[i - 1][1] for i in range(1, steps)]
    pdur += [b - a for a, b, _, _ in pl]
    p_after += [pl[i][0] - st[i][1] for i in range(steps)]
print("step kernel launch: mean %.1f us (CTA mean %.1f); gap between a slice's consecutive step kernels: mean %.1f us; planes launch mean %.1f us, starts %.1f us after its step kernel ended"
      % (np.mean(durs), np.mean(ctat), np.mean(gaps), np.mean(pdur), np.mean(p_after)))
print("period per step of a slice = %.1f us" % (np.mean(durs) + np.mean(gaps)))
