"""Where a self-play move's time goes beyond the network: times begin / 50 graph-replayed waves / tail check / read-back /
step with a sync after each, and the bare network forward on the same stream.  usage: python profiles/wave_gap_probe.py [n] [sims]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import hive_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 50
torch.manual_seed(0)
folded = hive_b200.FoldedNet(hive_b200.HiveNet().eval(), device="cuda")
stream = torch.cuda.Stream()
folded.attach_trunk(stream_ptr=stream.cuda_stream, max_boards=n)
x = torch.zeros(n, 56, 12, 12, device="cuda", dtype=torch.bfloat16)
def sync(): torch.cuda.synchronize()
with torch.cuda.stream(stream):
    for _ in range(3): folded(x)
    sync(); t0 = time.perf_counter()
    for _ in range(10): folded(x)
    sync(); net_ms = (time.perf_counter() - t0) / 10 * 1e3
    wg = hive_b200.WaveGraph(stream)
    sp = hive_b200.SelfPlayBatch(n, sims, hive_b200.LeafEvaluator(folded), stream=stream.cuda_stream, seed=1, wave_graph=wg)
    for _ in range(8):
        sp.env.step_random(3, 55, False)
    sp.play_moves(2)
    m = sp.mcts
    for rep in range(3):
        sync(); t = [time.perf_counter()]
        m.begin(None); sync(); t.append(time.perf_counter())
        wg.ensure(m, sp.evaluator, False)
        for _ in range(sims): wg.replay()
        sync(); t.append(time.perf_counter())
        left = m.descend(); sync(); t.append(time.perf_counter())
        a = m.actions(); sync(); t.append(time.perf_counter())
        sp.env.step(a); sp.env.sync(); t.append(time.perf_counter())
        d = [round((b - a_) * 1e3, 2) for a_, b in zip(t[:-1], t[1:])]
        print("net fwd", round(net_ms, 2), "ms | begin", d[0], "| %d waves" % sims, d[1], "= %.2f per wave" % (d[1] / sims), "| tail descend", d[2], "(pending %d)" % left,
              "| actions", d[3], "| step", d[4], flush=True)
    # one wave as separate pieces
    ev0, ev1, ev2, ev3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    m.begin(None); sync()
    from importlib import import_module
    L = import_module("hive-alphazero_b200._capi").lib()
    ev0.record(stream); L.mcts_descend(m._h, None); ev1.record(stream)
    sp.evaluator(m.dev_leaf_planes, m.dev_leaf_policy, m.dev_leaf_value, m.dev_pending_mask, m.n); ev2.record(stream)
    m.expand(); ev3.record(stream); sync()
    print("eager wave: descend+leaf eval", round(ev0.elapsed_time(ev1), 3), "ms | network", round(ev1.elapsed_time(ev2), 3), "| expand", round(ev2.elapsed_time(ev3), 3))
