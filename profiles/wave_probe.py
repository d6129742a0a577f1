"""Inside a real search: how long does the network take per wave vs the whole wave?"""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
import hive_b200
n, sims = 2048, 50
torch.manual_seed(0)
stream = torch.cuda.Stream()
folded = hive_b200.FoldedNet(hive_b200.HiveNet().eval(), device="cuda").attach_trunk(stream_ptr=stream.cuda_stream, max_boards=n)
inner = hive_b200.LeafEvaluator(folded)
evs = []
def timed(*a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); inner(*a); e1.record(stream); evs.append((e0, e1))
with torch.cuda.stream(stream):
    sp = hive_b200.SelfPlayBatch(n, sims, timed, stream=stream.cuda_stream, seed=1)
    for _ in range(7): sp.env.step_random(3, 55, False)
    sp.play_moves(1); evs.clear()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record(stream); t0 = time.perf_counter()
    r = sp.play_moves(2)
    w1.record(stream); torch.cuda.synchronize(); dt = time.perf_counter() - t0
net_ms = [a.elapsed_time(b) for a, b in evs]
print(json.dumps(dict(waves=r["waves"], wall_ms_per_wave=dt * 1e3 / r["waves"], gpu_ms_per_wave=w0.elapsed_time(w1) / r["waves"],
                      net_ms_mean=float(np.mean(net_ms)), net_ms_min=float(np.min(net_ms)), net_ms_max=float(np.max(net_ms)))))
