"""Where does the host-driven (e2e) loop of bench.py spend its time?  Same loop (`parts` pipelined parts of a 16,384-game
batch, graph-replayed hive_step_host_async, hive_wait_results), with the three calls of an iteration timed separately.
usage: python profiles/e2e_probe.py [parts] [iterations]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import hive_b200
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 3
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
n = 16384
P = []
for i in range(parts):
    cnt = n // parts + (1 if i < n % parts else 0)
    st = torch.cuda.Stream()
    b = hive_b200.HiveBatch(cnt, stream=st.cuda_stream)
    mask_h = torch.empty((cnt, 25), dtype=torch.int64).pin_memory(); count_h = torch.empty(cnt, dtype=torch.int32).pin_memory()
    status_h = torch.empty(cnt, dtype=torch.int32).pin_memory(); actions_h = torch.empty(cnt, dtype=torch.int32).pin_memory()
    ep = np.zeros(cnt, dtype=np.uint32)
    b.legal_into(mask_h.data_ptr(), count_h.data_ptr()); b.status_packed_into(status_h.data_ptr())
    P.append(dict(b=b, st=st, keep=(mask_h, count_h, status_h, actions_h, ep), n=cnt,
                  p=(mask_h.data_ptr(), count_h.data_ptr(), status_h.data_ptr(), ep.ctypes.data, actions_h.data_ptr())))
T = dict(wait=0.0, pick=0.0, launch=0.0)
for it in range(K + 30):
    if it == 30:
        for k in T: T[k] = 0.0
        torch.cuda.synchronize(); t_start = time.perf_counter()
    for h in P:
        pm, pc, ps, pe, pa = h["p"]
        t0 = time.perf_counter(); h["b"].wait_results()
        t1 = time.perf_counter(); hive_b200.host_pick_actions_ptr(h["n"], pm, pc, ps, pe, 7, 55, pa)
        t2 = time.perf_counter(); h["b"].step_async_ptr(pa, pm, pc, ps)
        t3 = time.perf_counter()
        T["wait"] += t1 - t0; T["pick"] += t2 - t1; T["launch"] += t3 - t2
for h in P: h["b"].sync()
total = time.perf_counter() - t_start
print({k: round(v / (K * parts) * 1e6, 1) for k, v in T.items()}, "us per part-iteration;", parts, "parts; full step",
      round(total / K * 1e6, 1), "us ->", round(n / (total / K) / 1e6, 1), "M env-steps/s (upper bound: live games only)")
