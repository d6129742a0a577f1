"""Where does the host-driven (e2e) environment step spend its time?  Prints per-call means (us)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import hive_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
b = hive_b200.HiveBatch(n)
mask_h = torch.empty((n, 25), dtype=torch.int64).pin_memory(); count_h = torch.empty(n, dtype=torch.int32).pin_memory()
status_h = torch.empty(n, dtype=torch.int32).pin_memory(); actions_h = torch.empty(n, dtype=torch.int32).pin_memory()
mask_np, count_np = mask_h.numpy().view(np.uint64), count_h.numpy()
status_np, actions_np = status_h.numpy().view(np.uint32), actions_h.numpy()
episodes = np.zeros(n, dtype=np.uint32)
T = dict(legal=0.0, status=0.0, pick=0.0, step=0.0, sync=0.0)
K = 300
for it in range(K + 20):
    if it == 20:
        for k in T: T[k] = 0.0
    t0 = time.perf_counter(); b.legal_into(mask_h.data_ptr(), count_h.data_ptr())
    t1 = time.perf_counter(); b.status_packed_into(status_h.data_ptr())
    t2 = time.perf_counter(); hive_b200.host_pick_actions(mask_np, count_np, status_np, episodes, 7, 55, actions_np)
    t3 = time.perf_counter(); b.step_ptr(actions_h.data_ptr())
    t4 = time.perf_counter(); b.sync(); t5 = time.perf_counter()
    T['legal'] += t1 - t0; T['status'] += t2 - t1; T['pick'] += t3 - t2; T['step'] += t4 - t3; T['sync'] += t5 - t4
print({k: round(v / K * 1e6, 1) for k, v in T.items()}, 'us per step; n =', n)
