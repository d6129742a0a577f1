import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
dt = t(lambda: x.zero_()); print("write-only (zero_ 1 GiB): %.0f GB/s" % (2**30 / dt / 1e9))
dt = t(lambda: y.copy_(x)); print("copy (read+write 2 GiB): %.0f GB/s" % (2 * 2**30 / dt / 1e9))
dt = t(lambda: x.sum()); print("read-only (sum 1 GiB): %.0f GB/s" % (2**30 / dt / 1e9))
