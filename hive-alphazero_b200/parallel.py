"""Multi-GPU plumbing of self-play: games shard independently, one process per GPU.
The only collectives are a broadcast of the network weights from rank 0 (model update) and an
all-gather of finished self-play samples; nothing sits on the per-simulation path.

Backends: "nccl" on the B200 box, "gloo" in the CPU tests (world_size 2).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_games(total_games, world_size, rank):
    """Contiguous block of games owned by `rank` (block sizes differ by at most one)."""
    base, extra = divmod(int(total_games), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def broadcast_weights(module, src=0, wire_dtype=None):
    """Every rank ends up with rank `src`'s parameters and buffers (one flat broadcast per tensor dtype group).
    wire_dtype=torch.bfloat16 sends the floating-point tensors as bf16 (51.8 M parameters -> 103.6 MB instead of 207 MB,
    SURVEY.md 8e); every rank, the source included, then holds the bf16-rounded values, so all replicas stay identical.
    Returns the number of bytes broadcast."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    nbytes = 0
    groups = {}
    for t in tensors:
        groups.setdefault(t.dtype, []).append(t)
    for dtype, ts in groups.items():
        flat = torch.cat([t.reshape(-1) for t in ts])
        if wire_dtype is not None and dtype.is_floating_point and dtype != wire_dtype:
            flat = flat.to(wire_dtype)
        dist.broadcast(flat, src=src)
        nbytes += flat.numel() * flat.element_size()
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
    return nbytes


def pack_samples(planes_bits, pi_sparse_idx, pi_sparse_val, value):
    """One flat uint8 record per sample: 55 binary planes x 18 B + turn byte (991 B), up to 160
    (u16 action, f32 prob) policy entries (count byte + 960 B), value byte -> fixed 1,960-byte rows."""
    n = len(value)
    rec = np.zeros((n, 1960), dtype=np.uint8)
    rec[:, :991] = planes_bits
    cnt = np.minimum((pi_sparse_idx >= 0).sum(axis=1), 160).astype(np.uint8)
    rec[:, 991] = cnt
    rec[:, 992:992 + 320] = pi_sparse_idx.astype(np.uint16).view(np.uint8).reshape(n, -1)[:, :320]
    rec[:, 1312:1312 + 640] = pi_sparse_val.astype(np.float32).view(np.uint8).reshape(n, -1)[:, :640]
    rec[:, 1952] = (np.asarray(value) + 1).astype(np.uint8)
    return rec


def allgather_samples(local, device=None):
    """All-gather variable-length [k_i, row] uint8 sample blocks; returns the concatenation in rank
    order on every rank (torch tensor on `device`)."""
    local = torch.as_tensor(local)
    if device is not None:
        local = local.to(device)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    count = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    kmax = int(max(int(c.item()) for c in counts))
    padded = torch.zeros((kmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.cat([o[:int(c.item())] for o, c in zip(out, counts)], dim=0)
