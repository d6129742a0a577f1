"""N>1 host logic on CPU: world_size-2 gloo processes exercise the weight broadcast, the sample
all-gather and the game sharding used by the multi-GPU self-play path."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import importlib
    par = importlib.import_module("hive-alphazero_b200.parallel")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # different weights per rank before the broadcast
    net = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Linear(5, 3))
    nbytes = par.broadcast_weights(net, src=0)
    flat = torch.cat([p.data.reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    local = torch.full((rank + 2, 7), rank + 1, dtype=torch.uint8)      # ragged: 2 rows on rank 0, 3 on rank 1
    allrows = par.allgather_samples(local)
    start, cnt = par.shard_games(8193, world, rank)
    out.put((rank, same, nbytes, allrows.shape[0], allrows[:, 0].tolist(), start, cnt))
    dist.destroy_process_group()


def test_broadcast_allgather_shard_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, nbytes, rows, col, start, cnt in res:
        assert same and nbytes > 0
        assert rows == 5 and col == [1, 1, 2, 2, 2]
    assert (res[0][5], res[0][6]) == (0, 4097) and (res[1][5], res[1][6]) == (4097, 4096)


def test_pack_samples_roundtrip_shapes():
    import importlib
    par = importlib.import_module("hive-alphazero_b200.parallel")
    n = 3
    bits = np.random.RandomState(0).randint(0, 256, size=(n, 991)).astype(np.uint8)
    idx = np.full((n, 160), -1, dtype=np.int64); idx[:, :4] = [[1, 5, 9, 1583]] * n
    val = np.zeros((n, 160), dtype=np.float32); val[:, :4] = 0.25
    rec = par.pack_samples(bits, np.where(idx < 0, 0, idx), val, np.array([1, -1, 0]))
    assert rec.shape == (n, 1960) and (rec[:, :991] == bits).all()
    assert rec[:, 1952].tolist() == [2, 0, 1]
