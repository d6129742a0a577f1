"""BASELINE configs[3] and configs[4] (run under torchrun, one rank per GPU):
  [3] self-play, 250 sims/move, 8,192 games sharded over the ranks, NCCL weight broadcast + sample all-gather
  [4] evaluator match, 500 sims/move, new vs best net, 1,024 games per GPU
Prints one JSON line per config on rank 0 (moves/s whole job, max-over-ranks time)."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
import hive_b200
from importlib import import_module
par = import_module("hive-alphazero_b200.parallel")

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
def allmax(x):
    if world == 1: return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
def allsum(x):
    if world == 1: return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())
def barrier():
    if world > 1: dist.barrier()

moves3 = int(sys.argv[1]) if len(sys.argv) > 1 else 2
stream = torch.cuda.Stream()
torch.manual_seed(rank)                                       # differing weights until the broadcast
net = hive_b200.HiveNet().eval().cuda()
bytes_b = par.broadcast_weights(net, src=0)
start, cnt = par.shard_games(8192, world, rank)
folded = hive_b200.FoldedNet(net, device="cuda").attach_trunk(stream_ptr=stream.cuda_stream, max_boards=max(cnt, 1024))
with torch.cuda.stream(stream):
    sp = hive_b200.SelfPlayBatch(cnt, 250, hive_b200.LeafEvaluator(folded), device=lr, stream=stream.cuda_stream, seed=100 + rank,
                                 wave_graph=hive_b200.WaveGraph(stream))
    for _ in range(7): sp.env.step_random(5 + rank, 55, False)
    sp.play_moves(1)                                          # warm-up move: autotune, graph capture, allocations
    barrier(); torch.cuda.synchronize()
    r = sp.play_moves(moves3)
    torch.cuda.synchronize()
rows = torch.from_numpy(np.packbits((sp.env.planes_bf16() != 0).reshape(cnt, -1), axis=1)[:, :991].copy())
gathered = int(par.allgather_samples(rows, device="cuda").shape[0])
secs, moves = allmax(r["seconds"]), allsum(float(r["moves"]))
if rank == 0:
    print(json.dumps(dict(config="configs[3]: self-play 250 sims/move, 8192 games sharded over %d B200" % world, games_per_gpu=cnt,
                          moves=int(moves), seconds=secs, moves_per_s=moves / secs, sims_per_s=moves * 250 / secs,
                          tensor_util=moves * 250 / secs * 6.56e9 / (1411e12 * world), weights_broadcast_bytes=int(bytes_b),
                          samples_allgathered=gathered)), flush=True)
del sp
torch.manual_seed(1)
net_b = hive_b200.HiveNet().eval().cuda()
par.broadcast_weights(net_b, src=0)
folded_b = hive_b200.FoldedNet(net_b, device="cuda").attach_trunk(stream_ptr=stream.cuda_stream, max_boards=1024)
with torch.cuda.stream(stream):
    ev = hive_b200.EvaluatorMatch(1024, 500, hive_b200.LeafEvaluator(folded), hive_b200.LeafEvaluator(folded_b), device=lr,
                                  stream=stream.cuda_stream, seed=7 + rank, torch_stream=stream)
    ev.play(max_plies=5)                                       # 4 random plies + one searched ply (warm-up)
    barrier(); torch.cuda.synchronize()
    r0 = ev.waves
    r = ev.play(max_plies=2)                                   # two searched plies (one by each colour), timed
    r["waves"] = ev.waves - r0
    torch.cuda.synchronize()
searched = 2 * 1024.0
secs, moves = allmax(r["seconds"]), allsum(searched)
if rank == 0:
    print(json.dumps(dict(config="configs[4]: evaluator match 500 sims/move, 1024 games per GPU on %d B200" % world,
                          searched_moves=int(moves), seconds=secs, moves_per_s=moves / secs, sims_per_s=moves * 500 / secs,
                          tensor_util=moves * 500 / secs * 6.56e9 / (1411e12 * world), waves=r["waves"])), flush=True)
if world > 1:
    dist.destroy_process_group()
