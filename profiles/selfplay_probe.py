import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
import hive_b200
n = int(sys.argv[1]); sims = int(sys.argv[2]); moves = int(sys.argv[3])
torch.manual_seed(0)
folded = hive_b200.FoldedNet(hive_b200.HiveNet().eval(), device="cuda")
stream = torch.cuda.Stream()
if len(sys.argv) > 4 and sys.argv[4] == "tc":
    folded.attach_trunk(stream_ptr=stream.cuda_stream, max_boards=n)
# pure net throughput
x = torch.zeros(n, 56, 12, 12, device="cuda", dtype=torch.bfloat16)
with torch.cuda.stream(stream):
    for _ in range(3): folded(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): folded(x)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("net fwd batch", n, "trunk", "tcgen05" if folded.trunk else "cudnn", "ms", dt * 1e3, "TFLOP/s", n * 6.56e9 / dt / 1e12, flush=True)
with torch.cuda.stream(stream):
    wg = hive_b200.WaveGraph(stream) if (len(sys.argv) > 5 and sys.argv[5] == 'graph') else None
    sp = hive_b200.SelfPlayBatch(n, sims, hive_b200.LeafEvaluator(folded), stream=stream.cuda_stream, seed=1, wave_graph=wg)
    sp.play_moves(1)
    r = sp.play_moves(moves)
    print(json.dumps(dict(phase='opening', n=n, sims=sims, **r, moves_per_s=r['moves'] / r['seconds'])), flush=True)
    for _ in range(6):
        sp.env.step_random(3, 55, False)            # leave the opening schedule (turn > 6)
    sp.play_moves(1)
    r = sp.play_moves(moves)
print(json.dumps(dict(n=n, sims=sims, **r, moves_per_s=r["moves"] / r["seconds"], sims_per_s=r["moves"] * sims / r["seconds"])))
