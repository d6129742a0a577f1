"""Where the host time of a collecting self-play move goes (SelfPlayBatch.play_moves with collect=True): times the read-backs
and the sample assembly of one move of 2,048 games.  usage: python profiles/collect_probe.py [n_games]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import hive_b200
from importlib import import_module
C = import_module("hive-alphazero_b200.config")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.manual_seed(0)
folded = hive_b200.FoldedNet(hive_b200.HiveNet().eval(), device="cuda")
stream = torch.cuda.Stream()
folded.attach_trunk(stream_ptr=stream.cuda_stream, max_boards=n)
with torch.cuda.stream(stream):
    wg = hive_b200.WaveGraph(stream)
    sp = hive_b200.SelfPlayBatch(n, 50, hive_b200.LeafEvaluator(folded), stream=stream.cuda_stream, seed=1, wave_graph=wg, collect=True)
    sp.play_moves(2)
    for rep in range(3):
        t = [time.perf_counter()]
        turn, winner, done = sp.env.status(); t.append(time.perf_counter())
        live = np.ones(n, dtype=np.uint8)
        sp.mcts.search_device(sp.evaluator, tree_mask=live, graph=wg); t.append(time.perf_counter())
        pin_pi, pin_planes, pin_action, pin_sum = sp._pinned()
        pi, act, _ = sp.mcts.policy(out=(pin_pi, pin_action, pin_sum)); t.append(time.perf_counter())
        mask, count = sp.env.legal_mask(); t.append(time.perf_counter())
        bits = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little")[:, :C.ACTION_SPACE].astype(bool); t.append(time.perf_counter())
        planes = sp.env.planes_bf16(out=pin_planes); t.append(time.perf_counter())
        chosen = sp._choose_batch(pi, act, bits, count, turn); t.append(time.perf_counter())
        live_idx = np.arange(n)
        kept_planes, kept_pi, side = planes[live_idx], pi[live_idx].astype(np.float32), (turn[live_idx] % 2).tolist()
        for k, g in enumerate(live_idx.tolist()):
            sp.samples[g].append((kept_planes[k], kept_pi[k], side[k]))
        t.append(time.perf_counter())
        sp.env.step(chosen); sp.env.sync(); t.append(time.perf_counter())
        names = ["status", "search", "policy()", "legal_mask()", "unpackbits", "planes_bf16()", "choose", "append loop", "step"]
        print({k: round((b - a) * 1e3, 2) for k, a, b in zip(names, t[:-1], t[1:])}, flush=True)
    # the product loop itself
    for rep in range(3):
        t0 = time.perf_counter(); r = sp.play_moves(1); dt = time.perf_counter() - t0
        print("play_moves(1):", round(dt * 1e3, 1), "ms; waves", r["waves"], flush=True)
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable(); sp.play_moves(2); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
