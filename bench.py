#!/usr/bin/env python
"""bench.py -- BASELINE.json metric on its quoted configuration.

Workload (config.workload): BASELINE configs[1] -- batched legal-move generation + step +
56-plane encode for 16,384 concurrent random games per GPU (on-device counter-based policy,
auto-reset at game end / turn 55).  One "step" = one launch of the env kernel over the whole
batch = one GamePlay.move() per live game.

  value     env-steps/s, whole job, state resident in HBM (device-timed, CUDA events on the
            stream the kernel runs on, max over ranks)
  e2e       same metric through the reference-facing C-ABI path with HOST buffers: every step
            copies the actions H2D from pinned memory, runs the kernel, reads legal masks /
            counts / status back D2H and picks the next actions on the host
  roofline  HBM: 17,094 algorithmic bytes per env-step (SURVEY 8d / DESIGN.md) over the measured
            copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the oracle port (oracle/hive_oracle.c) on the host cores, bounded sample

`--impl reference` times the CPU arm alone (the reference is pure Python and cannot travel to the
GPU box; the C restatement pinned to it by tests/golden stands in, kind="port").
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_ENV_STEP = 16128 + 198 + 768          # bf16 planes + legal mask + state record r/w
METRIC = "env_steps_per_s"
UNIT = "env-steps/s"
WORKLOAD = ("configs[1]: batched legal-move gen + step + plane encode, 16,384 concurrent random games "
            "per B200 (bit-exact vs ref)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=16384, help="concurrent games per GPU")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--max-turn", type=int, default=55)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-window-ms", type=float, default=50.0, help="the K-step block is repeated until the timed region is this long")
    ap.add_argument("--chunk", type=int, default=50, help="rollout steps per CUDA-graph launch (1 = one launch per step)")
    ap.add_argument("--no-selfplay", action="store_true", help="skip the self-play (configs[2]) side measurement")
    ap.add_argument("--selfplay-games", type=int, default=2048)
    ap.add_argument("--selfplay-sims", type=int, default=50)
    ap.add_argument("--selfplay-moves", type=int, default=2)
    ap.add_argument("--no-selfplay-deep", action="store_true", help="skip configs[3] (250 sims, sharded) and configs[4] (evaluator match)")
    ap.add_argument("--sharded-games", type=int, default=8192, help="configs[3]: games sharded over the ranks")
    ap.add_argument("--sharded-sims", type=int, default=250)
    ap.add_argument("--evaluator-games", type=int, default=1024, help="configs[4]: games per GPU")
    ap.add_argument("--evaluator-sims", type=int, default=500)
    ap.add_argument("--e2e-drivers", type=int, default=0, help="host-driven path: native driver threads (0: one per part)")
    ap.add_argument("--e2e-parts", type=int, default=0, help="host-driven path: parts of the batch, each its own handle and streams (0: two per driver thread)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hive_env_kernel_dram_bytes_per_launch")
        except Exception:
            pass
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:      # noqa
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_port_rate(seed, budget_s, threads, max_turn):
    """Bounded CPU sample: full counter-seeded games on `threads` host threads for ~budget_s."""
    from oracle import hive_oracle as ho
    t0 = time.perf_counter()
    pilot_games = 2 * threads
    pilot_steps = ho.play_many(seed, 0, pilot_games, max_turn, threads)
    pilot_dt = max(time.perf_counter() - t0, 1e-6)
    games = max(pilot_games, int(pilot_games * budget_s / pilot_dt))
    t0 = time.perf_counter()
    steps = ho.play_many(seed, pilot_games, games, max_turn, threads)
    dt = time.perf_counter() - t0
    return steps / dt, games, steps, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import hive_oracle as ho
    threads = os.cpu_count() or 1
    # size one "step" (a bounded sample of full games) so the whole run ends within ~2 minutes
    t0 = time.perf_counter()
    pilot = ho.play_many(args.seed, 0, 2 * threads, args.max_turn, threads)
    rate = pilot / max(time.perf_counter() - t0, 1e-6)
    per_step_s = 100.0 / max(args.steps + args.warmup, 1)
    games_per_step = max(threads, int(rate * per_step_s / 54.0))
    first = 2 * threads
    for _ in range(args.warmup):
        ho.play_many(args.seed, first, games_per_step, args.max_turn, threads)
        first += games_per_step
    total = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        total += ho.play_many(args.seed, first, games_per_step, args.max_turn, threads)
        first += games_per_step
    dt = time.perf_counter() - t0
    value = total / dt
    sample = "%d steps x %d full random games (counter-seeded, to terminal/turn %d) = %d env steps in %.1f s" % (
        args.steps, games_per_step, args.max_turn, total, dt)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic (counter-seeded random-legal-move games from reset)",
        "config": {"workload": WORKLOAD, "reference_arm": "oracle/hive_oracle.c port of the Python reference "
                   "(the reference itself is pure Python at ~38 env-steps/s/core and is absent on the GPU box)",
                   "games_per_step": games_per_step, "max_turn": args.max_turn},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _cpu_mcts_worker(job):
    """One process of the CPU search baseline: sequential 50-simulation searches of the oracle's HivePlayer restatement
    (hash-net stand-in for the network) from random positions; returns (moves, seconds)."""
    seed, budget_s, sims = job
    import numpy as np
    from oracle.hive_oracle import OracleEnv
    from oracle.mcts_oracle import MctsOracle, hash_net
    rng = np.random.RandomState(seed)
    env = OracleEnv()
    for _ in range(6 + seed % 12):
        la = env.actions()
        env.move(int(la[rng.randint(len(la))]) if len(la) else -1)
    np.random.seed(seed)
    moves, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        if env.game_is_over() or env.turn >= 55:
            env.reset()
        a = MctsOracle(hash_net, sims).action(env)[0]
        env.move(int(a))
        moves += 1
    return moves, time.perf_counter() - t0


def cpu_mcts_rate(budget_s, procs, sims=50):
    """moves/s of the sequential CPU search (oracle/mcts_oracle.py) on `procs` host processes, ~budget_s of wall time."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    t0 = time.perf_counter()
    with ProcessPoolExecutor(procs, mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_cpu_mcts_worker, [(100 + i, budget_s, sims) for i in range(procs)]))
    moves = sum(r[0] for r in res)
    dt = max(r[1] for r in res)
    return moves / dt, moves, dt, time.perf_counter() - t0


def run_selfplay(args, hive_b200, torch, dist, rank, world, local_rank, allsum, allmax, barrier):
    """BASELINE configs[2], [3], [4] -- AlphaZero self-play (50 sims, 2,048 games per GPU), sharded self-play (250 sims,
    8,192 games over the ranks) and the evaluator match (500 sims, two nets, 1,024 games per GPU).  Whole-job moves/s,
    sims/s and the tensor-pipe utilisation they imply; the NCCL weight broadcast (bf16 on the wire, weights reloaded in
    place) and the all-gather of the packed samples of the timed moves sit INSIDE the timed regions."""
    import numpy as np
    from importlib import import_module
    par = import_module("hive-alphazero_b200.parallel")
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        tf_peak, tf_src = float(peaks["bf16_tflops_sustained"]), "measured sustained (MEASURED_PEAKS.json)"
    except Exception:
        tf_peak, tf_src = 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained)"
    hbm_peak, _ = measured_peak()
    stream = torch.cuda.Stream()
    n2 = args.selfplay_games
    _, n3 = par.shard_games(args.sharded_games, world, rank)
    n4 = args.evaluator_games
    torch.manual_seed(0)
    net = hive_b200.HiveNet().eval().cuda()
    torch.manual_seed(1)
    net_best = hive_b200.HiveNet().eval().cuda()
    for m in (net, net_best):
        par.broadcast_weights(m, src=0, wire_dtype=torch.bfloat16)
    folded = hive_b200.FoldedNet(net, device="cuda").attach_trunk(stream_ptr=stream.cuda_stream, max_boards=max(n2, n3, n4))
    side_stream = torch.cuda.Stream()                          # the second network of the evaluator match runs beside the first
    folded_best = hive_b200.FoldedNet(net_best, device="cuda").attach_trunk(stream_ptr=side_stream.cuda_stream, max_boards=n4)
    if dist is not None:                                       # warm-up of the in-place reload (first launch of the packing kernels), untimed
        folded.reload(net); folded_best.reload(net_best)
        torch.cuda.synchronize()

    def timed_collectives_before(model, fold):
        """weight broadcast from rank 0 (bf16 on the wire) + in-place reload of the folded network; seconds, bytes"""
        t0 = time.perf_counter()
        nbytes = par.broadcast_weights(model, src=0, wire_dtype=torch.bfloat16) if dist is not None else 0
        if dist is not None:
            fold.reload(model)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, nbytes

    def selfplay_config(tag, n, sims, skip_plies, open_moves, timed_moves):
        with torch.cuda.stream(stream):
            sp = hive_b200.SelfPlayBatch(n, sims, hive_b200.LeafEvaluator(folded), device=local_rank, stream=stream.cuda_stream,
                                         seed=args.seed + rank, collect=True, wave_graph=hive_b200.WaveGraph(stream))
            for _ in range(skip_plies):                         # random plies from reset (synthetic random-opening positions)
                sp.env.step_random(args.seed + 5 + rank, args.max_turn, False)
            sp.play_moves(1)                                    # warm-up move: graph capture, allocations
            r_open = None
            if open_moves:
                barrier(); torch.cuda.synchronize()
                r_open = sp.play_moves(open_moves)              # the reference's opening schedule (noise mix, policy read-back)
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            t_bcast, nbytes = timed_collectives_before(net, folded)
            r = sp.play_moves(timed_moves)                      # search + policy read-back + sample assembly (collect=True)
            torch.cuda.synchronize()
            t_pack0 = time.perf_counter()
            rows = sp.packed_rows(timed_moves)
            gathered = par.allgather_samples(torch.from_numpy(rows), device="cuda") if dist is not None else torch.from_numpy(rows)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            stats = sp.mcts.tree_stats(256)
        secs = allmax(t1 - t0)
        moves = allsum(float(r["moves"]))
        sims_per_s = moves * sims / secs
        e_bar, d_bar = stats["edges_per_node"], stats["select_depth"]
        bytes_per_sim = d_bar * (16 * e_bar + 16) + (16 * e_bar + 384) + 2 * 16128 + 3168
        out = {"workload": tag, "games_this_rank": n, "sims_per_move": sims,
               "moves_per_s": moves / secs, "sims_per_s": sims_per_s, "moves": int(moves), "seconds": secs,
               "seconds_search_and_samples": allmax(r["seconds"]), "seconds_weight_broadcast": allmax(t_bcast),
               "seconds_pack_and_allgather": allmax(t1 - t_pack0), "weights_broadcast_bytes": int(nbytes),
               "samples_allgathered": int(gathered.shape[0]), "sample_row_bytes": int(rows.shape[1]) if rows.ndim == 2 else 0,
               "waves": int(r["waves"]), "tensor_util": sims_per_s * 6.56e9 / (tf_peak * 1e12 * world),
               "search_hbm": {"edges_per_node": e_bar, "select_depth": d_bar, "bytes_per_sim": bytes_per_sim,
                              "achieved_gbs_per_gpu": sims_per_s / world * bytes_per_sim / 1e9,
                              "frac_of_hbm_peak": sims_per_s / world * bytes_per_sim / 1e9 / hbm_peak}}
        if r_open is not None:
            om, osec = allsum(float(r_open["moves"])), allmax(r_open["seconds"])
            out["opening_moves_per_s"] = om / osec
            out["full_game_moves_per_s_estimate"] = 54.0 / (5.0 / (om / osec) + 49.0 / (moves / secs))
        del sp
        return out

    res = {"tensor_peak_tflops": tf_peak, "tensor_peak_source": tf_src,
           "net": "BN-folded bf16, every kernel of a wave hand-written sm_100a: 39 trunk 3x3 convs = tcgen05 implicit GEMM on CTA pairs (cta_group::2, M = 256; TMA halo "
                  "tile, TMEM accumulators, fused bias/residual/ReLU); heads = two tcgen05 GEMMs (both 1x1 convs as one, policy fc) "
                  "+ softmax / value-MLP kernel writing into the search's leaf arenas (net_forward)"}
    c2 = selfplay_config("configs[2]: AlphaZero self-play, %d sims/move, model_hive net random-init, %d concurrent games per GPU"
                         % (args.selfplay_sims, n2), n2, args.selfplay_sims, 0, 5, args.selfplay_moves)
    res.update(c2)                                              # configs[2] stays at the top level of the block (as in round 1)
    res["trunk_kernel_launches"] = int(folded.trunk.launches)
    if not args.no_selfplay_deep:
        res["config3"] = selfplay_config("configs[3]: self-play %d sims/move, %d games sharded over %d B200, NCCL weight broadcast + "
                                         "sample all-gather" % (args.sharded_sims, args.sharded_games, world),
                                         n3, args.sharded_sims, 7, 0, 1)
        # configs[4]: evaluator match, new vs best net
        with torch.cuda.stream(stream):
            ev = hive_b200.EvaluatorMatch(n4, args.evaluator_sims, hive_b200.LeafEvaluator(folded),
                                          hive_b200.LeafEvaluator(folded_best, stream=side_stream),
                                          device=local_rank, stream=stream.cuda_stream, seed=7 + rank, torch_stream=stream)
            ev.play(max_plies=5)                                # 4 random plies + one searched ply (warm-up, graph capture)
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            t_bcast, nbytes = timed_collectives_before(net, folded)       # the candidate's weights arrive
            w0 = ev.waves
            r = ev.play(max_plies=2)                            # two searched plies (one by each colour)
            torch.cuda.synchronize()
            tally = torch.tensor([r["new_wins"], r["best_wins"], r["draws"]], dtype=torch.int64, device="cuda")
            if dist is not None:
                dist.all_reduce(tally)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        secs, moves = allmax(t1 - t0), allsum(2.0 * n4)
        res["config4"] = {"workload": "configs[4]: evaluator match, %d sims/move, new vs best net, %d games per GPU on %d B200"
                                      % (args.evaluator_sims, n4, world),
                          "moves_per_s": moves / secs, "sims_per_s": moves * args.evaluator_sims / secs, "searched_moves": int(moves),
                          "seconds": secs, "seconds_weight_broadcast": allmax(t_bcast), "weights_broadcast_bytes": int(nbytes),
                          "waves": int(ev.waves - w0), "tensor_util": moves * args.evaluator_sims / secs * 6.56e9 / (tf_peak * 1e12 * world),
                          "tally_after_two_plies": [int(x) for x in tally.tolist()]}
    # the reference's sequential CPU search next to it (rank 0, N = 1 only; bounded sample)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = max(1, min(os.cpu_count() or 1, 16))
        rate, moves, dt, wall = cpu_mcts_rate(8.0, procs, args.selfplay_sims)
        res["cpu_baseline"] = {"value": rate, "unit": "moves/s", "cores": procs, "kind": "port", "sims_per_move": args.selfplay_sims,
                               "sample": "%d sequential %d-simulation searches (oracle/mcts_oracle.py, the restatement of "
                                         "HivePlayer.action, hash-net stand-in for the network, no GPU) on %d processes in %.1f s"
                                         % (moves, args.selfplay_sims, procs, dt)}
    return res


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import hive_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    try:
        host_cores = len(os.sched_getaffinity(0))       # the ranks of one box share its host cores
    except AttributeError:
        host_cores = os.cpu_count() or 1
    hive_b200.build()
    n = args.games
    stream = torch.cuda.Stream()
    batch = hive_b200.HiveBatch(n, device=local_rank, stream=stream.cuda_stream)
    seed = args.seed + 1000003 * rank            # independent games per rank (sharded, no data-path collective)

    # ------------------------------------------------------------------ resident (device-timed)
    # The timed region is at least --min-window-ms long whatever --steps says: the K-step block (one CUDA-graph launch of
    # min(chunk, K) steps, then single steps up to K) is repeated R times between the two events, and everything is
    # reported per env step (ms_per_step = window / (K * R); config.timed_steps = K * R).  A 20-step window is 1.5 ms:
    # one graph-launch latency and a max over ranks of eight such windows would be most of what it measures.
    for _ in range(max(args.warmup, 3)):
        batch.step_random(seed, args.max_turn, True)
    full = min(args.chunk, args.steps)                  # the multi-step graph captured during warm-up
    if full > 1:
        batch.step_random_multi(seed, full, args.max_turn, True)      # graph capture + one replay, untimed

    def k_steps():                                      # EXACTLY args.steps steps: graph launches of `full` steps, then single steps
        done_steps = 0
        while done_steps < args.steps:
            if full > 1 and args.steps - done_steps >= full:
                batch.step_random_multi(seed, full, args.max_turn, True)
                done_steps += full
            else:                                       # (a shorter multi-step graph would be captured inside the timed region)
                batch.step_random(seed, args.max_turn, True)
                done_steps += 1

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    batch.sync()
    ev0.record(stream)
    k_steps()                                           # pilot block (untimed for the result): sizes R
    ev1.record(stream)
    torch.cuda.synchronize()
    pilot_ms = max(allmax(ev0.elapsed_time(ev1)), 1e-3)
    repeats = max(1, int(np.ceil(args.min_window_ms / pilot_ms)))
    steps0 = int(batch.counters()[0].astype(np.int64).sum())
    launches0 = batch.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(repeats):
        k_steps()
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    timed_steps = args.steps * repeats
    launches = batch.launches - launches0
    env_steps = int(batch.counters()[0].astype(np.int64).sum()) - steps0
    ms_max = allmax(ms)
    env_steps_all = allsum(float(env_steps))
    value = env_steps_all / (ms_max * 1e-3)

    # roofline, this rank.  One "launch" of the hot path = one step = step kernel + plane store over
    # the whole batch (cut into slices inside one CUDA graph); algorithmic bytes as in SURVEY 8d / DESIGN.md.
    peak, peak_src = measured_peak()
    n_steps = max(timed_steps, 1)
    per_launch_bytes = BYTES_PER_ENV_STEP * env_steps / n_steps
    launch_s = ms * 1e-3 / n_steps
    achieved = per_launch_bytes / launch_s / 1e9
    # the dominant kernel alone (hive_planes_kernel: reads 1,120 B of bit planes, writes 16,128 B of bf16 planes per
    # game through the TMA): timed live with events around the two kernels of un-sliced steps
    prof = [batch.profile_step(seed, args.max_turn) for _ in range(12)][2:]
    kms = {k: sum(p[k] for p in prof) / len(prof) for k in prof[0]}
    PLANES_BYTES = 16128 + 1120
    enc_bytes = PLANES_BYTES * (env_steps / n_steps)
    dominant = {"kernel": "hive_planes_kernel", "share_of_step": kms["planes"] / sum(kms.values()), "launch_us": kms["planes"] * 1e3,
                "algorithmic_bytes_per_env_step": PLANES_BYTES, "achieved": enc_bytes / (kms["planes"] * 1e-3) / 1e9,
                "per_kernel_us": {k: v * 1e3 for k, v in kms.items()}}
    dominant["frac"] = dominant["achieved"] / peak
    # the step is ~96 % writes (16,128 + 200 of 17,094 B): next to the contract's copy figure, report what a
    # write-only stream over the same planes arena reaches on this GPU (own 16-byte-store kernel, nothing read)
    write_only = batch.probe_write_stream(20)
    batch.step_random(seed, args.max_turn, True)       # planes hold real data again
    batch.sync()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "write_only_stream_gbs": write_only, "frac_of_write_only_stream": achieved / write_only,
                "traffic": ncu_traffic(), "kernel": "env step = hive_step_kernel + hive_planes_kernel",
                "peak_source": peak_src, "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP,
                "env_steps_per_launch": env_steps / n_steps, "launch_us": launch_s * 1e6,
                "kernels_per_step": launches / n_steps, "dominant_kernel": dominant}

    # ------------------------------------------------------------------ e2e (host buffers)
    # The same 16,384 games driven from the host through the C ABI (hive_host_loop_*): every step of a part copies its
    # actions H2D from pinned memory, runs the kernels and reads legal masks, counts and status back D2H; the policy
    # (the host twin of the device policy) runs on native driver threads, each walking over its own parts while the GPU
    # steps the others.  No Python inside the timed loop.
    # host threads of this rank: one driver thread per part, the rest is the policy pool the drivers share
    e2e_budget = int(os.environ.get("HIVE_B200_E2E_THREADS", str(max(2, min(16, host_cores // max(world, 1))))))
    # four parts in flight whatever the thread budget; with few host threads a driver walks over several parts (measured
    # with 4 host threads: 4 parts / 2 drivers 103 M env-steps/s, 2 parts / 2 drivers 84 M; with 16: 4 / 4 148 M)
    e2e_parts = args.e2e_parts if args.e2e_parts > 0 else 4
    e2e_drivers = min(e2e_parts, args.e2e_drivers) if args.e2e_drivers > 0 else max(1, min(e2e_parts, e2e_budget // 2 if e2e_budget < 8 else 4))
    os.environ["HIVE_B200_HOST_THREADS"] = str(max(1, e2e_budget - e2e_drivers) + 1)    # pool workers + the calling driver
    loop = hive_b200.HostLoop(n, device=local_rank, parts=e2e_parts, threads=e2e_drivers)
    loop.run(8, seed=seed, max_turn=args.max_turn)                                   # builds the per-part graphs
    pilot = loop.run(20, seed=seed, max_turn=args.max_turn)
    k_e2e = max(args.steps, int(np.ceil(20 * args.min_window_ms * 1e-3 / max(pilot["seconds"], 1e-6))))
    s0 = loop.env_steps()
    barrier()
    torch.cuda.synchronize()
    r_e2e = loop.run(k_e2e, seed=seed, max_turn=args.max_turn)
    torch.cuda.synchronize()
    barrier()
    e2e_steps = loop.env_steps() - s0
    dt = r_e2e["seconds"]
    e2e_value = allsum(float(e2e_steps)) / allmax(dt)
    busy = r_e2e["policy_seconds"] / max(dt, 1e-9)
    use_lists = os.environ.get("HIVE_B200_HOST_LISTS", "1") != "0"
    d2h_per_game = (96 + 4) if use_lists else (200 + 4 + 4)         # compact legal lists (96 B per game) + status, or masks + counts + status
    pcie_gbs = (d2h_per_game + 4) * e2e_steps / max(dt, 1e-9) / 1e9
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": d2h_per_game * n,
           "legal_sets_as": "compact lists (hive_step_host_async_lists: 96 B per game)" if use_lists else "masks (208 B per game)",
           "steps": k_e2e, "parts": loop.parts, "host_threads_per_rank": e2e_budget, "driver_threads": loop.threads, "host_cores": host_cores,
           "policy_share_of_thread_time": busy, "wait_share_of_thread_time": r_e2e["wait_seconds"] / max(dt, 1e-9),
           "pcie_gbs_this_rank": pcie_gbs,
           "bound": "host threads" if busy > 0.6 else ("PCIe" if pcie_gbs > 40.0 else "GPU step latency of a part"),
           "note": "per-GPU bytes per step of all %d games; %d parts of the batch, each its own environment handle, streams and pinned "
           "buffers (one CUDA-graph launch per part and step: H2D actions, kernels, D2H of the new legal sets and status); %d native driver "
           "threads run the policy (C-ABI twin of the device policy) for their parts while the GPU steps the others; planes stay in "
           "HBM for the net" % (n, loop.parts, loop.threads)}
    loop.close()
    clocks = sampler.stop()                       # sampled over both timed regions (resident rollout + host-driven e2e)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, games, steps_cpu, dt_cpu = cpu_port_rate(args.seed, 12.0, threads, args.max_turn)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d full random games (same counter-based policy, from reset to terminal/turn %d) = "
                         "%d env steps in %.1f s on %d threads" % (games, args.max_turn, steps_cpu, dt_cpu, threads)}

    # ------------------------------------------------------------------ self-play side measurement
    selfplay = None
    if not args.no_selfplay:
        selfplay = run_selfplay(args, hive_b200, torch, dist, rank, world, local_rank, allsum, allmax, barrier)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / max(timed_steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic (random-legal-move games from reset, "
            "counter-based splitmix64 policy; planes emitted as bf16)",
            "config": {"workload": WORKLOAD, "games_per_gpu": n, "max_turn": args.max_turn, "auto_reset": True,
                       "steps_per_graph_launch": full, "timed_steps": timed_steps, "repeats_of_steps_block": repeats,
                       "timed_window_ms": ms_max,
                       "parallelism": "games sharded %d-way, no data-path collective" % world,
                       "cache": "per-GPU working set %.0f MB (state+legal+planes) > 126 MB L2: inputs larger than L2"
                                % (n * (384 + 200 + 8 + 16128) / 1e6)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "selfplay": selfplay,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
