"""Batched self-play on one GPU (reference game loop: woker/self_play_with_train.py:146-219, the
runnable statement of woker/self_play.py:116-193).

Every game of a ``HiveBatch`` plays with its own tree: per move one batched search
(``MctsBatch.search_device`` -> ``LeafEvaluator``), then the reference's opening schedule
(turns with ``error = 0.7 - int(turn+1)/2*0.15 >= 0.1`` sample from ``(1-error)*pi[legal] +
error*Dir(0.5)``), then one environment step.  Value targets follow the reference: +1/-1 from the
winner's side, and -1 for BOTH sides when the game is drawn or cut at MAX_GAME_LENGTH
(self_play_with_train.py:204-216).

For the single-game, RNG-stream-exact drop-in use ``GamePlay`` + ``HivePlayer`` directly.
"""
import datetime
import json
import os
import time

import numpy as np

from . import config as C
from .env import HiveBatch
from .mcts import MctsBatch


class SelfPlayBatch:
    def __init__(self, n_games, sims, evaluator, device=0, stream=None, seed=0, collect=False, edges_per_sim=0, wave_graph=None):
        self.env = HiveBatch(n_games, device=device, stream=stream)
        self.mcts = MctsBatch(self.env, sims, edges_per_sim=edges_per_sim)
        self.mcts.set_root_noise(None)                      # Dirichlet(0.3) rows sampled on the device
        self.mcts.set_params(sims, C.MAX_GAME_LENGTH, noise_seed=seed * 7919 + 17)
        self.evaluator = evaluator
        self.wave_graph = wave_graph                        # optional mcts.WaveGraph: replay captured waves
        self.n, self.sims = n_games, sims
        self.rng = np.random.RandomState(seed)
        self.collect = collect
        self.samples = [[] for _ in range(n_games)] if collect else None
        self.finished = []                                  # (value_white, turn) per finished game
        self.moves = 0
        self.waves = 0
        self.search_calls = 0
        self._pin = None                                    # page-locked landing buffers of the per-move read-backs (made on first use)

    def _pinned(self):
        """Page-locked host buffers for the read-backs of a move (policies 12.7 KB per game; the planes as BITS, 1,120 B per
        game instead of 16 KB: hive_bits_host): the downloads run at the PCIe rate instead of through the driver's staging of
        pageable memory (2,048 games: 15 ms -> 1 ms).  Their content is valid until the next move: samples copy what they keep."""
        if self._pin is None:
            import torch
            self._pin_t = (torch.empty((self.n, C.ACTION_SPACE), dtype=torch.float64).pin_memory(),
                           torch.empty((self.n, C.STATE_FEATURES, 5), dtype=torch.int32).pin_memory(),
                           torch.empty(self.n, dtype=torch.int32).pin_memory(), torch.empty(self.n, dtype=torch.int32).pin_memory())
            pi, bits, action, sum_n = (t.numpy() for t in self._pin_t)
            self._pin = (pi, bits.view(np.uint32), action, sum_n)
        return self._pin

    def _choose(self, pi, mcts_action, legal, turn):
        """The reference's move choice for one game (self_play_with_train.py:169-183)."""
        if len(legal) == 0:
            return -1
        action = int(mcts_action)
        if turn <= 2:
            action = int(self.rng.choice(legal))
        error = 0.7 - int(turn + 1) / 2 * 0.15
        if error >= 0.1:
            p = pi[legal]
            noise = self.rng.dirichlet([0.5] * len(legal))
            p = (1 - error) * p + error * noise
            p = p / p.sum()
            action = int(self.rng.choice(legal, p=p))
        return action

    def _choose_batch(self, pi, mcts_action, legal_bits, count, turn):
        """The reference's move choice (self_play_with_train.py:169-183) for every game at once:
        turns with error >= 0.1 sample from (1-error)*pi[legal] + error*Dir(0.5) (normalised)."""
        n = len(turn)
        chosen = np.asarray(mcts_action, dtype=np.int32).copy()
        chosen[count == 0] = -1
        error = 0.7 - ((turn + 1).astype(np.int64)).astype(np.float64) / 2 * 0.15
        noisy = (error >= 0.1) & (count > 0)
        if noisy.any():
            idx = np.nonzero(noisy)[0]
            rows, cols = np.nonzero(legal_bits[idx])                          # sparse legal entries, sorted by row
            m = len(idx)
            gam = self.rng.gamma(0.5, 1.0, size=len(rows))                    # Dirichlet(0.5) over each row's legal entries
            noise = gam / np.bincount(rows, weights=gam, minlength=m)[rows]
            e = error[idx][rows]
            p = (1 - e) * pi[idx[rows], cols] + e * noise
            p /= np.bincount(rows, weights=p, minlength=m)[rows]
            cum = np.cumsum(p)
            ends = np.cumsum(np.bincount(rows, minlength=m))                   # one past the last entry of each row
            starts = ends - np.bincount(rows, minlength=m)
            base = np.where(starts > 0, cum[np.maximum(starts - 1, 0)], 0.0)
            target = base + self.rng.random_sample(m) * (cum[ends - 1] - base)
            pos = np.minimum(np.maximum(np.searchsorted(cum, target, side="left"), starts), ends - 1)
            pick = cols[pos]
            chosen[idx] = pick
        return chosen

    def play_moves(self, n_moves, restart_finished=True):
        """Advance every live game by n_moves plies. Returns dict(moves, seconds, waves)."""
        t0 = time.perf_counter()
        moves0, waves0 = self.moves, self.waves
        for _ in range(n_moves):
            turn, winner, done = self.env.status()
            over = (done != 0) | (turn >= C.MAX_GAME_LENGTH)
            live = ~over
            self.waves += self.mcts.search_device(self.evaluator, tree_mask=live.astype(np.uint8), graph=self.wave_graph)
            self.search_calls += 1
            # after the opening (error < 0.1 from turn 7 on) the move is the search's own choice: no need to
            # read policies or legal lists back unless samples are being collected
            fast = (not self.collect) and bool((turn[live] > 6).all()) if live.any() else True
            actions = np.full(self.n, C.NOOP, dtype=np.int32)
            if fast:
                mcts_action = self.mcts.actions()
                actions[live] = mcts_action[live]
                for g in np.nonzero(over)[0]:
                    self._finish(int(g), int(winner[g]), int(turn[g]))
                    if restart_finished:
                        actions[g] = -3
            else:
                pin_pi, pin_bits, pin_action, pin_sum = self._pinned()
                pi, mcts_action, _ = self.mcts.policy(out=(pin_pi, pin_action, pin_sum))
                mask, count = self.env.legal_mask()
                legal_bits = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little")[:, :C.ACTION_SPACE].astype(bool)
                bits = self.env.planes_bits(out=pin_bits) if self.collect else None
                chosen = self._choose_batch(pi, mcts_action, legal_bits, count, turn)
                for g in np.nonzero(over)[0]:
                    self._finish(int(g), int(winner[g]), int(turn[g]))
                    if restart_finished:
                        actions[g] = -3                           # HIVE_RESET
                actions[live] = chosen[live]
                if self.collect:
                    # the landing buffers are copied out ONCE per move (contiguous when every game is live); a game's sample
                    # is a view into the copy: (bit rows of its planes uint32 [56,5], pi float32 [1584], side to move).  The planes
                    # themselves (16 KB) are made from the bits when the game is finished (_finish) -- the packed rows the ranks
                    # all-gather are those bits, never the expansion.
                    live_idx = np.nonzero(live)[0]
                    if len(live_idx) == self.n:
                        kept_bits, kept_pi = bits.copy(), pi.astype(np.float32)
                    else:
                        kept_bits, kept_pi = bits[live_idx], pi[live_idx].astype(np.float32)
                    if not (kept_bits[:, 31, 1] == 1).all():
                        raise RuntimeError("self-play: the bit rows of a live game are stale (it was not evaluated by the last step)")
                    side = (turn % 2).tolist()
                    for k, g in enumerate(live_idx.tolist()):
                        self.samples[g].append((kept_bits[k], kept_pi[k], side[g]))
            self.env.step(actions)
            self.moves += int(live.sum())
        self.env.sync()
        return dict(moves=self.moves - moves0, seconds=time.perf_counter() - t0, waves=self.waves - waves0)

    def packed_rows(self, last_moves):
        """The samples of the last `last_moves` plies of every game as fixed 1,960-byte records (parallel.pack_samples:
        55 binary planes + turn byte, sparse policy, value byte) -- what the ranks all-gather.  The value of a game
        still running is not known yet: its records carry 0 (the finished ones are rewritten by `_finish`)."""
        from .parallel import pack_samples
        rows = [smp for per_game in self.samples for smp in per_game[-last_moves:]] if last_moves > 0 else []
        if not rows:
            return np.zeros((0, 1960), dtype=np.uint8)
        bits = np.stack([r[0] for r in rows])                                            # (N, 56, 5) uint32 bit rows
        pi = np.stack([r[1] for r in rows])
        by = np.ascontiguousarray(bits).view(np.uint8).reshape(len(rows), C.STATE_FEATURES, 20)[:, :, :18]    # 144 cells = 18 bytes
        turn = bits[:, 31, 0].astype(np.uint8)
        planes_bits = np.concatenate([np.delete(by, 31, axis=1).reshape(len(rows), 55 * 18), turn[:, None]], axis=1)
        r, c = np.nonzero(pi > 0)                                                        # sparse policy: row-major, ascending actions
        starts = np.searchsorted(r, np.arange(len(rows)))
        pos = np.arange(len(r)) - starts[r]
        keep = pos < 160
        idx = np.full((len(rows), 160), -1, dtype=np.int32)
        val = np.zeros((len(rows), 160), dtype=np.float32)
        idx[r[keep], pos[keep]] = c[keep]
        val[r[keep], pos[keep]] = pi[r[keep], c[keep]]
        return pack_samples(planes_bits, idx, val, np.zeros(len(rows), dtype=np.int8))

    def _finish(self, g, winner, turn):
        value_white = 1 if winner == 1 else (-1 if winner == 2 else 0)
        self.finished.append((value_white, turn))
        if self.collect:
            data = []
            counts = {1: 0, 0: 0}
            for _, _, side_odd in self.samples[g]:
                counts[side_odd] += 1
            seen = {1: 0, 0: 0}
            from .env import bits_to_planes_bf16
            expanded = bits_to_planes_bf16(np.stack([smp[0] for smp in self.samples[g]])) if self.samples[g] else []
            for (_, pi, side_odd), planes in zip(self.samples[g], expanded):       # side_odd == 1: white to move
                seen[side_odd] += 1
                value = value_white if side_odd == 1 else -value_white
                if value_white == 0:
                    value = -1                                 # draw / cut game: -1 for both sides
                data.append((planes, pi, value, (counts[side_odd], seen[side_odd])))
            self.samples[g] = []
            self.finished_samples = getattr(self, "finished_samples", [])
            self.finished_samples.extend(data)


def self_play_buffer(cur=None, make_player=None, device=0):
    """Drop-in for the reference's one-game worker (woker/self_play_with_train.py:146-219, the runnable statement of
    woker/self_play.py:116-193), RNG-stream exact: one GamePlay, two HivePlayers, the opening schedule
    (random move on turns 1-2; while error = 0.7 - int(turn+1)/2*0.15 >= 0.1 sample from (1-error)*pi[legal] +
    error*Dir(0.5)), rows [planes (12,12,56) nested list, pi[1584], value, [game_len_for_side, move_idx_for_side]],
    value = +-1 from the winner's side and -1 for BOTH sides of a drawn / cut game.  Returns (data, [value_white]).
    `cur`: the reference's list of pipe bundles (popped / pushed back, handed to the players; may be None);
    `make_player(pipes)`: builds a configured hive_b200.HivePlayer (default: HivePlayer(pipes=pipes))."""
    from .env import GamePlay
    from .mcts import HivePlayer
    board = GamePlay(HEIGHT_MAP=C.HEIGHT - 100, WIDTH_MAP=C.WIDTH - 500, device=device)
    pipes = cur.pop() if cur else None
    make_player = make_player or (lambda p: HivePlayer(pipes=p))
    white, black = make_player(pipes), make_player(pipes)
    state_policy_player = []
    black_count = white_count = 0
    e = 0.7
    while not board.game_is_over():
        if board.state.player() == 0:
            action, policy = white.action(board)
            player = 'W'
            white_count += 1
            counter = white_count
        else:
            action, policy = black.action(board)
            player = 'B'
            black_count += 1
            counter = black_count
        if board.state.turn <= 2:
            action = np.random.choice(board.actions())
        policy = policy[0]
        error = e - int(board.state.turn + 1) / 2 * 0.15
        actions = board.actions()
        if error >= 0.1 and len(actions) != 0:
            p = np.array(policy)[actions]
            noise = np.random.dirichlet([0.5] * len(actions))
            p = (1 - error) * np.array(p) + error * noise
            p /= p.sum()
            action = np.random.choice(board.actions(), p=p)
        state = board.encode_board(player)
        state_policy_player.append([state.tolist(), policy, player, counter])
        board.move(action)
        if board.state.turn >= C.MAX_GAME_LENGTH:
            break
    value_white = 0
    if board.game_is_over():
        if board.state.winner == C.PIECE_WHITE:
            value_white = 1
        elif board.state.winner == C.PIECE_BLACK:
            value_white = -1
    white.finish_game(value_white)
    black.finish_game(-value_white)
    data = []
    for state, policy, player, counter in state_policy_player:
        value, game_lens = (value_white, white_count) if player == "W" else (-value_white, black_count)
        if value_white == 0:
            value = -1
        data.append([state, policy, value, [game_lens, counter]])
    if cur is not None:
        cur.append(pipes)
    return data, [value_white]


def evaluation_report(win_lose, game_lens, state_keys):
    """What the reference's evaluation worker prints every ten games (woker/evaluation.py:66-88): white win rate =
    share of +1 in win_lose, mean game length, and the share of distinct final positions (de-dup by final state_key)."""
    wl = np.asarray(win_lose)
    values, counts = np.unique(wl, return_counts=True)
    return dict(total_games=int(len(wl)), white_win_rate=float((wl == 1).sum() / max(len(wl), 1)),
                mean_game_len=float(np.round(np.mean(game_lens), 2)) if len(game_lens) else 0.0,
                distinct_final_positions=float(len(np.unique(list(state_keys))) / max(len(wl), 1)),
                counter={int(v): int(c) for v, c in zip(values, counts)})


def accept_new_network(new_wins, best_wins, threshold=0.55):
    """The new-vs-best gate of the evaluator (AlphaZero's rule, hive-report.pdf p.6): the candidate replaces the best
    network when it wins at least `threshold` of the decided games (draws do not count).  (accepted, win_share)"""
    decided = new_wins + best_wins
    share = new_wins / decided if decided else 0.0
    return bool(decided > 0 and share >= threshold), float(share)


def sample_to_reference_row(planes_bf16, pi, value, lens):
    """One finished sample in the reference's on-disk form (self_play.py:160,190):
    [planes 12x12x56 nested list, pi[1584], value, [game_len_for_side, move_idx_for_side]]."""
    f = (np.asarray(planes_bf16, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(C.STATE_FEATURES, 12, 12)
    hwc = f.transpose(1, 2, 0).astype(np.float64)
    return [hwc.tolist(), [float(x) for x in pi], int(value), [int(lens[0]), int(lens[1])]]


def write_play_file(samples, directory="."):
    """Flush samples like SelfPlayWorker.flush_buffer (self_play.py:100-112, sl.py:49-60):
    ``play_%Y%m%d-%H%M%S.%f.json`` holding the list of rows.  Returns the path."""
    os.makedirs(directory, exist_ok=True)
    game_id = datetime.datetime.now().strftime("%Y%m%d-%H%M%S.%f")
    path = os.path.join(directory, "play_%s.json" % game_id)
    with open(path, "wt") as f:
        json.dump([sample_to_reference_row(*s) for s in samples], f)
    return path
