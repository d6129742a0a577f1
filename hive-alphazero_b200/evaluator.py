"""New-vs-best evaluator match (reference: woker/evaluation.py:128-310 and hive-report.pdf p.6;
BASELINE configs[4]): two networks play each other with colours split evenly, the first four plies
are uniformly random, afterwards every side moves by its own PUCT search; wins are tallied per net.
Games are independent, one tree per game, both nets evaluated as batched leaf inference.
"""
import time

import numpy as np

from . import config as C
from .env import HiveBatch
from .mcts import MctsBatch
from .net import SplitEvaluator


class EvaluatorMatch:
    def __init__(self, n_games, sims, eval_new, eval_best, device=0, stream=None, seed=0, random_plies=4, torch_stream=None):
        self.env = HiveBatch(n_games, device=device, stream=stream)
        self.mcts = MctsBatch(self.env, sims)
        self.mcts.set_root_noise(None)
        self.mcts.set_params(sims, C.MAX_GAME_LENGTH, noise_seed=seed * 104729 + 5)
        self.eval_new, self.eval_best = eval_new, eval_best
        self.n, self.sims, self.random_plies = n_games, sims, random_plies
        self.rng = np.random.RandomState(seed)
        self.new_is_white = np.arange(n_games) < n_games // 2          # colours split evenly
        self.moves = 0
        self.waves = 0
        # two wave graphs (white to move / black to move differ in which net evaluates which half)
        self._graphs = None
        if torch_stream is not None:
            from .mcts import WaveGraph
            self._graphs = {True: WaveGraph(torch_stream), False: WaveGraph(torch_stream)}
        self._evals = {True: SplitEvaluator(eval_new, eval_best, n_games // 2), False: SplitEvaluator(eval_best, eval_new, n_games // 2)}

    def play(self, max_plies=C.MAX_GAME_LENGTH):
        """Play every game to the end (or max_plies).  Returns the tally and timing."""
        t0 = time.perf_counter()
        for _ in range(max_plies):
            turn, winner, done = self.env.status()
            over = (done != 0) | (turn >= C.MAX_GAME_LENGTH)
            live = ~over
            if not live.any():
                break
            actions = np.full(self.n, C.NOOP, dtype=np.int32)
            t = int(turn[live][0])                                       # all live games are at the same ply
            if t <= self.random_plies:                                   # evaluation.py:169-170,187-188
                legal = self.env.actions()
                for g in np.nonzero(live)[0]:
                    actions[g] = int(self.rng.choice(legal[g])) if len(legal[g]) else -1
            else:
                # one search for all games: the half where the new net is to move is evaluated by the new
                # net, the other half by the best net (rows are ordered new-is-white first)
                white_to_move = (t % 2) == 1
                g = self._graphs[white_to_move] if self._graphs else None
                self.waves += self.mcts.search_device(self._evals[white_to_move], tree_mask=live.astype(np.uint8), graph=g)
                a = self.mcts.actions()
                actions[live] = a[live]
            self.env.step(actions)
            self.moves += int(live.sum())
        self.env.sync()
        turn, winner, done = self.env.status()
        white_won, black_won = winner == 1, winner == 2
        new_wins = int((white_won & self.new_is_white).sum() + (black_won & ~self.new_is_white).sum())
        best_wins = int((white_won & ~self.new_is_white).sum() + (black_won & self.new_is_white).sum())
        # what the reference's evaluation worker prints every 10 games (evaluation.py:76-88): mean game length and the
        # share of distinct final positions (state_key) among the games played
        keys = [self.env.state_key(g) for g in range(self.n)]
        return dict(games=self.n, new_wins=new_wins, best_wins=best_wins, draws=self.n - new_wins - best_wins,
                    white_win_rate=float(white_won.mean()), mean_game_len=float(np.mean(turn)),
                    distinct_final_positions=len(set(keys)) / self.n,
                    moves=self.moves, waves=self.waves, seconds=time.perf_counter() - t0)
