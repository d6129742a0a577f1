"""Recorded-game ingest for supervised pre-training: replays human / bot game records through
the device environment and turns every recorded move into a training sample, as the reference's
``woker/sl.py::get_buffer`` (:146-231) does one game at a time on the CPU.

A record is the reference's own row format, one row per recorded move (what ``woker/read_sgn_data_file.py`` leaves
in its pickles): ``[piece, x, y, player, bot]`` with piece in {"Q","B1","B2","S1","S2","G1".."G3","A1".."A3"}
(sl.py:33-46), x in ``index_char`` (H..S), y in ``index_number`` ("7".."18"), player "W"/"B", bot 0/1.

  get_buffer(game)        drop-in for sl.py::get_buffer on the ``GamePlay`` facade (one game): same return value
                          ``(data, game)``, data rows ``[planes 12x12x56 nested list, policy[1584], value,
                          [game_len_for_side, counter]]``.
  get_buffers(games)      the B200 form: ALL games advance in lock step on one ``HiveBatch`` (one env step per
                          tick for the whole set; a game whose next recorded mover is not the side to move spends
                          the tick on the skip, sl.py:157-161); per game the same samples, kept compact
                          (bf16 planes, policy as (index, weight)); ``IngestResult.rows(g)`` expands to the
                          reference's rows.

Reference behaviour kept: a row that is not a legal action discards the whole game (``data == []``, the "CCC"
branch sl.py:186-189); the value is taken from the final position only -- +1 / -1 from White's point of view if
the game is over there, else 0 (:205-228); bot moves weigh ``BOT_WEIGHT`` (0.24) in the policy (:194-199);
``[game_lens, counter]`` counts recorded moves per colour (:162-171,220-225).
"""
import numpy as np

from . import config as C
from .env import GamePlay, HiveBatch

PIECE_TYPES = ["G1", "G2", "G3", "A1", "A2", "A3", "S1", "S2", "B1", "B2", "Q"]          # sl.py:46
_SHORT_TO_K = {"Q": 0, "B1": 1, "B2": 2, "S1": 3, "S2": 4, "G1": 5, "G2": 6, "G3": 7, "A1": 8, "A2": 9, "A3": 10}


def decode_piece(piece):
    """sl.py:33-43: record id -> the env's piece key (``str(type) + index``), None for anything else."""
    if piece == "Q":
        return "<class 'pieces.Queen'>0"
    name = {"G": "Grasshopper", "B": "Beetle", "S": "Spider", "A": "Ant"}.get(piece[0])
    return None if name is None else "<class 'pieces.%s'>%d" % (name, int(piece[1]) - 1)


def record_action(row):
    """Action id of one record row: ``board_matrix[x, y]`` x piece index (sl.py:173-183, env_hive.py:287-304).
    Raises ValueError for a coordinate outside the 12x12 window (``list.index``) and IndexError for an unknown
    piece id (``np.where(...)[0][0]`` on an empty match), like the reference."""
    piece, x, y = row[0], row[1], row[2]
    yi = C.index_number.index(y)
    xi = C.index_char.index(x)
    k = _SHORT_TO_K.get(piece) if isinstance(piece, str) else None
    if k is None:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    return (xi * C.MAX_MAP_FULL + yi) * 11 + k


def _value_white(done, winner):
    # sl.py:205-216 (PIECE_WHITE wins -> +1, PIECE_BLACK wins -> -1, otherwise 0)
    if done and winner == 1:
        return 1
    if done and winner == 2:
        return -1
    return 0


def get_buffer(game, env=None, device=0):
    """Drop-in for ``woker/sl.py::get_buffer`` on the device environment (one game through the facade)."""
    board = env if env is not None else GamePlay(HEIGHT_MAP=C.HEIGHT - 100, WIDTH_MAP=C.WIDTH - 500, device=device)
    if env is not None:
        board.new_game()
    state_policy_player = []
    black_count = white_count = 0
    for step in game:
        player, bot = step[3], step[4]
        if (board.player() == 1 and player == "W") or (board.player() == 0 and player == "B"):
            board.skip_turn()
        if player == "W":
            white_count += 1
            counter = white_count
        else:
            black_count += 1
            counter = black_count
        action = record_action(step)
        if action not in board.actions():
            state_policy_player = []
            break
        policy = np.zeros(C.ACTION_SPACE)
        policy[action] = C.BOT_WEIGHT if bot == 1 else 1
        state = board.encode_board(player)
        state_policy_player.append([state.tolist(), policy, player, counter])
        board.move(action, with_skip=False)
    done = board.game_is_over()
    w = board.state.winner
    value_white = _value_white(done, 1 if w == C.PIECE_WHITE else 2 if w == C.PIECE_BLACK else 0)
    data = []
    for state, policy, player, counter in state_policy_player:
        value = value_white if player == "W" else -value_white
        game_lens = white_count if player == "W" else black_count
        if value_white == 0:
            value = 0
        data.append([state, policy.tolist(), value, [game_lens, counter]])
    return data, game


class IngestResult:
    """Samples of a set of replayed games, compact.  Per game g: ``planes[g]`` uint16 (bf16 bits) [m,56,144],
    ``policy_index[g]`` int32 [m], ``policy_weight[g]`` float64 [m], ``value[g]`` int32 [m], ``lens[g]`` int32 [m,2],
    ``discarded[g]`` (an illegal row emptied the game), ``ticks`` env steps the batch took."""

    def __init__(self, n):
        self.planes = [[] for _ in range(n)]
        self.policy_index = [[] for _ in range(n)]
        self.policy_weight = [[] for _ in range(n)]
        self.player = [[] for _ in range(n)]
        self.counter = [[] for _ in range(n)]
        self.value = [None] * n
        self.lens = [None] * n
        self.discarded = [False] * n
        self.ticks = 0

    def n_samples(self):
        return sum(len(p) for p in self.planes)

    def rows(self, g):
        """The reference's rows for game g (what sl.py::get_buffer returns as ``data``)."""
        out = []
        for i in range(len(self.planes[g])):
            f = (np.asarray(self.planes[g][i], dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)
            hwc = f.reshape(C.STATE_FEATURES, 12, 12).transpose(1, 2, 0).astype(np.float64)
            policy = np.zeros(C.ACTION_SPACE)
            policy[self.policy_index[g][i]] = self.policy_weight[g][i]
            out.append([hwc.tolist(), policy.tolist(), int(self.value[g][i]), [int(self.lens[g][i][0]), int(self.lens[g][i][1])]])
        return out


def get_buffers(games, device=0, stream=None, batch=None):
    """Replays ``games`` (a list of records) in lock step on one HiveBatch and returns an IngestResult holding,
    per game, exactly the samples ``get_buffer`` returns for it."""
    n = len(games)
    res = IngestResult(n)
    if n == 0:
        return res
    hb = batch if batch is not None else HiveBatch(n, device=device, stream=stream)
    if batch is not None:
        if hb.n != n:
            raise ValueError("batch holds %d games, %d records given" % (hb.n, n))
        hb.reset()
    cursor = np.zeros(n, dtype=np.int64)
    length = np.array([len(g) for g in games], dtype=np.int64)
    white_count = np.zeros(n, dtype=np.int64)
    black_count = np.zeros(n, dtype=np.int64)
    active = length > 0
    actions = np.empty(n, dtype=np.int32)
    while active.any():
        turn, _, _ = hb.status()
        mask, _ = hb.legal_mask()
        planes = None
        actions[:] = C.NOOP
        for g in np.nonzero(active)[0]:
            row = games[g][cursor[g]]
            side = (int(turn[g]) + 1) % 2                            # game_state.py:58-62
            mover = 0 if row[3] == "W" else 1
            if side != mover:
                actions[g] = -1                                      # skip_turn (sl.py:157-161); the row is handled next tick
                continue
            if mover == 0:
                white_count[g] += 1
                counter = white_count[g]
            else:
                black_count[g] += 1
                counter = black_count[g]
            a = record_action(row)
            if not (int(mask[g, a >> 6]) >> (a & 63)) & 1:           # "CCC": the game is dropped
                res.discarded[g] = True
                active[g] = False
                continue
            if planes is None:
                planes = hb.planes_bf16()                            # one download per tick, shared by every game
            res.planes[g].append(planes[g].copy())
            res.policy_index[g].append(a)
            res.policy_weight[g].append(C.BOT_WEIGHT if row[4] == 1 else 1.0)
            res.player[g].append(mover)
            res.counter[g].append(int(counter))
            actions[g] = a
            cursor[g] += 1
            if cursor[g] >= length[g]:
                active[g] = False
        hb.step(actions)
        res.ticks += 1
    _, winner, done = hb.status()
    for g in range(n):
        if res.discarded[g]:
            res.planes[g], res.policy_index[g], res.policy_weight[g], res.player[g], res.counter[g] = [], [], [], [], []
        vw = _value_white(bool(done[g]), int(winner[g]))
        m = len(res.planes[g])
        res.value[g] = np.array([(vw if p == 0 else -vw) for p in res.player[g]], dtype=np.int32).reshape(m)
        res.lens[g] = np.array([[white_count[g] if p == 0 else black_count[g], c] for p, c in zip(res.player[g], res.counter[g])],
                               dtype=np.int32).reshape(m, 2)
        res.planes[g] = np.asarray(res.planes[g], dtype=np.uint16).reshape(m, C.STATE_FEATURES, 144)
        res.policy_index[g] = np.asarray(res.policy_index[g], dtype=np.int32)
        res.policy_weight[g] = np.asarray(res.policy_weight[g], dtype=np.float64)
    if batch is None:
        hb.close()
    return res
