"""TEST INFRASTRUCTURE ONLY -- harness around the *unmodified* Python reference.

Imports the reference's own env/player modules from ``/root/reference`` (present
only in the build container, never on the GPU box) with two inert stubs
(``oracle/stubs``: pygame, matplotlib) and exposes helpers to

* run seeded rollouts (SURVEY.md section 8d "Config 1" procedure),
* extract the full position (turn, per-piece cell/level) from a ``GamePlay``,
* dump legal lists / planes / state keys / status per ply.

It is used by ``oracle/gen_golden.py`` to produce the fixtures committed under
``tests/golden/`` and by the container-only tests that pin the C restatement
(``oracle/hive_oracle.c``) against the real reference.  Nothing under
``hive-alphazero_b200/`` imports this file.

Reference entry points driven here:
  hive_engine/env_hive.py:24   GamePlay
  hive_engine/env_hive.py:99   GamePlay.move
  hive_engine/env_hive.py:182  GamePlay.actions
  hive_engine/env_hive.py:306  GamePlay.encode_board
  move_checker.py:140          game_is_over
  woker/solo_play.py:69        HivePlayer
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np

REF_ROOT = os.environ.get("HIVE_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")

PIECE_KEYS = ["Q0", "B0", "B1", "S0", "S1", "G0", "G1", "G2", "A0", "A1", "A2"]


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "hive_engine", "env_hive.py"))


_loaded = {}


def load():
    """Import the reference modules (idempotent). Returns a dict of handles."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for p in (REF_ROOT, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, _STUBS)
    from hive_engine.env_hive import GamePlay  # noqa
    from settings import HEIGHT, WIDTH, PIECE_WHITE, PIECE_BLACK  # noqa
    import move_checker  # noqa
    _loaded.update(GamePlay=GamePlay, HEIGHT=HEIGHT, WIDTH=WIDTH,
                   PIECE_WHITE=PIECE_WHITE, PIECE_BLACK=PIECE_BLACK,
                   move_checker=move_checker)
    return _loaded


def new_env():
    """GamePlay(HEIGHT-100, WIDTH-500) exactly as woker/self_play.py:117 does."""
    m = load()
    return m["GamePlay"](m["HEIGHT"] - 100, m["WIDTH"] - 500)


def load_player():
    """Import woker/solo_play.py::HivePlayer (needs torch importable)."""
    load()
    from woker import solo_play  # noqa
    return solo_play


def position(env):
    """(turn, cells[22], levels[22]) -- white pieces 0..10 then black 0..10, in the
    fixed order Q,B0,B1,S0,S1,G0,G1,G2,A0,A1,A2 (env_hive.py:71-87).  cell = q*12+r,
    255 = in hand.  ``level`` is the stored stack index (env_hive.py:119,125)."""
    cells = np.full(22, 255, dtype=np.uint8)
    levels = np.zeros(22, dtype=np.uint8)
    for side, pset in enumerate((env.white_pieces_set, env.black_pieces_set)):
        for k, (_, val) in enumerate(pset.items()):
            tile, level = val[0], val[1]
            if tile.axial_coords != (99, 99):
                cells[side * 11 + k] = tile.index_xy[0] * 12 + tile.index_xy[1]
                levels[side * 11 + k] = level
    return int(env.state.turn), cells, levels


def status(env):
    """(done, winner) with winner 0 none / 1 white / 2 black (move_checker.py:140-165)."""
    m = load()
    done = bool(env.game_is_over())
    w = env.state.winner
    winner = 0 if w is None else (1 if w == m["PIECE_WHITE"] else 2)
    return done, winner


def planes_bits(env):
    """encode_board() for the side to move -> (packed 55 binary planes as bits
    [56][144] uint8 with plane 31 zeroed, turn value of plane 31).  Raises if a
    plane holds anything but {0,1} (other than plane 31)."""
    p = np.asarray(env.encode_board())
    assert p.shape == (12, 12, 56)
    chw = p.transpose(2, 0, 1).reshape(56, 144)
    t = chw[31]
    assert np.all(t == t[0])
    rest = np.delete(chw, 31, axis=0)
    assert np.all((rest == 0) | (rest == 1))
    out = chw.astype(np.uint8)
    out[31] = 0
    return np.packbits(out, axis=1, bitorder="little"), int(t[0])


def rollout(seed, record=False, max_turn=55):
    """SURVEY 8d Config 1: rng=RandomState(seed); a = A[rng.randint(len(A))] or -1.
    Returns dict with transcript (+ per-ply records when ``record``)."""
    rng = np.random.RandomState(seed)
    env = new_env()
    transcript, legal_counts, recs = [], [], []
    while True:
        done, winner = status(env)
        if record:
            turn, cells, levels = position(env)
            bits, tval = planes_bits(env)
            recs.append(dict(turn=turn, cells=cells, levels=levels,
                             legal=np.array(env.actions(), dtype=np.int32),
                             planes=bits, plane31=tval, key=env.state_key,
                             done=done, winner=winner))
        if done or env.state.turn >= max_turn:
            break
        acts = env.actions()
        legal_counts.append(len(acts))
        a = int(acts[rng.randint(len(acts))]) if acts else -1
        transcript.append(a)
        env.move(a)
    return dict(seed=seed, transcript=np.array(transcript, dtype=np.int32),
                legal_sum=int(sum(legal_counts)), records=recs,
                final_turn=int(env.state.turn))


def transcript_hash(transcript):
    return hashlib.sha256(np.asarray(transcript, dtype="<i4").tobytes()).hexdigest()[:16]


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


if __name__ == "__main__":
    # Reproduce SURVEY.md Appendix D pins.
    import time
    for s in range(5):
        t0 = time.time()
        r = rollout(s)
        print(s, r["transcript"][:8].tolist(), r["legal_sum"],
              transcript_hash(r["transcript"]), "%.1fs" % (time.time() - t0))


def inject(turn, cells, levels):
    """Build a reference GamePlay holding an arbitrary position (history empty, no pending push):
    pieces are taken from their inventory tiles and stacked bottom-up, then the same tail as
    GamePlay.move (frontier with the turn-2 rule, state_key, pre_actions, make_state_value) is run."""
    env = new_env()
    keys = list(env.white_pieces_set.keys())
    for lvl in range(5):
        for p in range(22):
            if cells[p] == 255 or levels[p] != lvl:
                continue
            pset = env.white_pieces_set if p < 11 else env.black_pieces_set
            key = keys[p % 11]
            inv_tile, _, piece = pset[key]
            inv_tile.remove_piece()
            tile = env.board_matrix[int(cells[p]) // 12, int(cells[p]) % 12]
            assert len(tile.pieces) == lvl
            tile.add_piece(piece)
            pset[key] = [tile, lvl, piece]
    env.state.turn = int(turn)
    env.next_move_tiles = []
    state_key = ""
    for tile in env.state.board_tiles:
        if tile.axial_coords != (99, 99):
            if tile.has_pieces():
                for adj in tile.adjacent_tiles:
                    if not adj.has_pieces() and adj not in env.next_move_tiles:
                        if env.state.turn == 2:
                            if adj.core_index == ('M', '13'):
                                env.next_move_tiles.append(adj)
                        else:
                            env.next_move_tiles.append(adj)
                for piece in tile.pieces:
                    state_key += env.pieces_keys[piece]
            else:
                state_key += "."
    state_key += str(env.state.player())
    env.state_key = state_key
    env.history_white, env.history_black = [], []
    env.add_history = False
    env.encoded_action = env.pre_actions()
    env.state_final = env.make_state_value()
    return env
