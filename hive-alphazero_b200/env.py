"""Host-side mirror of the reference's environment interface on top of the C ABI.

* ``HiveBatch``  -- n independent games on one GPU (the batched API the hot path is built for).
* ``GamePlay``   -- the reference's single-game class (hive_engine/env_hive.py:24), same method
                    names, argument meaning and return types, backed by a 1-game ``HiveBatch``.

Everything that computes (move generation, step, plane encoding, terminal test) runs in the
sm_100a kernels of csrc/; this file only moves bytes and reshapes them.
"""
import copy
import ctypes

import numpy as np

from . import config as C
from ._capi import HiveError, check, lib


def bits_to_planes_bf16(bits):
    """uint32 [..., 56, 5] bit rows (hive_bits_host) -> bf16 bit patterns uint16 [..., 56, 144], the planes they stand for."""
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    lead = bits.shape[:-2]
    cells = np.unpackbits(bits.view(np.uint8).reshape(lead + (C.STATE_FEATURES, 20)), axis=-1, bitorder="little")[..., :144]
    out = cells.astype(np.uint16) * np.uint16(0x3F80)                       # bf16(1.0)
    turn = bits[..., 31, 0].astype(np.float32)
    out[..., 31, :] = (turn.view(np.uint32) >> 16).astype(np.uint16)[..., None]          # plane 31 = the turn number in every cell
    return out


def _bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


class HiveBatch:
    """n_games concurrent games resident in HBM (include/hive_b200.h)."""

    def __init__(self, n_games, device=0, stream=None):
        self.n = int(n_games)
        self.device = int(device)
        h = ctypes.c_void_p()
        check(lib().hive_create(self.n, self.device, stream, ctypes.byref(h)), "hive_create")
        self._h = h

    @classmethod
    def borrowed(cls, handle, device=0):
        """View of a hive_env_t* owned by someone else (a part of a HostLoop): never destroyed from here."""
        b = cls.__new__(cls)
        b._h = ctypes.c_void_p(handle)
        b._owned = False
        b.n = lib().hive_num_games(b._h)
        b.device = int(device)
        return b

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_owned", True):
                lib().hive_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stepping
    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        if m is not None and m.shape != (self.n,):
            raise ValueError("mask must have shape (n_games,)")
        check(lib().hive_reset(self._h, None if m is None else m.ctypes.data), "hive_reset")

    def step(self, actions):
        """GamePlay.move for every game; actions int32[n] in host memory (-1 pass, NOOP skip)."""
        a = np.ascontiguousarray(actions, dtype=np.int32)
        if a.shape != (self.n,):
            raise ValueError("actions must have shape (n_games,)")
        check(lib().hive_step_host(self._h, a.ctypes.data), "hive_step_host")

    def step_ptr(self, host_ptr):
        """Same, from a raw host pointer (e.g. pinned memory)."""
        check(lib().hive_step_host(self._h, host_ptr), "hive_step_host")

    def step_async_ptr(self, actions_ptr, mask_ptr, count_ptr, status_ptr):
        """Queue H2D(actions) -> step -> D2H(results) without waiting (pinned host pointers)."""
        check(lib().hive_step_host_async(self._h, actions_ptr, mask_ptr, count_ptr, status_ptr), "hive_step_host_async")

    def wait_results(self):
        """Blocks until the downloads of the last step_async_ptr have landed (the planes may still be in flight)."""
        check(lib().hive_wait_results(self._h), "hive_wait_results")

    def step_device(self, actions_dev_ptr):
        check(lib().hive_step(self._h, actions_dev_ptr), "hive_step")

    def step_random(self, seed, max_turn=C.MAX_GAME_LENGTH, auto_reset=True, chosen_dev_ptr=None):
        check(lib().hive_step_random(self._h, seed, max_turn, 1 if auto_reset else 0, chosen_dev_ptr),
              "hive_step_random")

    def step_random_multi(self, seed, n_steps, max_turn=C.MAX_GAME_LENGTH, auto_reset=True):
        """n_steps rollout steps as one CUDA graph launch."""
        check(lib().hive_step_random_multi(self._h, seed, max_turn, 1 if auto_reset else 0, int(n_steps)),
              "hive_step_random_multi")

    def sync(self):
        check(lib().hive_sync(self._h), "hive_sync")

    # ---- results (host copies)
    def legal_mask(self):
        """(mask uint64[n,25], count int32[n])."""
        mask = np.empty((self.n, 25), dtype=np.uint64)
        count = np.empty(self.n, dtype=np.int32)
        check(lib().hive_legal_host(self._h, mask.ctypes.data, count.ctypes.data), "hive_legal_host")
        return mask, count

    def legal_into(self, mask_ptr, count_ptr):
        check(lib().hive_legal_host(self._h, mask_ptr, count_ptr), "hive_legal_host")

    def actions(self, g=None):
        """Sorted legal action ids (GamePlay.actions) of game g, or a list for all games."""
        mask, _ = self.legal_mask()
        bits = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little")[:, :C.ACTION_SPACE]
        if g is not None:
            return np.nonzero(bits[g])[0].astype(np.int32)
        return [np.nonzero(b)[0].astype(np.int32) for b in bits]

    def planes_bf16(self, out=None):
        """bf16 bit patterns [n,56,144]; `out`: array to fill (a page-locked one makes the download several times faster)."""
        if out is None:
            out = np.empty((self.n, C.STATE_FEATURES, 144), dtype=np.uint16)
        assert out.shape == (self.n, C.STATE_FEATURES, 144) and out.dtype == np.uint16 and out.flags.c_contiguous
        check(lib().hive_encode_host(self._h, out.ctypes.data), "hive_encode_host")
        return out

    def planes_bits(self, out=None):
        """The planes as bits: uint32 [n,56,5] (hive_bits_host; plane 31: word 0 = turn, word 1 = row valid)."""
        if out is None:
            out = np.empty((self.n, C.STATE_FEATURES, 5), dtype=np.uint32)
        assert out.shape == (self.n, C.STATE_FEATURES, 5) and out.dtype == np.uint32 and out.flags.c_contiguous
        check(lib().hive_bits_host(self._h, out.ctypes.data), "hive_bits_host")
        return out

    def planes(self):
        """float32 [n,56,12,12] (CHW, what the net consumes)."""
        return _bf16_to_f32(self.planes_bf16()).reshape(self.n, C.STATE_FEATURES, 12, 12)

    def status(self):
        turn = np.empty(self.n, dtype=np.int32)
        winner = np.empty(self.n, dtype=np.int8)
        done = np.empty(self.n, dtype=np.uint8)
        check(lib().hive_status_host(self._h, turn.ctypes.data, winner.ctypes.data, done.ctypes.data),
              "hive_status_host")
        return turn, winner, done

    def status_packed_into(self, host_ptr):
        check(lib().hive_status_packed_host(self._h, host_ptr), "hive_status_packed_host")

    def counters(self):
        steps = np.empty(self.n, dtype=np.uint32)
        episodes = np.empty(self.n, dtype=np.uint32)
        check(lib().hive_counters_host(self._h, steps.ctypes.data, episodes.ctypes.data), "hive_counters_host")
        return steps, episodes

    def state_key(self, g):
        buf = ctypes.create_string_buffer(256)
        n = check(lib().hive_state_key(self._h, int(g), buf, 256), "hive_state_key")
        return buf.raw[:n].decode()

    def load_state(self, g, turn, cells, levels):
        c = np.ascontiguousarray(cells, dtype=np.uint8)
        l = np.ascontiguousarray(levels, dtype=np.uint8)
        if c.shape != (22,) or l.shape != (22,):
            raise ValueError("cells/levels must have 22 entries")
        check(lib().hive_load_state(self._h, int(g), int(turn), c.ctypes.data, l.ctypes.data), "hive_load_state")

    def record(self, g):
        """The raw 384-byte record of game g (uint8[384]): cells, levels, header, counters, plane history."""
        rec = np.zeros(384, dtype=np.uint8)
        check(lib().hive_record_host(self._h, int(g), rec.ctypes.data), "hive_record_host")
        return rec

    def dump_state(self, g):
        turn = ctypes.c_int32()
        c = np.empty(22, dtype=np.uint8)
        l = np.empty(22, dtype=np.uint8)
        check(lib().hive_dump_state(self._h, int(g), ctypes.byref(turn), c.ctypes.data, l.ctypes.data),
              "hive_dump_state")
        return turn.value, c, l

    def copy_state_from(self, g, other, og):
        check(lib().hive_copy_state(self._h, int(g), other._h, int(og)), "hive_copy_state")

    # ---- device arenas
    @property
    def dev_planes(self): return lib().hive_dev_planes(self._h)
    @property
    def dev_legal(self): return lib().hive_dev_legal(self._h)
    @property
    def dev_count(self): return lib().hive_dev_count(self._h)
    @property
    def dev_state(self): return lib().hive_dev_state(self._h)
    @property
    def dev_status(self): return lib().hive_dev_status(self._h)
    @property
    def launches(self): return lib().hive_launch_count(self._h)

    def profile_step(self, seed, max_turn=C.MAX_GAME_LENGTH):
        """One rollout step timed kernel by kernel: dict(step, planes) in ms."""
        ms = (ctypes.c_float * 2)()
        check(lib().hive_profile_step(self._h, seed, max_turn, ms), "hive_profile_step")
        return dict(zip(("step", "planes"), [float(x) for x in ms]))

    def probe_write_stream(self, reps=20):
        """GB/s of a write-only stream over this batch's planes arena (roofline aid; clobbers the planes)."""
        g = ctypes.c_double()
        check(lib().hive_probe_write_stream(self._h, reps, ctypes.byref(g)), "hive_probe_write_stream")
        return g.value

    def set_timing(self, on): check(lib().hive_set_timing(self._h, 1 if on else 0), "hive_set_timing")
    def last_kernel_ms(self): return float(lib().hive_last_kernel_ms(self._h))


def host_pick_actions(mask, count, packed_status, episodes, seed, max_turn, actions_out):
    """hive_host_pick_actions on numpy arrays (all preallocated, C-contiguous)."""
    n = len(count)
    check(lib().hive_host_pick_actions(n, mask.ctypes.data, count.ctypes.data, packed_status.ctypes.data,
                                       episodes.ctypes.data, seed, max_turn, actions_out.ctypes.data),
          "hive_host_pick_actions")


def host_pick_actions_ptr(n, mask_ptr, count_ptr, status_ptr, episodes_ptr, seed, max_turn, actions_ptr):
    """hive_host_pick_actions on raw host addresses (a tight host loop computes them once: ndarray.ctypes.data costs
    microseconds per access)."""
    check(lib().hive_host_pick_actions(n, mask_ptr, count_ptr, status_ptr, episodes_ptr, seed, max_turn, actions_ptr),
          "hive_host_pick_actions")


class HostLoop:
    """The host-driven game loop of a whole batch inside the library (hive_host_loop_*, include/hive_b200.h): the batch is
    cut into parts, native driver threads wait for a part's legal masks / counts / status, run the policy and launch the
    part's next step while the GPU steps the other parts (woker/self_play.py:54-56,116-193 is the loop this replaces).
    policy=None: the host twin of the on-device random policy; else a callable
    policy(part, first_game, mask u64[n,25], count i32[n], status u32[n], actions i32[n]) writing `actions` in place
    (called from the driver threads; Python callables serialise on the interpreter lock)."""

    def __init__(self, n_games, device=0, parts=8, threads=4):
        from ._capi import POLICY_FN
        self._fn_type = POLICY_FN
        self.n = int(n_games)
        self.device = int(device)
        h = ctypes.c_void_p()
        check(lib().hive_host_loop_create(self.n, int(device), int(parts), int(threads), ctypes.byref(h)), "hive_host_loop_create")
        self._h = h
        self.parts = lib().hive_host_loop_parts(h)
        self.threads = lib().hive_host_loop_threads(h)

    def close(self):
        if getattr(self, "_h", None):
            lib().hive_host_loop_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def env_steps(self):
        return int(lib().hive_host_loop_env_steps(self._h))

    def part(self, i):
        """(borrowed HiveBatch view of part i, index of its first game in the batch)"""
        first = ctypes.c_int()
        eh = lib().hive_host_loop_part(self._h, int(i), ctypes.byref(first))
        if not eh:
            raise IndexError(i)
        return HiveBatch.borrowed(eh, self.device), first.value

    def run(self, n_steps, seed=0, max_turn=C.MAX_GAME_LENGTH, policy=None):
        """n_steps steps of every part.  Returns dict(seconds, policy_seconds, wait_seconds)."""
        cb = None
        if policy is not None:
            def tramp(user, part, first, n, mask, count, status, actions):
                m = np.ctypeslib.as_array(ctypes.cast(mask, ctypes.POINTER(ctypes.c_uint64)), shape=(n, 25))
                c = np.ctypeslib.as_array(ctypes.cast(count, ctypes.POINTER(ctypes.c_int32)), shape=(n,))
                st = np.ctypeslib.as_array(ctypes.cast(status, ctypes.POINTER(ctypes.c_uint32)), shape=(n,))
                a = np.ctypeslib.as_array(ctypes.cast(actions, ctypes.POINTER(ctypes.c_int32)), shape=(n,))
                policy(part, first, m, c, st, a)
            cb = self._fn_type(tramp)
        sec, pol, wait = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        check(lib().hive_host_loop_run(self._h, int(n_steps), int(seed), int(max_turn), cb, None,
                                       ctypes.byref(sec), ctypes.byref(pol), ctypes.byref(wait)), "hive_host_loop_run")
        return {"seconds": sec.value, "policy_seconds": pol.value, "wait_seconds": wait.value}


class _State:
    """The slice of Game_State (game_state.py:10-119) the hot-path callers read.  `turn` is cached by the facade after
    every step (one status read per move, not one per access)."""

    def __init__(self, env):
        self._env = env
        self.winner = None
        self.turn = 1

    def player(self):          # game_state.py:58-62
        return 0 if self.turn % 2 == 1 else 1


class _TileView:
    """What callers read of a reference Tile (tile.py:10-30): board coordinates and the printable label."""

    def __init__(self, q, r):
        self.index_xy = [q, r]
        self.core_index = (C.index_char[q], C.index_number[r])
        self.axial_coords = (q, r)

    def __repr__(self):
        return "Tile%s" % (self.core_index,)


_HAND_TILE = type("_HandTile", (), {"index_xy": [99, 99], "core_index": ("-", "-"), "axial_coords": (99, 99),
                                    "__repr__": lambda self: "Tile(hand)"})()


class GamePlay:
    """Drop-in for hive_engine/env_hive.py::GamePlay on the hot path.

    ``debug=True`` makes ``move`` raise ValueError on an action outside ``actions()``; the default
    mirrors the reference, which does not validate (assert commented out, env_hive.py:129-144).
    """

    def __init__(self, HEIGHT_MAP=C.HEIGHT - 100, WIDTH_MAP=C.WIDTH - 500, second_force=False, device=0, debug=False):
        self.HEIGHT_MAP, self.WIDTH_MAP = HEIGHT_MAP, WIDTH_MAP
        self.second_force = True
        self.debug = debug
        self._device = device
        self._batch = HiveBatch(1, device=device)
        self.state = _State(self)
        self._stale_key_player = None
        # board_matrix[q, r] (env_hive.py:33,90): one view object per cell, read by solo_play.py:127-134 for printing
        self.board_matrix = np.empty((C.MAX_MAP_FULL, C.MAX_MAP_FULL), dtype=object)
        for q in range(C.MAX_MAP_FULL):
            for r in range(C.MAX_MAP_FULL):
                self.board_matrix[q, r] = _TileView(q, r)
        self._pushes = [1, 0]                             # history entries pushed per side (reset pushes white's first)
        self._refresh()

    # -- internal
    def _refresh(self):
        """One read of the new position per step: legal list, turn, winner, done; the key is rebuilt on first use."""
        self.encoded_action = self._batch.actions(0).tolist()
        turn, winner, done = self._batch.status()
        self.state.turn = int(turn[0])
        self._winner_code, self._done = int(winner[0]), bool(done[0])
        self._key = None
        self._pos = None

    def _position(self):
        if self._pos is None:
            self._pos = self._batch.dump_state(0)
        return self._pos

    def _pieces_set(self, color):
        """{key: [tile, level, piece type name]} in the reference's order (env_hive.py:71-87; keys "<class 'pieces.Queen'>0" ...)."""
        _, cells, levels = self._position()
        out = {}
        for k in range(11):
            c = int(cells[color * 11 + k])
            tile = _HAND_TILE if c == 255 else self.board_matrix[c // 12, c % 12]
            out["<class 'pieces.%s'>%d" % (C.PIECE_CLASS[k], C.PIECE_NUM[k])] = [tile, int(levels[color * 11 + k]), C.PIECE_CLASS[k]]
        return out

    @property
    def white_pieces_set(self):
        return self._pieces_set(0)

    @property
    def black_pieces_set(self):
        return self._pieces_set(1)

    def _history(self, side):
        """history_white / history_black (env_hive.py:38-39,436-445): newest first, each entry (12,12,2) = the mover's own
        pieces and the opponent's, as pushed when that side was to move; ages without a push yet are absent."""
        rec = self._batch.record(0)
        words = rec[64:64 + 320].view(np.uint32).reshape(2, 4, 2, 5)[side]
        out = []
        for age in range(min(self._pushes[side], 4)):      # the device keeps the four newest; older ones never reach the planes
            bits = np.unpackbits(words[age].view(np.uint8).reshape(2, 20), axis=1, bitorder="little")[:, :144]
            out.append(bits.reshape(2, 12, 12).transpose(1, 2, 0).astype(np.float64))
        return out

    @property
    def history_white(self):
        return self._history(0)

    @property
    def history_black(self):
        return self._history(1)

    def human_play(self):                                 # env_hive.py:185-194
        """The reference re-derives the frontier, the legal list and the planes after a piece was dragged by hand in the
        GUI.  Here a position only changes through move(), whose kernel has already done all of that: the call re-reads
        the device's view."""
        self._refresh()

    def new_game(self):                                   # env_hive.py:61-97
        self._batch.reset()
        self.state.winner = None
        self._stale_key_player = None
        self._pushes = [1, 0]
        self._refresh()

    def move(self, move, with_skip=False):                # env_hive.py:99-171
        move = int(move)
        if self.debug and move != -1 and move not in self.encoded_action:
            raise ValueError("illegal action %d at turn %d" % (move, self.state.turn))
        if move < -1 or move >= C.ACTION_SPACE:
            raise IndexError("action %d out of range" % move)
        self._batch.step(np.array([move], dtype=np.int32))
        self._stale_key_player = None
        self._refresh()
        if move >= 0:                                     # a real move pushes the new mover's boards; a pass does not (env_hive.py:100-103)
            self._pushes[self.state.player()] += 1

    def actions(self):                                    # env_hive.py:182
        return self.encoded_action

    def encode_board(self, player="N"):                   # env_hive.py:306-318
        side = "W" if self.state.player() == 0 else "B"
        if player == "N":
            player = side
        if player != side:
            raise KeyError(player)                        # state_final only holds the side to move
        p = self._batch.planes()[0]                       # (56,12,12) float32
        return np.ascontiguousarray(p.transpose(1, 2, 0)).astype(np.float64)

    def game_is_over(self):                               # move_checker.py:140-165
        if self._winner_code == 1:
            self.state.winner = C.PIECE_WHITE
        elif self._winner_code == 2:
            self.state.winner = C.PIECE_BLACK
        return self._done

    def turn(self):
        return self.state.turn

    def player(self):
        return self.state.player()

    @property
    def state_key(self):                                  # env_hive.py:150-168
        if self._key is None:
            self._key = self._batch.state_key(0)
        key = self._key
        if self._stale_key_player is not None:            # skip_turn leaves the key untouched (:493-496)
            key = key[:-1] + self._stale_key_player
        return key

    def skip_turn(self):                                  # env_hive.py:493-496
        stale = self.state_key[-1]
        self._batch.step(np.array([-1], dtype=np.int32))
        self._stale_key_player = stale
        self._refresh()

    def decode_action(self, action):                      # env_hive.py:498-507
        action = int(action)
        cell, k = divmod(action, 11)
        q, r = divmod(cell, 12)
        return "<class 'pieces.%s'>%d" % (C.PIECE_CLASS[k], C.PIECE_NUM[k]), (C.index_char[q], C.index_number[r])

    def position(self):
        """(turn, cells[22], levels[22]) -- see hive_dump_state."""
        return self._position()

    def load_position(self, turn, cells, levels):
        self._batch.load_state(0, turn, cells, levels)
        self._stale_key_player = None
        self._pushes = [0, 0]                             # history is cleared by a load
        self._refresh()

    def __deepcopy__(self, memo):                         # solo_play.py:158 deep-copies the env
        other = GamePlay.__new__(GamePlay)
        other.HEIGHT_MAP, other.WIDTH_MAP = self.HEIGHT_MAP, self.WIDTH_MAP
        other.second_force, other.debug, other._device = self.second_force, self.debug, self._device
        other._batch = HiveBatch(1, device=self._device)
        other._batch.copy_state_from(0, self._batch, 0)
        other._batch.sync()
        other.state = _State(other)
        other.state.winner, other.state.turn = self.state.winner, self.state.turn
        other._stale_key_player = self._stale_key_player
        other.encoded_action = list(self.encoded_action)
        other.board_matrix = self.board_matrix
        other._winner_code, other._done, other._key, other._pos = self._winner_code, self._done, self._key, self._pos
        other._pushes = list(self._pushes)
        return other
