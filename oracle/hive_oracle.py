"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/hive_oracle.c (the CPU
checker).  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by hive-alphazero_b200/.

Parity pinned: tests/golden/*.npz (vectors produced by the real reference, see
oracle/gen_golden.py) and the container-only live differential test.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhive_oracle.so")
_lib = None


def build(force=False):
    """gcc-compile hive_oracle.c -> libhive_oracle.so (idempotent)."""
    src = os.path.join(_HERE, "hive_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libhive_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, u64, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_long
        L.ho_sizeof.restype = i32
        L.ho_reset.argtypes = [vp]
        L.ho_load.argtypes = [vp, i32, vp, vp]
        L.ho_load.restype = i32
        L.ho_move.argtypes = [vp, i32]
        L.ho_legal.argtypes = [vp, vp]
        L.ho_legal.restype = i32
        L.ho_planes.argtypes = [vp, vp]
        L.ho_game_is_over.argtypes = [vp]
        L.ho_game_is_over.restype = i32
        L.ho_turn.argtypes = [vp]
        L.ho_turn.restype = i32
        L.ho_winner.argtypes = [vp]
        L.ho_winner.restype = i32
        L.ho_position.argtypes = [vp, vp, vp]
        L.ho_state_key.argtypes = [vp, ctypes.c_char_p]
        L.ho_state_key.restype = i32
        L.ho_neighbors.argtypes = [vp]
        L.ho_splitmix64.argtypes = [u64]
        L.ho_splitmix64.restype = u64
        L.ho_pick_action.argtypes = [vp, u64, u64]
        L.ho_pick_action.restype = i32
        L.ho_play_game.argtypes = [vp, u64, u64, i32, vp]
        L.ho_play_game.restype = i32
        L.ho_play_many.argtypes = [u64, cl, cl, i32, i32]
        L.ho_play_many.restype = cl
        _lib = L
    return _lib


class OracleEnv:
    """Single-game CPU checker with the reference's GamePlay vocabulary."""

    def __init__(self):
        self._L = lib()
        self._buf = ctypes.create_string_buffer(self._L.ho_sizeof())
        self._p = ctypes.addressof(self._buf)
        self._L.ho_reset(self._p)

    def clone(self):
        o = OracleEnv.__new__(OracleEnv)
        o._L = self._L
        o._buf = ctypes.create_string_buffer(self._buf.raw, len(self._buf))
        o._p = ctypes.addressof(o._buf)
        return o

    def reset(self):
        self._L.ho_reset(self._p)

    def load(self, turn, cells, levels):
        c = np.ascontiguousarray(cells, dtype=np.uint8)
        l = np.ascontiguousarray(levels, dtype=np.uint8)
        if self._L.ho_load(self._p, int(turn), c.ctypes.data, l.ctypes.data) != 0:
            raise ValueError("inconsistent stacks")

    def move(self, action):
        self._L.ho_move(self._p, int(action))

    def actions(self):
        out = np.empty(1584, dtype=np.int32)
        n = self._L.ho_legal(self._p, out.ctypes.data)
        return out[:n].copy()

    def planes(self):
        """(56,144) uint8, CHW, plane 31 = turn."""
        out = np.empty((56, 144), dtype=np.uint8)
        self._L.ho_planes(self._p, out.ctypes.data)
        return out

    def encode_board(self):
        """(12,12,56) float64 like GamePlay.encode_board (env_hive.py:306)."""
        return self.planes().reshape(56, 12, 12).transpose(1, 2, 0).astype(np.float64)

    def game_is_over(self):
        return bool(self._L.ho_game_is_over(self._p))

    @property
    def turn(self):
        return self._L.ho_turn(self._p)

    @property
    def winner(self):
        return self._L.ho_winner(self._p)

    def position(self):
        c = np.empty(22, dtype=np.uint8)
        l = np.empty(22, dtype=np.uint8)
        self._L.ho_position(self._p, c.ctypes.data, l.ctypes.data)
        return self.turn, c, l

    @property
    def state_key(self):
        buf = ctypes.create_string_buffer(512)
        n = self._L.ho_state_key(self._p, buf)
        return buf.raw[:n].decode()

    def pick_action(self, seed, game_id):
        return self._L.ho_pick_action(self._p, seed, game_id)

    def play_game(self, seed, game_id, max_turn=55):
        tr = np.empty(128, dtype=np.int32)
        n = self._L.ho_play_game(self._p, seed, game_id, max_turn, tr.ctypes.data)
        return tr[:n].copy()


def neighbors():
    out = np.empty((144, 6), dtype=np.int32)
    lib().ho_neighbors(out.ctypes.data)
    return out


def splitmix64(x):
    return lib().ho_splitmix64(ctypes.c_uint64(x & (2 ** 64 - 1)))


def play_many(seed, first, count, max_turn=55, threads=1):
    """Total env steps of `count` counter-seeded games on `threads` host threads."""
    return lib().ho_play_many(seed, first, count, max_turn, threads)
