"""TEST INFRASTRUCTURE: Python front end of the CPU lock-step warp emulator that runs the
product's CUDA device source (hive-alphazero_b200/csrc/*.cuh, included verbatim) without a GPU."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "hive-alphazero_b200", "csrc")
_LIB = os.path.join(_HERE, "libhive_emu.so")
_lib = None

OP_RESET, OP_STEP, OP_EVAL, OP_RANDOM, OP_INIT = 0, 1, 2, 3, 4
NOOP = -2


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("cuda_emu.cpp", "emu_env.cpp", "cuda_emu.h")]
    srcs += [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < newest:
        cpps = [s for s in srcs if s.startswith(_HERE) and s.endswith(".cpp")]
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-o", _LIB] + cpps)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        vp = ctypes.c_void_p
        L.emu_env_run.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_uint64,
                                  ctypes.c_int, ctypes.c_int, vp, ctypes.c_uint64]
        L.emu_env_run.restype = ctypes.c_int
        L.emu_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


class EmuBatch:
    """Host-memory twin of the device arenas of hive_env (same layouts)."""

    def __init__(self, n, sched_seed=1):
        self.n = n
        self.recs = np.zeros((n, 384), dtype=np.uint8)
        self.legal = np.zeros((n, 50), dtype=np.uint32)
        self.count = np.zeros(n, dtype=np.int32)
        self.status = np.zeros(n, dtype=np.uint32)
        self.planes = np.zeros((n, 56 * 144), dtype=np.uint16)
        self.chosen = np.zeros(n, dtype=np.int32)
        self.sched_seed = sched_seed
        self._run(OP_INIT)

    def _run(self, op, actions=None, mask=None, seed=0, max_turn=55, auto_reset=0):
        self.sched_seed += 7919
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = lib().emu_env_run(self.recs.ctypes.data, self.legal.ctypes.data, self.count.ctypes.data,
                               self.status.ctypes.data, self.planes.ctypes.data, self.n, op,
                               None if a is None else a.ctypes.data, None if m is None else m.ctypes.data,
                               seed, max_turn, auto_reset, self.chosen.ctypes.data, self.sched_seed)
        if rc:
            raise RuntimeError("emulator: " + lib().emu_last_error().decode())

    def reset(self, mask=None):
        self._run(OP_RESET, mask=mask)

    def step(self, actions):
        self._run(OP_STEP, actions=actions)

    def step_random(self, seed, max_turn=55, auto_reset=1):
        self._run(OP_RANDOM, seed=seed, max_turn=max_turn, auto_reset=auto_reset)

    def load(self, g, turn, cells, levels):
        self.recs[g, :] = 0
        self.recs[g, 0:22] = cells
        self.recs[g, 22:44] = levels
        self.recs[g, 44] = turn
        mask = np.zeros(self.n, dtype=np.uint8)
        mask[g] = 1
        self._run(OP_EVAL, mask=mask)

    # ---- views
    def turn(self, g): return int(self.recs[g, 44])
    def winner(self, g): return int(self.recs[g, 45])
    def done(self, g): return bool(self.recs[g, 46])
    def cells(self, g): return self.recs[g, 0:22].copy()
    def levels(self, g): return self.recs[g, 22:44].copy()

    def actions(self, g):
        bits = np.unpackbits(self.legal[g].view(np.uint8), bitorder="little")[:1584]
        return np.nonzero(bits)[0].astype(np.int32)

    def planes_u8(self, g):
        """bf16 planes -> (56,144) uint8 values (exact for the small integers involved)."""
        u = self.planes[g].astype(np.uint32) << 16
        return u.view(np.float32).reshape(56, 144).astype(np.uint8)
