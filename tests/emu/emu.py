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
    srcs = [os.path.join(_HERE, f) for f in ("cuda_emu.cpp", "emu_env.cpp", "emu_mcts.cpp", "cuda_emu.h")]
    srcs += [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < newest:
        cpps = [s for s in srcs if s.startswith(_HERE) and s.endswith(".cpp")]
        defs = os.environ.get("HIVE_EMU_DEFS", "").split()          # e.g. -DHIVE_STEP_WARPS=6: emulate a variant build
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", _LIB] + defs + cpps)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        vp = ctypes.c_void_p
        L.emu_env_run.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_uint64,
                                  ctypes.c_int, ctypes.c_int, vp, ctypes.c_uint64, vp]
        L.emu_env_run.restype = ctypes.c_int
        L.emu_env_rollout.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        L.emu_env_rollout.restype = ctypes.c_int
        L.emu_env_rollout_q.argtypes = L.emu_env_rollout.argtypes
        L.emu_env_rollout_q.restype = ctypes.c_int
        L.emu_env_lists.restype = vp
        L.emu_last_error.restype = ctypes.c_char_p
        L.emu_mcts_create.restype = vp
        L.emu_mcts_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.emu_mcts_destroy.argtypes = [vp]
        L.emu_mcts_set_noise.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
        L.emu_mcts_begin.argtypes = [vp, vp, vp, vp, vp, vp]
        L.emu_mcts_descend.argtypes = [vp, vp]
        L.emu_mcts_planes.restype = vp
        L.emu_mcts_planes.argtypes = [vp]
        L.emu_mcts_pending_mask.restype = vp
        L.emu_mcts_pending_mask.argtypes = [vp]
        L.emu_mcts_set_leaf.argtypes = [vp, vp, vp]
        L.emu_mcts_expand.argtypes = [vp]
        L.emu_mcts_finalize.argtypes = [vp, vp, vp, vp]
        L.emu_mcts_root.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp]
        L.emu_mcts_error.argtypes = [vp]
        L.emu_mcts_error.restype = ctypes.c_uint
        _lib = L
    return _lib


class EmuBatch:
    """Host-memory twin of the device arenas of hive_env (same layouts)."""

    def __init__(self, n, sched_seed=1, full_store=True):
        self.n = n
        # delta plane store (HIVE_B200_DELTA_STORE=1): the bit image of `planes`, kept by the store kernel; None = full store (the default)
        self.shadow = None if full_store else np.zeros((n, 280), dtype=np.uint32)
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
                               seed, max_turn, auto_reset, self.chosen.ctypes.data, self.sched_seed,
                               None if self.shadow is None else self.shadow.ctypes.data)
        if rc:
            raise RuntimeError("emulator: " + lib().emu_last_error().decode())

    def legal_lists(self):
        """The compact legal lists the last step / reset wrote (EnvArgs::lists), decoded: per game the ascending action ids,
        or None where the group's overflow flag is set."""
        nb = (self.n + 31) // 32
        raw = np.ctypeslib.as_array(ctypes.cast(lib().emu_env_lists(), ctypes.POINTER(ctypes.c_uint8)), shape=(nb, 3072)).copy()
        out = []
        for g in range(self.n):
            blk, hdr = raw[g // 32], raw[g // 32][(g % 32) * 12:(g % 32) * 12 + 12]
            if hdr[9] & 1:
                out.append(None)
                continue
            off, cum = int(hdr[0]) | (int(hdr[1]) << 8), [int(x) for x in hdr[2:9]]
            ids, page = [], 0
            for k in range(cum[6]):
                while cum[page] <= k:
                    page += 1
                ids.append(page * 256 + int(blk[32 * 12 + off + k]))
            out.append(ids)
        return out

    def reset(self, mask=None):
        self._run(OP_RESET, mask=mask)

    def step(self, actions):
        self._run(OP_STEP, actions=actions)

    def step_random(self, seed, max_turn=55, auto_reset=1):
        self._run(OP_RANDOM, seed=seed, max_turn=max_turn, auto_reset=auto_reset)

    def step_random_multi(self, seed, n_steps, max_turn=55, auto_reset=1):
        """n_steps random steps through the rollout kernel (one launch; every CTA loops and stores its own planes)."""
        self.sched_seed += 7919
        rc = lib().emu_env_rollout(self.recs.ctypes.data, self.legal.ctypes.data, self.count.ctypes.data, self.status.ctypes.data,
                                   self.planes.ctypes.data, self.n, seed, max_turn, auto_reset, n_steps, self.sched_seed)
        if rc:
            raise RuntimeError("emulator: " + lib().emu_last_error().decode())

    def step_random_queue(self, seed, n_steps, max_turn=55, auto_reset=1):
        """n_steps (1 or 2) random steps through the queue-driven rollout kernels (ticket logic, both bit-plane buffers)."""
        self.sched_seed += 7919
        rc = lib().emu_env_rollout_q(self.recs.ctypes.data, self.legal.ctypes.data, self.count.ctypes.data, self.status.ctypes.data,
                                     self.planes.ctypes.data, self.n, seed, max_turn, auto_reset, n_steps, self.sched_seed)
        if rc:
            raise RuntimeError("emulator: " + lib().emu_last_error().decode())

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


def planes_bf16_to_hwc_f64(bf16_row):
    """bf16 CHW [56*144] -> (12,12,56) float64, the array GamePlay.encode_board returns."""
    f = (bf16_row.astype(np.uint32) << 16).view(np.float32).reshape(56, 12, 12)
    return f.transpose(1, 2, 0).astype(np.float64)


class EmuMcts:
    """Host twin of hive_mcts for the emulator: same wave protocol as the C ABI."""

    def __init__(self, batch, sims, edges_per_sim=96):
        self.batch, self.n, self.sims = batch, batch.n, sims
        self._h = lib().emu_mcts_create(self.n, sims, edges_per_sim)

    def __del__(self):
        try:
            lib().emu_mcts_destroy(self._h)
        except Exception:
            pass

    def set_noise(self, noise):
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        assert noise.ndim == 3 and noise.shape[0] == self.n
        lib().emu_mcts_set_noise(self._h, noise.ctypes.data, noise.shape[1], noise.shape[2])

    def _chk(self, rc):
        if rc:
            raise RuntimeError("emulator: " + lib().emu_last_error().decode())

    def search(self, net):
        """net(planes_hwc_f64) -> (p float32[1584], v float).  Runs all simulations."""
        self._chk(lib().emu_mcts_begin(self._h, self.batch.recs.ctypes.data, self.batch.legal.ctypes.data,
                                       self.batch.count.ctypes.data, self.batch.planes.ctypes.data,
                                       None if self.batch.shadow is None else self.batch.shadow.ctypes.data))
        pending = ctypes.c_int(0)
        p = np.zeros((self.n, 1584), dtype=np.float32)
        v = np.zeros(self.n, dtype=np.float64)
        waves = 0
        while True:
            self._chk(lib().emu_mcts_descend(self._h, ctypes.byref(pending)))
            if pending.value == 0:
                break
            planes = np.ctypeslib.as_array(ctypes.cast(lib().emu_mcts_planes(self._h), ctypes.POINTER(ctypes.c_uint16)),
                                           shape=(self.n, 56 * 144))
            mask = np.ctypeslib.as_array(ctypes.cast(lib().emu_mcts_pending_mask(self._h), ctypes.POINTER(ctypes.c_uint8)),
                                         shape=(self.n,))
            for t in range(self.n):
                if mask[t]:
                    p[t], v[t] = net(planes_bf16_to_hwc_f64(planes[t]))
            lib().emu_mcts_set_leaf(self._h, p.ctypes.data, v.ctypes.data)
            self._chk(lib().emu_mcts_expand(self._h))
            waves += 1
        return waves

    def errors(self):
        """OR of 1 << error code over the trees of the last search (2 node arena, 4 edge arena, 8 depth)."""
        return int(lib().emu_mcts_error(self._h))

    def policy(self):
        pi = np.zeros((self.n, 1584), dtype=np.float64)
        action = np.zeros(self.n, dtype=np.int32)
        sum_n = np.zeros(self.n, dtype=np.int32)
        self._chk(lib().emu_mcts_finalize(self._h, pi.ctypes.data, action.ctypes.data, sum_n.ctypes.data))
        return pi, action, sum_n

    def root_stats(self, t, max_edges=256):
        a = np.zeros(max_edges, dtype=np.int32); n = np.zeros(max_edges, dtype=np.int32)
        w = np.zeros(max_edges, dtype=np.float64); q = np.zeros(max_edges, dtype=np.float64)
        p = np.zeros(max_edges, dtype=np.float32); info = np.zeros(6, dtype=np.int32)
        lib().emu_mcts_root(self._h, t, max_edges, a.ctypes.data, n.ctypes.data, w.ctypes.data, q.ctypes.data,
                            p.ctypes.data, info.ctypes.data)
        k = int(info[0])
        return dict(action=a[:k], n=n[:k], w=w[:k], q=q[:k], p=p[:k], sum_n=int(info[1]), n_nodes=int(info[2]),
                    sims_done=int(info[3]), error=int(info[4]), root_selects=int(info[5]))
