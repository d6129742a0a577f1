"""ctypes binding of include/hive_b200.h.  There is deliberately no fallback: if the CUDA
library is missing or no GPU is present, calls fail loudly."""
import ctypes
import os

from ._build import LIB_PATH

_lib = None

ENV_SYMBOLS = [
    "hive_last_error", "hive_abi_version", "hive_create", "hive_destroy", "hive_num_games", "hive_sync",
    "hive_reset", "hive_step_host", "hive_step", "hive_step_host_async", "hive_step_host_async_lists", "hive_host_pick_actions_lists", "hive_wait_results", "hive_step_random", "hive_step_random_multi", "hive_legal_host", "hive_encode_host", "hive_bits_host",
    "hive_status_host", "hive_status_packed_host", "hive_host_pick_actions", "hive_counters_host", "hive_state_key", "hive_load_state", "hive_dump_state",
    "hive_copy_state", "hive_record_host", "hive_dev_state", "hive_dev_legal", "hive_dev_count", "hive_dev_status", "hive_dev_planes",
    "hive_launch_count", "hive_profile_step", "hive_probe_write_stream", "hive_set_timing", "hive_last_kernel_ms",
    "hive_host_loop_create", "hive_host_loop_destroy", "hive_host_loop_parts", "hive_host_loop_threads", "hive_host_loop_part",
    "hive_host_loop_run", "hive_host_loop_env_steps",
]
# void policy(void* user, int part, int first_game, int n, const u64* mask, const i32* count, const u32* status, i32* actions)
POLICY_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                             ctypes.c_void_p, ctypes.c_void_p)
NET_SYMBOLS = ["net_create", "net_destroy", "net_load_conv_host", "net_trunk_forward", "net_launch_count", "net_load_heads_host", "net_forward", "net_load_conv_dev", "net_load_heads_dev"]
MCTS_SYMBOLS = [
    "mcts_create", "mcts_destroy", "mcts_set_params", "mcts_set_root_noise_host", "mcts_begin", "mcts_descend",
    "mcts_expand", "mcts_pending_host", "mcts_dev_leaf_planes", "mcts_dev_leaf_policy", "mcts_dev_leaf_value", "mcts_dev_pending_mask",
    "mcts_leaf_planes_host", "mcts_set_leaf_eval_host", "mcts_policy_host", "mcts_root_stats_host", "mcts_launch_count",
    "mcts_error_host", "mcts_hash_eval_dev", "mcts_stream", "mcts_tree_stats_host",
]


class HiveError(RuntimeError):
    pass


class SearchError(HiveError):
    """A search tree ran out of its node / edge arena or path depth (HIVE_E_SEARCH): the move would come from
    fewer simulations than asked for, so the call fails instead."""


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HiveError("CUDA extension %s is missing -- run __graft_entry__.build() "
                        "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
    L.hive_last_error.restype = ctypes.c_char_p
    L.hive_abi_version.restype = i32
    L.hive_create.argtypes = [i32, i32, vp, ctypes.POINTER(vp)]
    L.hive_destroy.argtypes = [vp]
    L.hive_num_games.argtypes = [vp]
    L.hive_sync.argtypes = [vp]
    L.hive_reset.argtypes = [vp, vp]
    L.hive_step_host.argtypes = [vp, vp]
    L.hive_step.argtypes = [vp, vp]
    L.hive_step_host_async.argtypes = [vp, vp, vp, vp, vp]
    L.hive_step_host_async_lists.argtypes = [vp, vp, vp, vp]
    L.hive_host_pick_actions_lists.argtypes = [i32, vp, vp, vp, u64, i32, vp, vp]
    L.hive_bits_host.argtypes = [vp, vp]
    L.hive_wait_results.argtypes = [vp]
    L.hive_step_random.argtypes = [vp, u64, i32, i32, vp]
    L.hive_step_random_multi.argtypes = [vp, u64, i32, i32, i32]
    L.hive_legal_host.argtypes = [vp, vp, vp]
    L.hive_encode_host.argtypes = [vp, vp]
    L.hive_status_host.argtypes = [vp, vp, vp, vp]
    L.hive_counters_host.argtypes = [vp, vp, vp]
    L.hive_status_packed_host.argtypes = [vp, vp]
    L.hive_host_pick_actions.argtypes = [i32, vp, vp, vp, vp, u64, i32, vp]
    L.hive_state_key.argtypes = [vp, i32, ctypes.c_char_p, i32]
    L.hive_load_state.argtypes = [vp, i32, i32, vp, vp]
    L.hive_dump_state.argtypes = [vp, i32, vp, vp, vp]
    L.hive_copy_state.argtypes = [vp, i32, vp, i32]
    L.hive_record_host.argtypes = [vp, i32, vp]
    for name in ("hive_dev_state", "hive_dev_legal", "hive_dev_count", "hive_dev_status", "hive_dev_planes"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = vp
    L.hive_launch_count.argtypes = [vp]
    L.hive_launch_count.restype = ctypes.c_longlong
    L.hive_set_timing.argtypes = [vp, i32]
    L.hive_profile_step.argtypes = [vp, u64, i32, vp]
    L.hive_probe_write_stream.argtypes = [vp, i32, vp]
    L.hive_last_kernel_ms.argtypes = [vp]
    L.hive_last_kernel_ms.restype = ctypes.c_float
    L.hive_host_loop_create.argtypes = [i32, i32, i32, i32, ctypes.POINTER(vp)]
    L.hive_host_loop_destroy.argtypes = [vp]
    L.hive_host_loop_parts.argtypes = [vp]
    L.hive_host_loop_threads.argtypes = [vp]
    L.hive_host_loop_part.argtypes = [vp, i32, vp]
    L.hive_host_loop_part.restype = vp
    L.hive_host_loop_run.argtypes = [vp, i32, u64, i32, vp, vp, vp, vp, vp]
    L.hive_host_loop_env_steps.argtypes = [vp]
    L.hive_host_loop_env_steps.restype = ctypes.c_longlong
    f64p, dbl = vp, ctypes.c_double
    L.mcts_create.argtypes = [vp, i32, i32, ctypes.POINTER(vp)]
    L.mcts_destroy.argtypes = [vp]
    L.mcts_set_params.argtypes = [vp, i32, i32, u64]
    L.mcts_set_root_noise_host.argtypes = [vp, vp, i32, i32]
    L.mcts_begin.argtypes = [vp, vp]
    L.mcts_descend.argtypes = [vp, vp]
    L.mcts_expand.argtypes = [vp]
    L.mcts_pending_host.argtypes = [vp, vp]
    for name in ("mcts_dev_leaf_planes", "mcts_dev_leaf_policy", "mcts_dev_leaf_value", "mcts_dev_pending_mask"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = vp
    L.mcts_leaf_planes_host.argtypes = [vp, vp, vp]
    L.mcts_set_leaf_eval_host.argtypes = [vp, vp, vp]
    L.mcts_policy_host.argtypes = [vp, vp, vp, vp]
    L.mcts_root_stats_host.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.mcts_launch_count.argtypes = [vp]
    L.mcts_launch_count.restype = ctypes.c_longlong
    L.mcts_error_host.argtypes = [vp, vp]
    L.mcts_hash_eval_dev.argtypes = [vp, vp, vp, vp, i32, u64, vp]
    L.mcts_tree_stats_host.argtypes = [vp, i32, vp]
    L.mcts_stream.argtypes = [vp]
    L.mcts_stream.restype = vp
    L.net_create.argtypes = [i32, vp, i32, ctypes.POINTER(vp)]
    L.net_destroy.argtypes = [vp]
    L.net_load_conv_host.argtypes = [vp, i32, vp, vp, i32]
    L.net_trunk_forward.argtypes = [vp, vp, i32, ctypes.POINTER(vp)]
    L.net_load_heads_host.argtypes = [vp] * 11
    L.net_load_heads_dev.argtypes = [vp] * 11
    L.net_load_conv_dev.argtypes = [vp, i32, vp, vp, i32]
    L.net_forward.argtypes = [vp, vp, i32, vp, vp]
    L.net_launch_count.argtypes = [vp]
    L.net_launch_count.restype = ctypes.c_longlong
    _lib = L
    return L


def check(rc, what=""):
    if rc == -4:
        raise SearchError("%s failed (%d): %s" % (what, rc, lib().hive_last_error().decode()))
    if rc < 0:
        raise HiveError("%s failed (%d): %s" % (what, rc, lib().hive_last_error().decode()))
    return rc
