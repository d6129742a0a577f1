"""hive-alphazero_b200 -- B200-native Hive self-play hot path (env step, move generation, plane
encoder, PUCT search) behind the reference's GamePlay / HivePlayer interface.

The package name carries a hyphen; import it as ``import hive_b200`` (root-level alias) or with
``importlib.import_module("hive-alphazero_b200")``.
"""
from . import config
from ._build import LIB_PATH, build
from ._capi import ENV_SYMBOLS, MCTS_SYMBOLS, HiveError, SearchError, lib
from .env import GamePlay, HiveBatch, HostLoop, host_pick_actions, host_pick_actions_ptr
from .mcts import HashEvaluator, HivePlayer, MctsBatch, WaveGraph


def __getattr__(name):
    # torch-dependent parts are imported lazily so that the environment path does not need torch
    if name in ("HiveNet", "FoldedNet", "LeafEvaluator", "SplitEvaluator", "host_net_callable", "device_view"):
        from . import net
        return getattr(net, name)
    if name in ("SelfPlayBatch", "write_play_file", "sample_to_reference_row", "self_play_buffer", "evaluation_report", "accept_new_network"):
        from . import selfplay
        return getattr(selfplay, name)
    if name in ("Trainer", "alpha_loss", "samples_to_tensors", "discounted_value", "load_play_file", "load_play_files", "rows_to_tensors"):
        from . import train
        return getattr(train, name)
    if name in ("get_buffer", "get_buffers", "IngestResult", "record_action", "decode_piece"):
        from . import ingest
        return getattr(ingest, name)
    if name == "EvaluatorMatch":
        from .evaluator import EvaluatorMatch
        return EvaluatorMatch
    raise AttributeError(name)

__all__ = ["config", "build", "lib", "LIB_PATH", "ENV_SYMBOLS", "HiveError", "GamePlay", "HiveBatch", "HostLoop",
           "host_pick_actions", "host_pick_actions_ptr", "HivePlayer", "MctsBatch", "WaveGraph", "HashEvaluator", "SearchError", "MCTS_SYMBOLS"]
