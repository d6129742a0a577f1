"""Trainer step of the AlphaZero loop (row f2; reference: alpha_zero/alpha_net.py:98-162,
woker/optimize.py:42-65, woker/self_play_with_train.py:118-138).

This closes the loop on one box: samples from ``SelfPlayBatch`` -> discounted value targets ->
``AlphaLoss`` -> Adam step on the fp32 ``HiveNet``.  The math is plain torch autograd (library code):
the hot path this repository accelerates is self-play; training is host plumbing around it.
"""
import numpy as np
import torch

from . import config as C


def alpha_loss(value_pred, value, policy_pred, policy):
    """AlphaLoss (alpha_net.py:98-115): mean over the batch of
    (value - value_pred)^2 * LOSS_WEIGHT['value'] + sum(-policy * log(1e-6 + policy_pred)) * LOSS_WEIGHT['policy']."""
    value_error = (value - value_pred) ** 2
    policy_error = torch.sum(-policy * (1e-6 + policy_pred.float()).float().log(), 1)
    return (value_error.view(-1).float() * C.LOSS_WEIGHT["value"] + policy_error * C.LOSS_WEIGHT["policy"]).mean()


def discounted_value(value, game_len, step):
    """optimize.py:55-58 / self_play_with_train.py:124-127: the last move of a side keeps the raw value,
    earlier ones are discounted by DISCOUNTED_REWARD ** (game_len - step)."""
    return value if step == game_len else value * C.DISCOUNTED_REWARD ** (game_len - step)


def samples_to_tensors(samples, one_hot_policy=False):
    """samples: iterable of (planes bf16 [56*144] or (56,144), pi[1584], value, (game_len, step)) as
    produced by SelfPlayBatch -> (states (B,56,12,12) f32, policies (B,1584) f32, values (B,) f32).
    one_hot_policy reproduces self_play_with_train.py:128-130 (one-hot of the arg-max policy)."""
    xs, ps, vs = [], [], []
    for planes, pi, value, (game_len, step) in samples:
        f = (np.asarray(planes, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(C.STATE_FEATURES, 12, 12)
        pi = np.asarray(pi, dtype=np.float32)
        if one_hot_policy:
            oh = np.zeros_like(pi)
            oh[int(np.argmax(pi))] = 1.0
            pi = oh
        xs.append(f); ps.append(pi); vs.append(discounted_value(float(value), int(game_len), int(step)))
    return (torch.from_numpy(np.stack(xs)), torch.from_numpy(np.stack(ps)), torch.tensor(vs, dtype=torch.float32))


def load_play_file(path):
    """optimize.py:42-65 ``load_data``: one ``play_*.json`` file (rows ``[planes 12x12x56, policy[1584], value,
    [game_len_for_side, step]]``, what ``write_play_file`` / the reference's workers write) -> list of
    ``[state (12,12,56) float64 ndarray, policy float32[1584], discounted value]``."""
    import json
    with open(path, "rt") as f:
        data = json.load(f)
    out = []
    for state, policy, value, game_lens in data:
        out.append([np.array(state), np.array(policy, dtype=np.float32), discounted_value(value, game_lens[0], game_lens[1])])
    return out


def load_play_files(directory=".", pattern="play_*.json"):
    """All play files of a directory in name order (optimize.py:26-40,68-100), concatenated."""
    import glob
    import os
    rows = []
    for path in sorted(glob.glob(os.path.join(directory, pattern))):
        rows.extend(load_play_file(path))
    return rows


def rows_to_tensors(rows):
    """Rows of ``load_play_file`` -> (states (B,56,12,12) f32, policies (B,1584) f32, values (B,) f32) for ``Trainer``:
    the HWC planes are moved to the CHW layout the net takes (api_hive.py:61-62, alpha_net.py:140-147)."""
    xs = np.stack([np.asarray(r[0], dtype=np.float32).transpose(2, 0, 1) for r in rows])
    ps = np.stack([np.asarray(r[1], dtype=np.float32) for r in rows])
    vs = np.asarray([r[2] for r in rows], dtype=np.float32)
    return torch.from_numpy(xs), torch.from_numpy(ps), torch.from_numpy(vs)


class Trainer:
    """Adam(lr=1e-3) + MultiStepLR([100,200,300,400], 0.2) as alpha_net.py:121-123."""

    def __init__(self, net, lr=1e-3, device=None):
        self.net = net
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        self.opt = torch.optim.Adam(net.parameters(), lr=lr)
        self.sched = torch.optim.lr_scheduler.MultiStepLR(self.opt, milestones=[100, 200, 300, 400], gamma=0.2)

    def step(self, states, policies, values):
        """One optimisation step on a mini-batch; returns the loss."""
        self.net.train()
        states, policies, values = states.to(self.device).float(), policies.to(self.device).float(), values.to(self.device).float()
        self.opt.zero_grad()
        policy_pred, value_pred = self.net(states)
        loss = alpha_loss(value_pred[:, 0], values, policy_pred, policies)
        loss.backward()
        self.opt.step()
        return float(loss.item())

    def epoch(self, states, policies, values, batch_size=512, shuffle=True, generator=None):
        n = states.shape[0]
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        losses = [self.step(states[order[i:i + batch_size]], policies[order[i:i + batch_size]], values[order[i:i + batch_size]])
                  for i in range(0, n, batch_size)]
        self.sched.step()
        return float(np.mean(losses))

    def save(self, path):
        """Checkpoint in the reference's format {'state_dict': ...} (self_play_with_train.py:86-97)."""
        torch.save({"state_dict": self.net.state_dict()}, path)
