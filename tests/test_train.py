"""Row f2: the loss is the reference's AlphaLoss (checked against the real class when the reference
tree is present), value targets follow the discount rule, one step on a tiny batch lowers the loss."""
import os
import sys

import numpy as np
import pytest
import torch

import hive_b200
from oracle import ref_harness as rh


def test_alpha_loss_formula_and_discount():
    torch.manual_seed(0)
    vp, v = torch.rand(5) * 2 - 1, torch.tensor([1.0, -1.0, -1.0, 1.0, -1.0])
    pp = torch.softmax(torch.randn(5, 1584), 1)
    p = torch.softmax(torch.randn(5, 1584) * 3, 1)
    ours = hive_b200.alpha_loss(vp, v, pp, p)
    manual = ((v - vp) ** 2 + (-(p * (1e-6 + pp).log()).sum(1))).mean()
    assert torch.allclose(ours, manual)
    assert hive_b200.discounted_value(-1, 20, 20) == -1
    assert abs(hive_b200.discounted_value(1, 20, 17) - 0.99 ** 3) < 1e-12


@pytest.mark.skipif(not rh.available(), reason="reference tree not present")
def test_alpha_loss_matches_reference_class(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                        # alpha_net.py creates ./datasets/iter3 on import
    rh.load()
    from alpha_zero.alpha_net import AlphaLoss        # the reference's own module (matplotlib stubbed)
    torch.manual_seed(1)
    vp, v = torch.rand(7) * 2 - 1, torch.sign(torch.randn(7))
    pp = torch.softmax(torch.randn(7, 1584), 1)
    p = torch.softmax(torch.randn(7, 1584) * 2, 1)
    ref = AlphaLoss()(vp, v, pp, p)
    assert torch.equal(ref, hive_b200.alpha_loss(vp, v, pp, p))


def test_training_step_lowers_loss_on_a_tiny_batch():
    torch.manual_seed(0)
    net = hive_b200.HiveNet()
    # shrink the work: keep the architecture but only run a few samples on CPU
    planes = np.zeros((4, 56 * 144), dtype=np.uint16)
    planes[:, 100:140] = 0x3F80
    samples = []
    for i in range(4):
        pi = np.zeros(1584, dtype=np.float32); pi[10 * (i + 1)] = 1.0
        samples.append((planes[i], pi, 1 if i % 2 else -1, (10, 7 + i % 3)))
    x, p, v = hive_b200.samples_to_tensors(samples)
    assert x.shape == (4, 56, 12, 12) and p.shape == (4, 1584) and abs(float(v[0]) + 0.99 ** 3) < 1e-6
    tr = hive_b200.Trainer(net, lr=1e-3)
    l0 = tr.step(x, p, v)
    l1 = tr.step(x, p, v)
    l2 = tr.step(x, p, v)
    assert l2 < l0
    xo, po, _ = hive_b200.samples_to_tensors(samples, one_hot_policy=True)
    assert (po.sum(1) == 1).all()


def test_play_file_round_trip_matches_sample_tensors(tmp_path):
    """write_play_file (self_play.py:100-112 format) -> load_play_file (optimize.py:42-65) gives the same training
    tensors as the in-memory path, including the value discount 0.99 ** (game_len - step)."""
    rng = np.random.RandomState(3)
    samples = []
    for i in range(5):
        planes = np.where(rng.rand(56 * 144) < 0.05, 0x3F80, 0).astype(np.uint16)
        planes.reshape(56, 144)[31] = (np.float32(7 + i).view(np.uint32) >> 16).astype(np.uint16)     # turn plane
        pi = rng.rand(1584).astype(np.float32); pi /= pi.sum()
        samples.append((planes, pi, 1 if i % 2 else -1, (12, 8 + i)))
    path = hive_b200.write_play_file(samples, str(tmp_path))
    rows = hive_b200.load_play_file(path)
    assert len(rows) == 5 and rows[0][0].shape == (12, 12, 56) and rows[0][1].dtype == np.float32
    assert abs(rows[0][2] - (-1) * 0.99 ** 4) < 1e-12 and rows[4][2] == -1          # step 12 of 12: raw value
    x1, p1, v1 = hive_b200.rows_to_tensors(hive_b200.load_play_files(str(tmp_path)))
    x0, p0, v0 = hive_b200.samples_to_tensors(samples)
    assert torch.equal(x0, x1) and torch.allclose(p0, p1, atol=1e-7) and torch.allclose(v0, v1, atol=1e-6)


@pytest.mark.gpu
def test_closed_loop_on_the_device_selfplay_train_reload(tmp_path):
    """Row f2 on the GPU: samples from a batched self-play -> Adam steps of the fp32 network on the device (loss falls) ->
    the folded bf16 network and the tensor-core operands take the new weights over IN PLACE (the captured wave graph
    stays valid) -> the next search sees the trained network; checkpoint in the reference's {'state_dict'} format."""
    torch.manual_seed(0)
    net = hive_b200.HiveNet().cuda()
    stream = torch.cuda.Stream()
    folded = hive_b200.FoldedNet(net, device="cuda").attach_trunk(stream_ptr=stream.cuda_stream, max_boards=32)
    with torch.cuda.stream(stream):
        sp = hive_b200.SelfPlayBatch(16, 6, hive_b200.LeafEvaluator(folded), stream=stream.cuda_stream, seed=4, collect=True,
                                     wave_graph=hive_b200.WaveGraph(stream))
        sp.play_moves(56)                                       # every game is cut at turn 55 (or ends) and flushed
        assert len(sp.finished_samples) >= 16 * 50
        x, p, v = hive_b200.samples_to_tensors(sp.finished_samples[:256])
        planes = torch.from_numpy(sp.env.planes().copy()).cuda()
        p_before, v_before = folded(planes)
        tr = hive_b200.Trainer(net, lr=1e-3)
        losses = [tr.step(x, p, v) for _ in range(6)]
        assert losses[-1] < losses[0]
        net.eval()
        folded.reload(net)
        p_after, v_after = folded(planes)
        with torch.no_grad():
            p_ref, v_ref = net(planes)
        torch.cuda.synchronize()
        assert float((p_after - p_before).abs().max()) > 1e-4               # the weights really changed ...
        assert float((p_after - p_ref).abs().max()) <= 1e-2 and float((v_after - v_ref).abs().max()) <= 2e-2   # ... to the trained ones
        r = sp.play_moves(2)                                    # the captured wave graph runs on with the reloaded network
        assert r["moves"] == 32
    path = str(tmp_path / "iter.pth.tar")
    tr.save(path)
    ck = torch.load(path)
    assert set(ck) == {"state_dict"} and "outblock.fc.weight" in ck["state_dict"]
