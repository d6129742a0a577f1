"""Network: BN folding is exact in fp32 (CPU); the bf16 device path agrees with the fp32 reference
architecture within the tolerance BASELINE.json states (1e-2) (GPU)."""
import numpy as np
import pytest
import torch


def _randomize_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def test_state_dict_names_match_reference_layout():
    import hive_b200
    net = hive_b200.HiveNet()
    keys = set(net.state_dict().keys())
    for k in ("conv.conv1.weight", "conv.conv1.bias", "conv.bn1.running_mean", "res_0.conv1.weight", "res_18.bn2.weight",
              "outblock.conv.weight", "outblock.fc1.weight", "outblock.fc2.bias", "outblock.conv1.weight", "outblock.fc.weight"):
        assert k in keys
    assert net.outblock.fc.weight.shape == (1584, 18432)
    assert sum(p.numel() for p in net.parameters()) == 51803188


def test_bn_folding_is_exact_in_fp32_cpu():
    import hive_b200
    torch.manual_seed(0)
    net = hive_b200.HiveNet().eval()
    _randomize_bn(net, 1)
    folded = hive_b200.FoldedNet(net, device="cpu", dtype=torch.float32)
    x = (torch.rand(2, 56, 12, 12) < 0.1).float()
    with torch.no_grad():
        p0, v0 = net(x)
    p1, v1 = folded(x)
    assert torch.allclose(p0, p1, atol=2e-6, rtol=1e-4) and torch.allclose(v0, v1, atol=1e-5)
    assert abs(float(p1.sum(1)[0]) - 1.0) < 1e-5


@pytest.mark.gpu
def test_bf16_device_path_within_1e2_of_fp32_reference():
    import hive_b200
    torch.manual_seed(0)
    net = hive_b200.HiveNet().eval()
    _randomize_bn(net, 2)
    b = hive_b200.HiveBatch(64)
    for _ in range(20):
        b.step_random(5, 55, True)
    planes = torch.from_numpy(b.planes().copy())                          # real positions as input
    with torch.no_grad():
        p_ref, v_ref = net.cuda()(planes.cuda())
    folded = hive_b200.FoldedNet(net, device="cuda")
    p, v = folded(planes.cuda())
    assert float((p - p_ref).abs().max()) <= 1e-2
    assert float((v - v_ref).abs().max()) <= 1e-2
    assert torch.allclose(p.sum(1), torch.ones(64, device="cuda"), atol=1e-3)


@pytest.mark.gpu
def test_selfplay_batch_smoke():
    import hive_b200
    torch.manual_seed(0)
    folded = hive_b200.FoldedNet(hive_b200.HiveNet().eval(), device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        sp = hive_b200.SelfPlayBatch(32, 8, hive_b200.LeafEvaluator(folded), stream=stream.cuda_stream, seed=3, collect=True)
        r = sp.play_moves(6)
    assert r["moves"] == 32 * 6 and r["waves"] >= 6 * 7
    turn, _, _ = sp.env.status()
    assert (turn == 7).all()
    assert all(len(s) == 6 for s in sp.samples)


@pytest.mark.gpu
def test_tensor_core_trunk_matches_references():
    """tcgen05 trunk kernel: within 1e-2 of the fp32 reference architecture on p and v, and close to the
    library bf16 path layer for layer (same folded weights)."""
    import hive_b200
    torch.manual_seed(0)
    net = hive_b200.HiveNet().eval()
    _randomize_bn(net, 3)
    b = hive_b200.HiveBatch(67)                                            # odd count: exercises the tail pair
    for _ in range(25):
        b.step_random(9, 55, True)
    planes = torch.from_numpy(b.planes().copy()).cuda()
    with torch.no_grad():
        p_ref, v_ref = net.cuda()(planes)
    folded = hive_b200.FoldedNet(net, device="cuda")
    p_lib, v_lib = folded(planes, trunk="torch")
    x_lib = folded._trunk_torch(planes).float()
    folded.attach_trunk(stream_ptr=torch.cuda.current_stream().cuda_stream, max_boards=128)
    x_tc = folded._trunk_tc(planes.to(torch.bfloat16).contiguous()).float()
    torch.cuda.synchronize()
    scale = float(x_lib.abs().max())
    assert float((x_tc - x_lib).abs().max()) <= 0.05 * scale + 0.05          # 39 layers of bf16 rounding
    p, v = folded(planes)
    assert float((p - p_ref).abs().max()) <= 1e-2 and float((v - v_ref).abs().max()) <= 1e-2
    assert float((p - p_lib).abs().max()) <= 1e-2
    assert folded.trunk.launches >= 40


@pytest.mark.gpu
def test_selfplay_samples_and_evaluator_match(tmp_path):
    """Row f1/f3: finished games produce the reference's sample rows (value -1 for both sides on a cut
    game); the evaluator match plays two nets with split colours and tallies wins."""
    import json
    import hive_b200
    torch.manual_seed(0)
    net_a = hive_b200.HiveNet().eval()
    torch.manual_seed(1)
    net_b = hive_b200.HiveNet().eval()
    stream = torch.cuda.Stream()
    fa = hive_b200.FoldedNet(net_a, device="cuda").attach_trunk(stream_ptr=stream.cuda_stream, max_boards=64)
    fb = hive_b200.FoldedNet(net_b, device="cuda").attach_trunk(stream_ptr=stream.cuda_stream, max_boards=64)
    with torch.cuda.stream(stream):
        sp = hive_b200.SelfPlayBatch(8, 4, hive_b200.LeafEvaluator(fa), stream=stream.cuda_stream, seed=5, collect=True)
        sp.play_moves(56)                                    # every game reaches the turn-55 cut (or ends) and is flushed
        assert len(sp.finished) >= 8 and len(sp.finished_samples) >= 8 * 50
        planes, pi, value, lens = sp.finished_samples[0]
        assert value in (-1, 1) and lens[1] == 1 and abs(float(pi.sum()) - 1.0) < 1e-5
        path = hive_b200.write_play_file(sp.finished_samples[:3], str(tmp_path))
        rows = json.load(open(path))
        assert len(rows) == 3 and len(rows[0][0]) == 12 and len(rows[0][0][0][0]) == 56 and len(rows[0][1]) == 1584
        ev = hive_b200.EvaluatorMatch(16, 4, hive_b200.LeafEvaluator(fa), hive_b200.LeafEvaluator(fb), stream=stream.cuda_stream, seed=2)
        r = ev.play()
    assert r["games"] == 16 and r["new_wins"] + r["best_wins"] + r["draws"] == 16 and r["moves"] > 16 * 20
    assert 0 < r["distinct_final_positions"] <= 1 and 20 < r["mean_game_len"] <= 55 and 0 <= r["white_win_rate"] <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("n_boards", [67, 300])
def test_heads_kernels_match_references(n_boards):
    """The whole network as hand-written kernels (net_forward: tcgen05 trunk + tcgen05 heads + finishing kernel) against
    the fp32 reference architecture (<= 1e-2, BASELINE.json) and against the library heads on the SAME trunk output
    (isolates the three head kernels: 1x1 convs, policy fc, softmax / value MLP); partial row tiles (67, 300 boards)."""
    import hive_b200
    torch.manual_seed(0)
    net = hive_b200.HiveNet().eval()
    _randomize_bn(net, 5)
    with torch.no_grad():                                                  # lively heads: non-trivial biases everywhere
        for p in net.outblock.parameters():
            p.add_(0.004 * torch.randn_like(p))
    b = hive_b200.HiveBatch(n_boards)
    for _ in range(30):
        b.step_random(11, 55, True)
    planes = torch.from_numpy(b.planes().copy()).cuda()
    with torch.no_grad():
        p_ref, v_ref = net.cuda()(planes)
    folded = hive_b200.FoldedNet(net, device="cuda").attach_trunk(stream_ptr=torch.cuda.current_stream().cuda_stream, max_boards=384)
    p_lib, v_lib = folded(planes, trunk="tc")                              # tensor-core trunk + library heads
    launches0 = folded.trunk.launches
    p, v = folded(planes)                                                  # trunk + heads kernels, nothing else
    torch.cuda.synchronize()
    assert folded.trunk.launches - launches0 == 40 + 3
    # (the library path rounds the 1,584 logits to bf16 before the softmax; here they stay float32, so the fp32 reference is the closer one)
    assert float((p - p_lib).abs().max()) <= 1e-2 and float((v - v_lib).abs().max()) <= 1e-2
    assert float((p - p_ref).abs().max()) <= float((p_lib - p_ref).abs().max()) + 1e-3
    assert float((p - p_ref).abs().max()) <= 1e-2 and float((v - v_ref).abs().max()) <= 1e-2
    assert torch.allclose(p.sum(1), torch.ones(n_boards, device="cuda"), atol=1e-4)
    # written straight into caller-owned arenas (what the search's wave does)
    pol = torch.full((n_boards, 1584), -1.0, device="cuda")
    val = torch.full((n_boards,), -7.0, dtype=torch.float64, device="cuda")
    hive_b200.LeafEvaluator(folded)(planes.to(torch.bfloat16).contiguous().data_ptr(), pol.data_ptr(), val.data_ptr(), 0, n_boards)
    torch.cuda.synchronize()
    assert torch.equal(pol, p) and torch.equal(val.float().reshape(-1, 1), v)
