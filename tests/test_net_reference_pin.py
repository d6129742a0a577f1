"""Pins the repository's fp32 network (hive_b200.HiveNet -- the model every bf16 / tensor-core tolerance test compares
against) to the REFERENCE's ChessNet (alpha_zero/alpha_net.py:82-95):
  * tests/golden/net_pins.npz holds outputs, parameter checksums and the state_dict layout of the unmodified ChessNet
    (oracle/gen_golden_net.py); HiveNet rebuilt under the same seeds must reproduce them (runs anywhere);
  * in the build container ChessNet itself is imported: its state_dict loads into HiveNet unchanged and both give
    identical outputs (torch.equal)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PINS = np.load(os.path.join(ROOT, "tests", "golden", "net_pins.npz"))


def _hivenet_like_fixture():
    import hive_b200
    from oracle.gen_golden_net import fixed_inputs, randomize_bn
    torch.manual_seed(0)
    net = hive_b200.HiveNet().eval()
    with torch.no_grad():
        randomize_bn(net, 7)
    return net, fixed_inputs()


def test_hivenet_reproduces_chessnet_fixture():
    net, x = _hivenet_like_fixture()
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in PINS["keys"]]                       # same names, same order
    for k, shape, s, sq in zip(PINS["keys"], PINS["shapes"], PINS["sums"], PINS["sqsums"]):
        t = sd[str(k)]
        assert ",".join(str(d) for d in t.shape) == str(shape), k
        assert float(t.double().sum()) == s and float((t.double() ** 2).sum()) == sq, k    # identical initialisation stream
    with torch.no_grad():
        p, v = net(x)
    assert np.allclose(p.numpy(), PINS["p"], rtol=1e-5, atol=1e-8) and np.allclose(v.numpy(), PINS["v"], rtol=1e-5, atol=1e-7)
    assert p.shape == (4, 1584) and v.shape == (4, 1)


def test_chessnet_state_dict_loads_and_outputs_are_identical():
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    import tempfile
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)                                     # importing alpha_net creates ./datasets/iter3/ in the CWD
        try:
            rh.load()
            from alpha_zero.alpha_net import ChessNet
        finally:
            os.chdir(cwd)
    import hive_b200
    from oracle.gen_golden_net import fixed_inputs, randomize_bn
    torch.manual_seed(5)
    ref = ChessNet().eval()
    with torch.no_grad():
        randomize_bn(ref, 11)
    mine = hive_b200.HiveNet().eval()
    missing = mine.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = fixed_inputs()
    with torch.no_grad():
        p0, v0 = ref(x)
        p1, v1 = mine(x)
    assert torch.equal(p0, p1) and torch.equal(v0, v1)
