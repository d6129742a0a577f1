"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/net_pins.npz from the UNMODIFIED reference network
(alpha_zero/alpha_net.py:82-95, class ChessNet).  Build container only; run from a scratch directory
(importing alpha_net creates ./datasets/iter3/ in the CWD, alpha_net.py:164-165).

Stored: for ChessNet initialised under torch.manual_seed(0) (eval mode; BatchNorm statistics then randomised by a
generator seeded with 7), the outputs (p, v) on 4 fixed binary-plane inputs, a checksum of every parameter
tensor (sum and sum of squares as float64) and the list of state_dict keys with shapes.  The travelling test
rebuilds hive_b200.HiveNet under the same seeds and must reproduce all of it."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def randomize_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def fixed_inputs():
    g = torch.Generator().manual_seed(123)
    x = (torch.rand(4, 56, 12, 12, generator=g) < 0.08).float()
    x[:, 31] = torch.tensor([3.0, 17.0, 40.0, 54.0]).view(4, 1, 1)       # plane 31 = turn number
    return x


def main():
    rh.load()
    from alpha_zero.alpha_net import ChessNet
    torch.manual_seed(0)
    net = ChessNet().eval()
    with torch.no_grad():
        randomize_bn(net, 7)
        x = fixed_inputs()
        p, v = net(x)
    sd = net.state_dict()
    keys = list(sd.keys())
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "net_pins.npz"),
        p=p.numpy(), v=v.numpy(), keys=np.array(keys),
        shapes=np.array([",".join(str(d) for d in sd[k].shape) for k in keys]),
        sums=np.array([float(sd[k].double().sum()) for k in keys]),
        sqsums=np.array([float((sd[k].double() ** 2).sum()) for k in keys]))
    print("ChessNet pins written:", len(keys), "tensors; p max", float(p.max()), "v", v.view(-1).tolist())


if __name__ == "__main__":
    main()
