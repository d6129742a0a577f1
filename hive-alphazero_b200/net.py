"""Policy/value network of the hot path (reference: alpha_zero/alpha_net.py:25-95, class ChessNet).

* ``HiveNet``     -- fp32 torch module with the reference's architecture AND parameter names
                     (``conv.conv1``, ``res_%i.conv1`` ... ``outblock.fc``), so a reference
                     checkpoint ``{'state_dict': ...}`` (self_play.py:92-96) loads unchanged. It is
                     the fp32 reference the bf16 path is checked against (tolerance 1e-2).
* ``FoldedNet``   -- inference form: BatchNorm folded into the convolutions, bf16 weights,
                     channels-last activations.  forward(planes) -> (p softmax fp32, v tanh fp32).
* ``LeafEvaluator`` -- glues a FoldedNet to the device buffers of MctsBatch (zero-copy views of the
                     library's arenas through the CUDA array interface).

The 3x3 trunk convolutions are 98 % of the 6.56 GFLOP/sample: ``TensorCoreTrunk`` runs them in the
hand-written tcgen05 kernel of csrc/hive_conv_kernel.cuh (``net_*`` entry points of the C ABI);
the two small heads are plain library GEMMs / 1x1 convolutions through torch.  ``FoldedNet`` uses
the trunk kernel once it is attached (``attach_trunk``); ``forward(..., trunk="torch")`` runs the
same folded weights through the library instead (the comparison path of the parity tests and the
only path on a machine without a GPU).
"""
import ctypes
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import config as C

N_RES = 19
CH = 256
BOARD = C.MAX_MAP_FULL                      # 12
CELLS = BOARD * BOARD                       # 144
POLICY_CH = 128


class _Stem(nn.Module):                     # alpha_net.py:25-34
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(C.STATE_FEATURES, CH, 3, stride=1, padding=1)
        self.bn1 = nn.BatchNorm2d(CH)

    def forward(self, s):
        return F.relu(self.bn1(self.conv1(s)))


class _Residual(nn.Module):                 # alpha_net.py:36-54
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(CH, CH, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(CH)
        self.conv2 = nn.Conv2d(CH, CH, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(CH)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + x)


class _Heads(nn.Module):                    # alpha_net.py:56-80
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(CH, 1, kernel_size=1)           # value head
        self.bn = nn.BatchNorm2d(1)
        self.fc1 = nn.Linear(CELLS, 64)
        self.fc2 = nn.Linear(64, 1)
        self.conv1 = nn.Conv2d(CH, POLICY_CH, kernel_size=1)  # policy head
        self.bn1 = nn.BatchNorm2d(POLICY_CH)
        self.fc = nn.Linear(CELLS * POLICY_CH, C.ACTION_SPACE)

    def forward(self, s):
        v = F.relu(self.bn(self.conv(s))).view(-1, CELLS)
        v = torch.tanh(self.fc2(F.relu(self.fc1(v))))
        p = F.relu(self.bn1(self.conv1(s))).view(-1, CELLS * POLICY_CH)      # NCHW flatten: c*144 + h*12 + w
        p = F.log_softmax(self.fc(p), dim=1).exp()
        return p, v


class HiveNet(nn.Module):
    """fp32 reference network; state_dict-compatible with the reference's ChessNet."""

    def __init__(self):
        super().__init__()
        self.conv = _Stem()
        for i in range(N_RES):
            setattr(self, "res_%i" % i, _Residual())
        self.outblock = _Heads()

    def forward(self, s):
        s = self.conv(s)
        for i in range(N_RES):
            s = getattr(self, "res_%i" % i)(s)
        return self.outblock(s)


def _fold(conv_w, conv_b, bn):
    """conv followed by eval-mode BatchNorm -> one conv (w', b')."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = conv_w * scale.view(-1, 1, 1, 1)
    b = bn.bias - bn.running_mean * scale
    if conv_b is not None:
        b = b + conv_b * scale
    return w, b


class TensorCoreTrunk:
    """The 39 folded 3x3 convolutions on the tensor cores (net_create / net_load_conv_host /
    net_trunk_forward)."""

    def __init__(self, folded_fp32, device_index, stream_ptr, max_boards, heads_fp32=None):
        from ._capi import check, lib
        self._lib, self._check = lib(), check
        h = ctypes.c_void_p()
        check(self._lib.net_create(int(device_index), stream_ptr, int(max_boards), ctypes.byref(h)), "net_create")
        self._h, self.max_boards = h, int(max_boards)
        self.has_heads = False
        self.load(folded_fp32)
        if heads_fp32 is not None:
            self.load_heads(heads_fp32)

    def load_heads(self, heads_fp32):
        """The ten head tensors (fp32) in net_load_heads_host's order; reloads keep the device addresses.  CUDA tensors
        are packed by kernels on the device (net_load_heads_dev: the fc weight then comes in the REFERENCE's layout)."""
        if heads_fp32[0].is_cuda:
            ts = [t.detach().float().contiguous() for t in heads_fp32]
            torch.cuda.synchronize()                                # the sources are ready, whatever stream made them
            self._check(self._lib.net_load_heads_dev(self._h, *[t.data_ptr() for t in ts]), "net_load_heads_dev")
            torch.cuda.synchronize()                                # ... and stay alive until the pack kernels have read them
        else:
            arrs = [np.ascontiguousarray(t.detach().float().cpu().numpy()) for t in heads_fp32]
            self._check(self._lib.net_load_heads_host(self._h, *[a.ctypes.data for a in arrs]), "net_load_heads_host")
        self.has_heads = True

    def forward_into(self, planes_ptr, n_boards, policy_ptr, value_ptr):
        """Whole network: planes (device bf16 [n][56][144]) -> policy (device f32 [n][1584]), value (device f64 [n])."""
        self._check(self._lib.net_forward(self._h, planes_ptr, int(n_boards), policy_ptr, value_ptr), "net_forward")

    def load(self, folded_fp32):
        """(Re)load the 39 folded convolutions; the packed operands keep their device addresses."""
        if folded_fp32[0][0].is_cuda:                               # packed on the device (net_load_conv_dev)
            ts = [(w.detach().float().contiguous(), b.detach().float().contiguous()) for w, b in folded_fp32]
            torch.cuda.synchronize()
            for layer, (w, b) in enumerate(ts):
                self._check(self._lib.net_load_conv_dev(self._h, layer, w.data_ptr(), b.data_ptr(), w.shape[1]), "net_load_conv_dev")
            torch.cuda.synchronize()
            return
        for layer, (w, b) in enumerate(folded_fp32):
            w = np.ascontiguousarray(w.detach().float().cpu().numpy())
            b = np.ascontiguousarray(b.detach().float().cpu().numpy())
            self._check(self._lib.net_load_conv_host(self._h, layer, w.ctypes.data, b.ctypes.data, w.shape[1]), "net_load_conv_host")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward_ptr(self, planes_ptr, n_boards):
        """planes_ptr: device bf16 [n][56][144].  Returns the device pointer of [n][144][256] bf16 NHWC."""
        out = ctypes.c_void_p()
        self._check(self._lib.net_trunk_forward(self._h, planes_ptr, int(n_boards), ctypes.byref(out)), "net_trunk_forward")
        return out.value

    @property
    def launches(self):
        return self._lib.net_launch_count(self._h)


class FoldedNet:
    """BN-folded bf16 inference network built from a HiveNet (weights frozen at construction)."""

    def __init__(self, net, device="cuda", dtype=torch.bfloat16):
        net = net.eval()
        self.device, self.dtype = torch.device(device), dtype
        self.trunk = None
        with torch.no_grad():
            folded = [_fold(net.conv.conv1.weight, net.conv.conv1.bias, net.conv.bn1)]
            for i in range(N_RES):
                r = getattr(net, "res_%i" % i)
                folded += [_fold(r.conv1.weight, None, r.bn1), _fold(r.conv2.weight, None, r.bn2)]
            keep = (lambda t: t.detach().float().to(self.device).contiguous()) if self.device.type == "cuda" else (lambda t: t.detach().float().cpu())
            self._folded_fp32 = [(keep(w), keep(b)) for w, b in folded]
        with torch.no_grad():
            def put(w, b):
                w = w.to(self.device, dtype).contiguous(memory_format=torch.channels_last) if w.dim() == 4 else w.to(self.device, dtype)
                return w, b.to(self.device, dtype)
            self.stem = put(*_fold(net.conv.conv1.weight, net.conv.conv1.bias, net.conv.bn1))
            self.blocks = []
            for i in range(N_RES):
                r = getattr(net, "res_%i" % i)
                self.blocks.append((put(*_fold(r.conv1.weight, None, r.bn1)), put(*_fold(r.conv2.weight, None, r.bn2))))
            ob = net.outblock
            self.vconv = put(*_fold(ob.conv.weight, ob.conv.bias, ob.bn))
            self.pconv = put(*_fold(ob.conv1.weight, ob.conv1.bias, ob.bn1))
            # The heads on the NHWC activations the trunk produces: a 1x1 convolution is a plain GEMM over (board*cell,
            # channel) rows, and the policy fc takes the cell-major flatten if its columns are permuted once from the
            # reference's channel-major order (c*144 + cell, alpha_net.py:81) to cell*128 + c -- no layout copies per wave.
            self.vlin = (self.vconv[0].reshape(1, CH).contiguous(), self.vconv[1])
            self.plin = (self.pconv[0].reshape(POLICY_CH, CH).contiguous(), self.pconv[1])
            self.fc_nhwc = ob.fc.weight.detach().view(C.ACTION_SPACE, POLICY_CH, CELLS).permute(0, 2, 1).reshape(
                C.ACTION_SPACE, CELLS * POLICY_CH).to(self.device, dtype).contiguous()
            self.fc1 = (ob.fc1.weight.to(self.device, torch.float32), ob.fc1.bias.to(self.device, torch.float32))
            self.fc2 = (ob.fc2.weight.to(self.device, torch.float32), ob.fc2.bias.to(self.device, torch.float32))
            self.fc = (ob.fc.weight.to(self.device, dtype), ob.fc.bias.to(self.device, torch.float32))
            # the same heads for the hand-written kernels (net_load_heads_host): fp32 on the host, fc columns cell-major
            pw, pb = _fold(ob.conv1.weight, ob.conv1.bias, ob.bn1)
            vw, vb = _fold(ob.conv.weight, ob.conv.bias, ob.bn)
            # (on a CUDA device they stay there and the fc weight keeps the reference's layout: net_load_heads_dev permutes)
            on_dev = self.device.type == "cuda"
            fc_w = ob.fc.weight.detach() if on_dev else ob.fc.weight.detach().view(C.ACTION_SPACE, POLICY_CH, CELLS).permute(0, 2, 1).reshape(
                C.ACTION_SPACE, CELLS * POLICY_CH)
            self._heads_fp32 = [t.detach().float().to(self.device if on_dev else "cpu").contiguous() for t in (
                pw.reshape(POLICY_CH, CH), pb, vw.reshape(CH), vb.reshape(1), fc_w,
                ob.fc.bias, ob.fc1.weight, ob.fc1.bias, ob.fc2.weight.reshape(-1), ob.fc2.bias.reshape(1))]

    @torch.no_grad()
    def reload(self, net):
        """Take over the weights of `net` IN PLACE (after a weight broadcast): every device tensor and the tensor-core
        trunk's packed operands keep their addresses, so CUDA graphs captured over this network stay valid."""
        fresh = FoldedNet(net, device=self.device, dtype=self.dtype)

        def take(dst, src):
            for d, s_ in zip(dst, src):
                d.copy_(s_)
        take(self.stem, fresh.stem)
        for (a1, a2), (b1, b2) in zip(self.blocks, fresh.blocks):
            take(a1, b1); take(a2, b2)
        for name in ("vconv", "pconv", "vlin", "plin", "fc1", "fc2", "fc"):
            take(getattr(self, name), getattr(fresh, name))
        self.fc_nhwc.copy_(fresh.fc_nhwc)
        self._folded_fp32, self._heads_fp32 = fresh._folded_fp32, fresh._heads_fp32
        if self.trunk is not None:
            self.trunk.load(self._folded_fp32)
            self.trunk.load_heads(self._heads_fp32)
        return self

    def trunk_weights(self):
        """[(w (256,Cin,3,3), b (256,)) ...] of the 39 folded 3x3 convolutions, in execution order."""
        out = [self.stem]
        for c1, c2 in self.blocks:
            out += [c1, c2]
        return out

    def attach_trunk(self, stream_ptr=None, max_boards=4096):
        """Create the tensor-core trunk on this net's device; kernels are queued on `stream_ptr`
        (a cudaStream_t as int; None = the legacy default stream)."""
        if self.device.type != "cuda":
            raise RuntimeError("the tensor-core trunk needs a CUDA device (there is no CPU fallback)")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.trunk = TensorCoreTrunk(self._folded_fp32, idx, stream_ptr, max_boards, heads_fp32=self._heads_fp32)
        return self

    def _trunk_torch(self, planes):
        x = planes.to(self.dtype).contiguous(memory_format=torch.channels_last)
        x = F.relu(F.conv2d(x, self.stem[0], self.stem[1], padding=1))
        for (w1, b1), (w2, b2) in self.blocks:
            y = F.relu(F.conv2d(x, w1, b1, padding=1))
            x = F.relu(F.conv2d(y, w2, b2, padding=1) + x)
        return x

    def _trunk_tc(self, planes):
        """planes: contiguous bf16 (B,56,12,12) on the device."""
        n = planes.shape[0]
        outs = []
        for s in range(0, n, self.trunk.max_boards):
            e = min(n, s + self.trunk.max_boards)
            ptr = self.trunk.forward_ptr(planes[s:e].data_ptr(), e - s)
            nhwc = device_view(ptr, (e - s, BOARD, BOARD, CH), "<u2", self.device).view(torch.bfloat16)
            outs.append(nhwc if n <= self.trunk.max_boards else nhwc.clone())
        x = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return x.permute(0, 3, 1, 2)                         # NCHW view over NHWC memory (= channels_last)

    @torch.no_grad()
    def forward(self, planes, trunk=None):
        """planes (B,56,12,12) bf16/fp32 on the device -> (p (B,1584) fp32, v (B,1) fp32).
        trunk: None = the hand-written network (trunk + heads) if attached, else torch; "tc" = tensor-core trunk + library
        heads; "torch" forces the library path."""
        if self.trunk is not None and trunk is None and self.trunk.has_heads:
            # the whole network as hand-written sm_100a kernels (net_forward): trunk + heads, no library kernel
            pl = planes.to(self.dtype).contiguous()
            n = pl.shape[0]
            p = torch.empty((n, C.ACTION_SPACE), dtype=torch.float32, device=self.device)
            v = torch.empty((n,), dtype=torch.float64, device=self.device)
            for s in range(0, n, self.trunk.max_boards):
                e = min(n, s + self.trunk.max_boards)
                self.trunk.forward_into(pl[s:e].data_ptr(), e - s, p[s:e].data_ptr(), v[s:e].data_ptr())
            return p, v.float().reshape(-1, 1)
        if self.trunk is not None and trunk != "torch":
            x = self._trunk_tc(planes.to(self.dtype).contiguous())
        else:
            x = self._trunk_torch(planes)
        rows = x.permute(0, 2, 3, 1).reshape(-1, CH)         # (B*144, 256): a view when x is NHWC in memory (both trunks)
        v = F.relu(F.linear(rows, *self.vlin)).float().reshape(-1, CELLS)
        v = torch.tanh(F.linear(F.relu(F.linear(v, *self.fc1)), *self.fc2))
        p = F.relu(F.linear(rows, *self.plin)).reshape(-1, CELLS * POLICY_CH)     # cell-major flatten
        logits = F.linear(p, self.fc_nhwc).float() + self.fc[1]
        return torch.softmax(logits, dim=1), v

    __call__ = forward


class _DevArray:
    """Zero-copy view of a library-owned device arena for torch (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_view(ptr, shape, typestr, device="cuda"):
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


class LeafEvaluator:
    """evaluate_device callback for MctsBatch.search_device: planes arena -> policy / value arenas."""

    def __init__(self, folded, max_batch=4096, stream=None):
        """stream: the torch stream the network's kernels are queued on when it is NOT the caller's current stream (the
        trunk was attached with that stream's pointer); SplitEvaluator then forks / joins around the call."""
        self.net, self.max_batch = folded, max_batch
        self.stream = stream
        self.calls = 0

    def __call__(self, planes_ptr, policy_ptr, value_ptr, mask_ptr, n):
        trunk = self.net.trunk
        if trunk is not None and trunk.has_heads:
            # planes arena -> policy / value arenas by the library's own kernels: no copy, no torch kernel in the wave
            step = min(self.max_batch, trunk.max_boards)
            for s in range(0, n, step):
                e = min(n, s + step)
                trunk.forward_into(planes_ptr + s * C.STATE_FEATURES * CELLS * 2, e - s, policy_ptr + s * C.ACTION_SPACE * 4, value_ptr + s * 8)
            self.calls += 1
            return
        dev = self.net.device
        planes = device_view(planes_ptr, (n, C.STATE_FEATURES, BOARD, BOARD), "<u2", dev).view(torch.bfloat16)
        policy = device_view(policy_ptr, (n, C.ACTION_SPACE), "<f4", dev)
        value = device_view(value_ptr, (n,), "<f8", dev)
        for s in range(0, n, self.max_batch):
            e = min(n, s + self.max_batch)
            p, v = self.net(planes[s:e])
            policy[s:e].copy_(p)
            value[s:e].copy_(v.reshape(-1).double())
        self.calls += 1


class SplitEvaluator:
    """Two networks in one search: rows [0, split) are evaluated by `first`, rows [split, n) by `second`
    (the evaluator match keeps the games where the new net plays white in the first half).
    An evaluator with a `stream` of its own runs beside the other one: the call forks that stream off the caller's and
    joins it again, so the persistent convolution CTAs of one network fill the SMs the other network's last partial wave
    leaves idle (512 boards are 3.5 waves of 148 CTAs; two such launches side by side lose one tail instead of two)."""

    def __init__(self, first, second, split):
        self.first, self.second, self.split = first, second, int(split)

    @staticmethod
    def _run(ev, args):
        side = getattr(ev, "stream", None)
        if side is None:
            ev(*args)
            return None
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            ev(*args)
            done = torch.cuda.Event()
            done.record(side)
        return done

    def __call__(self, planes_ptr, policy_ptr, value_ptr, mask_ptr, n):
        k = self.split
        joins = []
        if k > 0:
            joins.append(self._run(self.first, (planes_ptr, policy_ptr, value_ptr, mask_ptr, k)))
        if n > k:
            joins.append(self._run(self.second, (planes_ptr + k * C.STATE_FEATURES * CELLS * 2, policy_ptr + k * C.ACTION_SPACE * 4,
                                                 value_ptr + k * 8, mask_ptr + k if mask_ptr else mask_ptr, n - k)))
        for done in joins:
            if done is not None:
                torch.cuda.current_stream().wait_event(done)


def host_net_callable(folded):
    """planes (B,56,12,12) float32 ndarray -> (p, v) ndarrays; for HivePlayer.net."""
    def run(planes):
        x = torch.as_tensor(np.asarray(planes, dtype=np.float32), device=folded.device)
        p, v = folded(x)
        return p.float().cpu().numpy(), v.float().cpu().numpy()
    return run
