import sys, time; sys.path.insert(0, '.')
import torch, hive_b200
torch.manual_seed(0)
net = hive_b200.HiveNet().eval().cuda()
st = torch.cuda.Stream()
f = hive_b200.FoldedNet(net, device="cuda").attach_trunk(stream_ptr=st.cuda_stream, max_boards=256)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f.reload(net); torch.cuda.synchronize(); print("reload ms", (time.perf_counter() - t0) * 1e3)
