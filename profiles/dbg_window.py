"""Debug helper (not a test): python profiles/dbg_window.py <stage> ; stages: eval | fwd | bwd [kind C H W]"""
import os, sys
os.environ.setdefault("CUDA_LAUNCH_BLOCKING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_window import make_net, runner_of

stage = sys.argv[1]
kind = sys.argv[2] if len(sys.argv) > 2 else "LIFFireFlowNet"
C, H, W = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (32, 16, 128)
net = make_net(kind, C)
g = torch.Generator().manual_seed(5)
T, B = 4, 2
cnt = torch.poisson(torch.full((T, B, 2, H, W), 0.25), generator=g).cuda()
runner_of(net, "layer_major")
if stage == "eval":
    with torch.no_grad():
        f = net.forward_window(cnt)
    torch.cuda.synchronize()
    print("eval ok", float(f.abs().mean()), [float(s[1].mean()) for s in net._states])
else:
    f = net.forward_window(cnt)
    torch.cuda.synchronize()
    print("fwd ok", float(f.abs().mean()), [float(s[1].mean()) for s in net._states])
    if stage == "bwd":
        f.square().sum().backward()
        torch.cuda.synchronize()
        print("bwd ok", {n: float(p.grad.abs().sum()) for n, p in net.named_parameters()})
