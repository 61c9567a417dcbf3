import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from snnflow_b200 import _lib
from test_gpu_tc import pack, run_tc
B, C, H, W = 8, 32, 128, 128
w = (torch.rand(C, C, 3, 3) - 0.5).cuda()
blob = pack(w, None, C, C)
x = (torch.rand(3, B, C, H, W) < 0.2).float().cuda()
lam, theta = torch.full((C,), 0.5).cuda(), torch.full((C,), 0.3).cuda()
run_tc(x, blob, False, lam, theta)
run_tc(x, blob, False, lam, theta)
print("ok")
