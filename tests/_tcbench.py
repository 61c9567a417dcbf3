import sys, torch, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from snnflow_b200 import _lib
from test_gpu_tc import pack, run_tc
L = _lib.lib()
for (B, C, H, W, rec) in [(8, 32, 128, 128, False), (8, 32, 128, 128, True), (16, 32, 256, 256, False)]:
    w = (torch.rand(C, C, 3, 3) - 0.5).cuda(); wr = (torch.rand(C, C, 3, 3) - 0.5).cuda() if rec else None
    blob = pack(w, wr, C, C)
    x = (torch.rand(4, B, C, H, W) < 0.2).float().cuda()
    lam, theta = torch.full((C,), 0.5).cuda(), torch.full((C,), 0.3).cuda()
    run_tc(x, blob, rec, lam, theta)
    _lib.profile(True)
    t0 = time.time(); run_tc(x, blob, rec, lam, theta); t1 = time.time()
    p = _lib.profile_summary(); _lib.profile(False)
    k = p["convlif_fwd_tc"]
    print((B, C, H, W, rec), "wall", round(t1 - t0, 4), "us/launch", round(1e3 * k["ms"] / k["launches"], 1), "GB/s", round(k["bytes"] / k["ms"] / 1e6, 1), "TF", round(k["flops"] / k["ms"] / 1e9, 1))
