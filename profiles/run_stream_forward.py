"""Streaming per-bin forward (eval_flow.py:220 pattern): wall clock vs device time, and the per-kernel table of the library's
live profiler.  python profiles/run_stream_forward.py [--kind LIFFireFlowNet|LIFFireNet] [--res 256] [--batch 16]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snnflow_b200 as snnflow  # noqa: E402
from snnflow_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="LIFFireFlowNet")
ap.add_argument("--res", type=int, default=256)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--channels", type=int, default=32)
ap.add_argument("--n", type=int, default=200)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = getattr(snnflow, a.kind)(dict(num_bins=2, encoding="cnt", base_num_channels=a.channels, kernel_size=3,
                                    neuron_kwargs=dict(leak=(0.0, 1.0), thresh=(0.3, 0.1)))).to(dev)
g = torch.Generator().manual_seed(7)
x = torch.poisson(torch.full((10, a.batch, 2, a.res, a.res), 0.06), generator=g).to(dev)
with torch.no_grad():
    for t in range(10):
        net(None, x[t])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(a.n):
        net(None, x[i % 10])
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    print(f"{a.kind} {a.res}x{a.res} batch {a.batch}: device {e0.elapsed_time(e1) / a.n * 1e3:.1f} us/forward, host issue "
          f"{t_issue / a.n * 1e6:.1f} us/forward, wall {t_wall / a.n * 1e6:.1f} us/forward")
    _lib.profile(True)
    for i in range(10):
        net(None, x[i % 10])
    torch.cuda.synchronize()
    prof = _lib.profile_summary()
    _lib.profile(False)
    tot = sum(p["ms"] for p in prof.values())
    for k, p in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"  {k:20s} x{p['launches']:3d}  {p['ms'] / p['launches'] * 1e3:8.1f} us/launch  {p['ms'] / tot * 100:5.1f} %  {p.get('bytes', 0) / max(p['ms'], 1e-9) / 1e6:8.0f} GB/s (algorithmic)")
    print(f"  total {tot / 10 * 1e3:.1f} us/forward of kernel time")
