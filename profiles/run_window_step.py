"""One training window of the bench workload (LIFFireNet C=32, 128x128, batch 8, 10 bins) through the layer-major
engine: forward + backward with a synthetic flow gradient.  Used under ncu (profiles/README.md); prints nothing timed.

    python profiles/run_window_step.py [--batch 8] [--res 128] [--bins 10] [--reps 2]
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--bins", type=int, default=10)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--kind", default="LIFFireNet")
ap.add_argument("--eval", action="store_true", help="forward only, no_grad (eval path)")
ap.add_argument("--time", type=int, default=0, help="also time this many further windows with CUDA events (not under ncu)")
a = ap.parse_args()
snnflow = importlib.import_module("snn_event-based_optical_flow_b200")
torch.manual_seed(0)
net = getattr(snnflow, a.kind)(dict(num_bins=2, encoding="cnt", base_num_channels=32, kernel_size=3,
                                    neuron_kwargs=dict(leak=(0.0, 1.0), thresh=(0.3, 0.1)))).cuda()
with torch.no_grad():
    net.pred.conv2d.weight.mul_(20)
g = torch.Generator().manual_seed(1)
cnt = torch.poisson(torch.full((a.bins, a.batch, 2, a.res, a.res), 0.06), generator=g).cuda()
gout = torch.randn(a.bins, a.batch, 2, a.res, a.res, generator=g).cuda()
for rep in range(a.reps):
    if a.eval:
        with torch.no_grad():
            flow = net.forward_window(cnt)
        continue
    net.zero_grad(set_to_none=True)
    flow = net.forward_window(cnt)
    (flow * gout).sum().backward()
    net.detach_states()
torch.cuda.synchronize()
if a.time:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(a.time):
        if a.eval:
            with torch.no_grad():
                flow = net.forward_window(cnt)
        else:
            net.zero_grad(set_to_none=True)
            flow = net.forward_window(cnt)
            (flow * gout).sum().backward()
            net.detach_states()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.time
    print(f"{ms:.3f} ms per window, {a.bins * a.batch / ms * 1e3:.0f} frames/s")
print("ok", float(flow.abs().mean()), [round(float(s[1].mean()), 4) for s in net._states])
