"""BASELINE.json configs[4] under ncu: 10 M synthetic events into 256x256 count / mask / voxel grids and the warp-splat
(forward round + bilinear, backward), plus one fused window loss at the training shape.  Prints nothing timed.

    python profiles/run_encode_iwe.py [--events 10000000] [--reps 2]
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--events", type=int, default=10_000_000)
ap.add_argument("--res", type=int, default=256)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
snnflow = importlib.import_module("snn_event-based_optical_flow_b200")
enc, iwe = snnflow.encodings, snnflow.iwe
N, H, W = a.events, a.res, a.res
g = torch.Generator().manual_seed(3)
xs = torch.randint(0, W, (N,), generator=g).float().cuda()
ys = torch.randint(0, H, (N,), generator=g).float().cuda()
ts = torch.sort(torch.rand(N, generator=g)).values.cuda()
ps = (torch.randint(0, 2, (N,), generator=g).float() * 2 - 1).cuda()
flow = torch.tanh(0.5 * torch.randn(1, 2, H, W, generator=g)).cuda().requires_grad_(True)
events = torch.stack([ts, ys, xs, ps], dim=1).unsqueeze(0).contiguous()
pos, neg = (ps > 0).float().reshape(1, N, 1), (ps < 0).float().reshape(1, N, 1)
for rep in range(a.reps):
    enc.events_to_channels(xs, ys, ps, (H, W))
    enc.events_to_voxel(xs, ys, ts, ps, 5, (H, W))
    enc.events_to_image(xs, ys, ps.abs(), (H, W), accumulate=False)
    with torch.no_grad():
        iwe.compute_pol_iwe(flow, events, (H, W), pos, neg, 128, True)
    img = iwe.compute_pol_iwe(flow, events, (H, W), pos, neg, 128, False)
    img.square().sum().backward()
# the fused window loss at the training shape (BASELINE configs[1]: 8 x 10 x 1000 events, 128x128)
T, B, n, R = 10, 8, 1000, 128
ev = torch.stack([torch.sort(torch.rand(T, B, n, generator=g), dim=2).values, torch.randint(0, R, (T, B, n), generator=g).float(),
                  torch.randint(0, R, (T, B, n), generator=g).float(), torch.randint(0, 2, (T, B, n), generator=g).float() * 2 - 1], dim=3).cuda()
pm = torch.stack([(ev[..., 3] > 0).float(), (ev[..., 3] < 0).float()], dim=3).contiguous()
fl = (0.05 * torch.randn(T, B, 2, R, R, generator=g)).cuda().requires_grad_(True)
lossf = snnflow.EventWarping({"loader": {"resolution": [R, R]}, "loss": {"flow_regul_weight": 0.001}, "model": {"mask_output": False}},
                             torch.device("cuda"))
for rep in range(a.reps):
    lossf.window_loss(fl, ev, pm, None).backward()
torch.cuda.synchronize()
print("ok")
