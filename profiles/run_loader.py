"""One loader window of 10 M synthetic events at 256x256 through snnflow_format_window (BASELINE.json configs[4] shape);
used under ncu (profiles/capture.sh).  Prints nothing timed."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

snnflow = importlib.import_module("snn_event-based_optical_flow_b200")
N, H, W = 10_000_000, 256, 256
g = torch.Generator().manual_seed(3)
fmt = snnflow.EventWindowFormatter(
    {"data": {"mode": "events"}, "loader": {"resolution": [H, W], "std_resolution": [H, W], "batch_size": 1,
                                            "augment": ["Horizontal", "Vertical", "Polarity"], "augment_prob": [1.0, 1.0, 1.0]},
     "hot_filter": {"enabled": True, "max_px": 100, "min_obvs": 5, "max_rate": 0.8}}, 5)
raw = [torch.randint(0, W, (1, N), generator=g).float().cuda(), torch.randint(0, H, (1, N), generator=g).float().cuda(),
       torch.sort(torch.rand(1, N, generator=g)).values.cuda(), torch.randint(0, 2, (1, N), generator=g).float().cuda()]
for _ in range(3):
    out = fmt.format_batch(*raw)
torch.cuda.synchronize()
print("ok", float(out["event_cnt"].sum()))
