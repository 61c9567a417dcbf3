"""Probe (not a test): e2e step time with / without the side-stream prefetch, plus raw H2D time."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
a = bench.parse()
snnflow = importlib.import_module("snn_event-based_optical_flow_b200")
TrainWindow = importlib.import_module("snn_event-based_optical_flow_b200.train").TrainWindow
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = snnflow.LIFFireNet(dict(num_bins=2, encoding="cnt", base_num_channels=32, kernel_size=3, neuron_kwargs=dict(leak=(0.0, 1.0), thresh=(0.3, 0.1)))).to(dev)
cfg = {"loader": {"resolution": [128, 128]}, "loss": {"flow_regul_weight": 0.001}, "model": {"mask_output": False}}
tw = TrainWindow(net, snnflow.EventWarping(cfg, dev), torch.optim.Adam(net.parameters(), lr=2e-4, capturable=True))
host = [{k: v.pin_memory() for k, v in bench.make_window(a, i).items()} for i in range(4)]
tw.capture({k: v.to(dev) for k, v in host[0].items()})
def timeit(fn, n=10):
    fn(0); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def plain(i): return float(tw.step_graphed(host[i % 4]).item())
def pref(i):
    if getattr(tw, "_staged", None) is None: tw.prefetch(host[i % 4])
    l = tw.step_graphed(host[i % 4]); tw.prefetch(host[(i + 1) % 4]); return float(l.item())
def pref_after(i):
    if getattr(tw, "_staged", None) is None: tw.prefetch(host[i % 4])
    l = tw.step_graphed(host[i % 4]); v = float(l.item()); tw.prefetch(host[(i + 1) % 4]); return v
def h2d_only(i):
    tw._load(host[i % 4]); torch.cuda.synchronize()
def replay_only(i):
    tw._graph.replay(); torch.cuda.synchronize()
print("h2d only ms", timeit(h2d_only)); print("replay only ms", timeit(replay_only))
print("plain ms", timeit(plain)); print("prefetch-before-item ms", timeit(pref)); tw._staged = None
print("prefetch-after-item ms", timeit(pref_after))
