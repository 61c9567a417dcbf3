"""torchrun helper (not collected by pytest): the peer-memory all-reduce kernel against NCCL on every rank.
    torchrun --nproc-per-node 2 tests/_dp_peer_check.py"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
train = importlib.import_module("snn_event-based_optical_flow_b200.train")
torch.manual_seed(0)
params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in [(32, 32, 3, 3), (32, 1, 1), (2, 32, 1, 1), (7,)]]
red = train.PeerGradAllReduce(params)
ok = True
for step in range(5):
    g = torch.Generator(device="cpu").manual_seed(100 * step + rank)
    for p in params:
        p.grad = torch.randn(p.shape, generator=g).to(dev)
    ref = [p.grad.clone() for p in params]
    for r in ref:
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    red()
    torch.cuda.synchronize()
    for p, r in zip(params, ref):
        if not torch.allclose(p.grad, r, rtol=1e-6, atol=1e-6):
            ok = False
            print(f"rank {rank} step {step}: mismatch {float((p.grad - r).abs().max())}")
print(f"rank {rank}: peer all-reduce {'OK' if ok else 'FAILED'}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
