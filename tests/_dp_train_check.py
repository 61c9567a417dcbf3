"""torchrun helper (run by tests/test_gpu_multi.py): N ranks, each owning B/N samples of ONE global window, must produce -
after the SUM all-reduce of the flat gradient - the gradients, the clipped norm and the updated parameters of a single
process that runs the whole global batch (SURVEY.md section 4 iii / 8e; loss/flow.py:228,261,291: the loss is a SUM
over samples).  Checked on every rank, for the peer-memory kernel inside the step and for the NCCL all-reduce.
    torchrun --nproc-per-node 2 tests/_dp_train_check.py"""
import copy
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import torch
import torch.distributed as dist

from snnflow_testutil import synth_window

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
snnflow = importlib.import_module("snn_event-based_optical_flow_b200")
train = importlib.import_module("snn_event-based_optical_flow_b200.train")

T, Bl, N, H, W, C = 4, 2, 400, 32, 64, 32
B = Bl * world
cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001}, "model": {"mask_output": False}}


def make_net():
    torch.manual_seed(0)
    net = snnflow.LIFFireNet(dict(num_bins=2, encoding="cnt", base_num_channels=C, kernel_size=3,
                                  neuron_kwargs=dict(leak=(0.0, 1.0), thresh=(0.3, 0.1))))
    with torch.no_grad():
        net.pred.conv2d.weight.mul_(20)
        for n, p in net.named_parameters():
            if n.endswith("weight"):
                p.copy_(torch.round(p * 4096) / 4096)
    return net.to(dev)


ok = True
window = synth_window(T, B, N, H, W, seed=77)                         # the same global window on every rank
glob = {k: v.to(dev) for k, v in window.items()}
shard = {k: v[:, rank * Bl:(rank + 1) * Bl].contiguous().to(dev) for k, v in window.items()}   # rank r: samples [r B/R, (r+1) B/R)

# single process, global batch
ref_net = make_net()
ref_opt = snnflow.FusedClipAdam(ref_net.parameters(), lr=1e-3, max_norm=1.0)
ref = train.TrainWindow(ref_net, snnflow.EventWarping(cfg, dev), ref_opt, clip_grad=1.0)
ref.reducer = type("NoReduce", (), {"__call__": lambda self, grads=None: None, "active": lambda self: False})()
ref_loss = ref._forward_backward(glob)
ref_grads = {n: p.grad.clone() for n, p in ref_net.named_parameters()}
ref._update()

for peer in (True, False):
    net = make_net()
    opt = snnflow.FusedClipAdam(net.parameters(), lr=1e-3, max_norm=1.0)
    tw = train.TrainWindow(net, snnflow.EventWarping(cfg, dev), opt, clip_grad=1.0, peer_allreduce=peer)
    name = type(tw.reducer).__name__
    loss = tw._forward_backward(shard)
    tw.reducer()
    total = loss.clone()
    dist.all_reduce(total)
    if abs(float(total) - float(ref_loss)) > 1e-5 * abs(float(ref_loss)):
        ok = False
        print(f"rank {rank} [{name}]: summed loss {float(total)} vs global-batch loss {float(ref_loss)}")
    for n, p in net.named_parameters():
        err = float((p.grad - ref_grads[n]).norm()) / (float(ref_grads[n].norm()) + 1e-30)
        if err > 2e-5:
            ok = False
            print(f"rank {rank} [{name}]: gradient of {n} differs from the global-batch gradient: rel {err:.3e}")
    tw._update()
    if abs(float(opt.grad_norm) - float(ref_opt.grad_norm)) > 1e-5 * float(ref_opt.grad_norm):
        ok = False
        print(f"rank {rank} [{name}]: clipped norm {float(opt.grad_norm)} vs {float(ref_opt.grad_norm)}")
    for (n, p), q in zip(net.named_parameters(), ref_net.parameters()):
        d = (p.detach() - q.detach()).abs()
        if float(d.max()) > 2.05e-3 or float((d <= 2e-5).float().mean()) < 0.99:
            ok = False
            print(f"rank {rank} [{name}]: updated {n} differs: max {float(d.max()):.3e}")
    # replicas stay identical: the reduced gradients are bit-identical on every rank
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        ok = False
        print(f"rank {rank} [{name}]: replicas diverged after one step")
    print(f"rank {rank}: {name}: sharded step == global-batch step: {'OK' if ok else 'FAILED'}")
# the direct step: gradients straight into the flat buffer, SUM over the ranks + clip + Adam in ONE launch (two-slot
# symmetric buffer, one flag barrier per step): three consecutive steps against the single-process global-batch run
ref_net2 = make_net()
ref2 = train.TrainWindow(ref_net2, snnflow.EventWarping(cfg, dev), snnflow.FusedClipAdam(ref_net2.parameters(), lr=1e-3, max_norm=1.0),
                         clip_grad=1.0)
ref2.reducer = type("NoReduce", (), {"__call__": lambda self, grads=None: None, "active": lambda self: False})()
net2 = make_net()
opt2 = snnflow.FusedClipAdam(net2.parameters(), lr=1e-3, max_norm=1.0)
tw2 = train.TrainWindow(net2, snnflow.EventWarping(cfg, dev), opt2, clip_grad=1.0, peer_allreduce=True)
direct_ok = True
for step in range(3):
    w = synth_window(T, B, N, H, W, seed=90 + step)
    glob2 = {k: v.to(dev) for k, v in w.items()}
    shard2 = {k: v[:, rank * Bl:(rank + 1) * Bl].contiguous().to(dev) for k, v in w.items()}
    assert ref2.direct_ok(glob2) and tw2.direct_ok(shard2), "direct step not available"
    ref2.step(glob2)
    tw2.step(shard2)
    if step == 0:
        if abs(float(opt2.grad_norm) - float(ref2.opt.grad_norm)) > 1e-5 * float(ref2.opt.grad_norm):
            direct_ok = False
            print(f"rank {rank} [direct]: norm {float(opt2.grad_norm)} vs {float(ref2.opt.grad_norm)}")
        for (n, p), q in zip(net2.named_parameters(), ref_net2.parameters()):
            d = (p.detach() - q.detach()).abs()
            if float(d.max()) > 2.05e-3 or float((d <= 2e-5).float().mean()) < 0.99:
                direct_ok = False
                print(f"rank {rank} [direct]: updated {n} differs: max {float(d.max()):.3e}")
    flat = torch.cat([p.detach().reshape(-1) for p in net2.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        direct_ok = False
        print(f"rank {rank} [direct]: replicas diverged after step {step}")
print(f"rank {rank}: one-launch DP update: sharded step == global-batch step: {'OK' if direct_ok else 'FAILED'}")
ok = ok and direct_ok
torch.cuda.synchronize()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
