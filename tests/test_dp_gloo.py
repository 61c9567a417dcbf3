"""CPU, world_size 2 over gloo: the data-parallel gradient exchange (one flat SUM all-reduce, clip AFTER the
reduction) reproduces the single-process global-batch step for a sum-over-batch loss like the contrast loss
(loss/flow.py:228,261,291)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3, padding=1, bias=False), torch.nn.Tanh(),
                               torch.nn.Conv2d(4, 2, 1))


def _loss(model, x):
    return (model(x) ** 2).sum()          # SUM over the batch, like EventWarping


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from snnflow_b200.train import FlatGradAllReduce
    torch.set_num_threads(1)
    model = _make_model()
    x = torch.Generator().manual_seed(1)
    data = torch.randn(8, 2, 6, 6, generator=x)
    shard = data[rank * 4:(rank + 1) * 4]          # rank r owns samples [r*B/R, (r+1)*B/R)
    red = FlatGradAllReduce(model.parameters())
    _loss(model, shard).backward()
    red()
    norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    ret[rank] = (float(norm), [p.grad.clone() for p in model.parameters()])
    dist.destroy_process_group()


def test_flat_sum_allreduce_matches_global_batch():
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()     # no fork() of this multi-threaded process
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    model = _make_model()
    data = torch.randn(8, 2, 6, 6, generator=torch.Generator().manual_seed(1))
    _loss(model, data).backward()
    norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    for rank in (0, 1):
        n, grads = ret[rank]
        assert abs(n - float(norm)) < 1e-4 * float(norm)
        for g, p in zip(grads, model.parameters()):
            assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-7)


def test_single_process_is_a_noop():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from snnflow_b200.train import FlatGradAllReduce
    model = _make_model()
    _loss(model, torch.randn(2, 2, 6, 6)).backward()
    before = [p.grad.clone() for p in model.parameters()]
    FlatGradAllReduce(model.parameters())()
    for b, p in zip(before, model.parameters()):
        assert torch.equal(b, p.grad)
