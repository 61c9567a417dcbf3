"""GPU: the fused gradient clipping + Adam update (snnflow_clip_adam) against torch.nn.utils.clip_grad_norm_ +
torch.optim.Adam - the pair train_flow.py:264-271 runs - on identical gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(32, 2, 3, 3), (32, 32, 3, 3), (32, 1, 1), (32, 1, 1), (2, 32, 1, 1), (2,), (7, 5)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in SHAPES]


@pytest.mark.parametrize("max_norm", [1.0, None])
def test_fused_clip_adam_matches_torch(max_norm):
    import snnflow_b200 as snnflow
    pa, pb = _params(3), _params(3)
    ref = torch.optim.Adam(pa, lr=2e-3)
    opt = snnflow.FusedClipAdam(pb, lr=2e-3, max_norm=max_norm)
    for a, b in zip(pa, pb):
        assert torch.equal(a, b), "flattening must preserve the parameter values"
    g = torch.Generator().manual_seed(4)
    for it in range(6):
        scale = [10.0, 1e-3, 1.0, 100.0, 1e-2, 0.3][it]          # norms far above and far below the clip threshold
        grads = [(torch.randn(s, generator=g) * scale).cuda() for s in SHAPES]
        skip = it == 2                                           # one step with a missing gradient
        for a, b, gr in zip(pa, pb, grads):
            a.grad, b.grad = gr.clone(), gr.clone()
        if skip:
            grads[-1].zero_()
            pa[-1].grad = torch.zeros_like(pa[-1])               # zero gradient == what the fused update assumes for None
            pb[-1].grad = None
        v0 = pb[0]._version
        total = torch.nn.utils.clip_grad_norm_(pa, max_norm) if max_norm else torch.stack([p.grad.norm() for p in pa]).norm()
        ref.step()
        opt.step()
        assert pb[0]._version > v0
        assert abs(float(opt.grad_norm) - float(total)) <= 1e-5 * float(total)
        for i, (a, b) in enumerate(zip(pa, pb)):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (it, i, float((a - b).abs().max()))
    assert int(opt.step_count) == 6


def test_fused_clip_adam_replays_in_a_cuda_graph():
    import snnflow_b200 as snnflow
    pa, pb = _params(5), _params(5)
    eager = snnflow.FusedClipAdam(pa, lr=1e-3, max_norm=1.0)
    graphed = snnflow.FusedClipAdam(pb, lr=1e-3, max_norm=1.0)
    static = [torch.zeros(s, device="cuda") for s in SHAPES]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graphed.step(static)                                     # warm-up on zero gradients: parameters unchanged
    torch.cuda.current_stream().wait_stream(side)
    eager.step(static)
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        graphed.step(static)                                     # capture only records: nothing runs here
    g = torch.Generator().manual_seed(6)
    for it in range(4):
        grads = [torch.randn(s, generator=g).cuda() * (5.0 if it % 2 else 0.05) for s in SHAPES]
        for s, gr in zip(static, grads):
            s.copy_(gr)
        cg.replay()
        eager.step(grads)
    torch.cuda.synchronize()
    assert int(graphed.step_count) == int(eager.step_count) == 5
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)                                 # same kernels, same order: bit-identical
