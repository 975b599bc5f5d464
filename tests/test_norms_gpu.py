"""GPU parity of the whole-model reductions around the update (csrc/norms.cu) against oracle/salun_oracle.py and
against the torch functions the reference calls: torch.nn.utils.clip_grad_norm_ (HF Trainer default max_grad_norm
= 1.0, mm/forget.py:151) and torch.norm(torch.cat(params), p = 1 / 0) (mm/trainer.py:451-481)."""
import types

import numpy as np
import pytest
import torch

from oracle import salun_oracle as SO

pytestmark = pytest.mark.gpu

SUM_RTOL = 1e-6  # double accumulation of fp32 8-element partial sums vs numpy double


def _rand_bf16(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(shape, generator=g)).to(torch.bfloat16)


@pytest.mark.parametrize("n", [1, 7, 8, 4099, 1_000_003])
def test_grad_sqnorm_and_coef(n):
    from multimodal_unlearning_b200 import norms
    dev = torch.device("cuda:0")
    gs = [_rand_bf16((n,), 1, 0.1), _rand_bf16((3, n), 2, 0.01)]
    g = torch.Generator().manual_seed(5)
    ms = [(torch.rand(n, generator=g) < 0.3).to(torch.int8), None]
    want = SO.grad_sqnorm(gs, [m.numpy() if m is not None else None for m in ms])
    got = norms.grad_sqnorm([x.to(dev) for x in gs], [m.to(dev) if m is not None else None for m in ms])
    assert got.dtype == torch.float64
    assert abs(got.item() - want) <= SUM_RTOL * want
    for max_norm in (1.0, 1e6):
        c = norms.clip_coef(got, max_norm).item()
        assert c == pytest.approx(float(SO.clip_coef(want, max_norm)), rel=1e-6)
        assert c <= 1.0


def test_clip_grad_norm_matches_torch():
    from multimodal_unlearning_b200 import norms
    dev = torch.device("cuda:0")
    shapes = [(257, 33), (1001,), (64, 64)]
    ours = [torch.nn.Parameter(torch.zeros(s, dtype=torch.bfloat16, device=dev)) for s in shapes]
    theirs = [torch.nn.Parameter(torch.zeros(s, dtype=torch.float32, device=dev)) for s in shapes]
    for i, s in enumerate(shapes):
        g = _rand_bf16(s, 10 + i, 0.05).to(dev)
        ours[i].grad = g.clone()
        theirs[i].grad = g.float()
    total = norms.clip_grad_norm_(ours, 1.0)
    want_total = torch.nn.utils.clip_grad_norm_(theirs, 1.0)
    assert total.item() == pytest.approx(want_total.item(), rel=1e-5)
    assert total.item() > 1.0  # the clip is active in this test
    for a, b in zip(ours, theirs):
        # fp32 gradient scaled by torch, rounded to bf16 == bf16 gradient scaled by the kernel (same coefficient to 1 ulp of fp32)
        diff = (a.grad.float() - b.grad.to(torch.bfloat16).float()).abs().max().item()
        assert diff <= 2 ** -8 * b.grad.abs().max().item()
    # bit-exact against the oracle given the device coefficient
    g0 = _rand_bf16(shapes[0], 10, 0.05)
    coef = SO.clip_coef(SO.grad_sqnorm([_rand_bf16(s, 10 + i, 0.05) for i, s in enumerate(shapes)]), 1.0)
    assert torch.equal(ours[0].grad.cpu().view(torch.int16), SO.scale_grad(g0, coef).view(torch.int16))


def test_masked_adamw_with_fused_clip_bit_exact():
    from multimodal_unlearning_b200 import norms, salun
    dev = torch.device("cuda:0")
    n = 100_003
    g = torch.Generator().manual_seed(3)
    p = torch.randn(n, generator=g).to(torch.bfloat16)
    mask = (torch.rand(n, generator=g) < 0.5).to(torch.int8)
    grad = (0.05 * torch.randn(n, generator=g)).to(torch.bfloat16)
    sumsq = norms.grad_sqnorm([grad.to(dev)], [mask.to(dev)])
    coef = norms.clip_coef(sumsq, 1.0)
    c = np.float32(coef.item())
    assert c < 1.0 and c == SO.clip_coef(SO.grad_sqnorm([grad], [mask.numpy()]), 1.0)
    masked = torch.where(mask.bool(), grad, torch.zeros_like(grad))
    want_p, want_m, want_v = SO.adamw_step(p, SO.scale_grad(masked, c), np.zeros(n, np.float32), np.zeros(n, np.float32),
                                           None, 1, lr=1e-3)
    pd, md, vd = p.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    salun.masked_adamw_flat(pd, grad.to(dev), md, vd, mask.to(dev), 1, lr=1e-3, grad_scale=coef)
    assert torch.equal(pd.cpu().view(torch.int16), want_p.view(torch.int16))
    assert np.array_equal(md.cpu().numpy().view(np.int32), want_m.view(np.int32))
    assert np.array_equal(vd.cpu().numpy().view(np.int32), want_v.view(np.int32))


def test_masked_adamw_optimizer_max_grad_norm_matches_torch():
    from multimodal_unlearning_b200 import salun
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    ws = [torch.randn(256, 128, device=dev).to(torch.bfloat16), torch.randn(777, device=dev).to(torch.bfloat16)]
    p1 = [torch.nn.Parameter(w.clone()) for w in ws]
    p2 = [torch.nn.Parameter(w.float().clone()) for w in ws]
    o1 = salun.MaskedAdamW(p1, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    o2 = torch.optim.AdamW(p2, lr=1e-3, weight_decay=0.01, foreach=False)
    for _ in range(3):
        for a, b in zip(p1, p2):
            g = (0.05 * torch.randn(a.shape, device=dev)).to(torch.bfloat16)
            a.grad = g.clone()
            b.grad = g.float()
        tn = torch.nn.utils.clip_grad_norm_(p2, 1.0)
        for b in p2:
            b.grad = b.grad.to(torch.bfloat16).float()  # the reference's gradients are bf16 tensors after mul_
        o1.step()
        o2.step()
        assert o1.last_grad_norm.item() == pytest.approx(tn.item(), rel=1e-5)
        for b in p2:
            b.data = b.data.to(torch.bfloat16).float()
    for a, b in zip(p1, p2):
        assert (a.float() - b).abs().max().item() <= 2 ** -7 * b.abs().max().item()


@pytest.mark.parametrize("with_ref", [False, True])
def test_param_l1l0(with_ref):
    from multimodal_unlearning_b200 import norms
    dev = torch.device("cuda:0")
    shapes = [(300, 70), (1001,), (8,), (3,)]
    ps = [_rand_bf16(s, 20 + i) for i, s in enumerate(shapes)]
    refs = None
    if with_ref:
        refs = [p.clone() for p in ps]
        refs[0][::3] += 0.25          # some equal, some different elements
        refs[1][5:50] = 0
    want = SO.param_l1l0(ps, refs)
    got = norms.param_l1l0([p.to(dev) for p in ps], [r.to(dev) for r in refs] if refs else None)
    assert got[1].item() == want[1]                       # the count is exact
    assert abs(got[0].item() - want[0]) <= SUM_RTOL * want[0]
    # and against the reference's own expression
    cat = torch.cat([(p - r).view(-1) if refs else p.view(-1) for p, r in zip(ps, refs or ps)]).float()
    assert got[0].item() == pytest.approx(torch.norm(cat, p=1).item(), rel=1e-5)
    assert got[1].item() == torch.norm(cat, p=0).item()


@pytest.mark.parametrize("origin", ["zero", "init"])
def test_weight_regulariser_autograd(origin):
    """loss + l1 * ||.||_1 + l0 * ||.||_0 and its gradient vs the reference's torch.cat expression."""
    from multimodal_unlearning_b200 import trainer as T
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    model = torch.nn.Sequential(torch.nn.Linear(64, 32), torch.nn.Linear(32, 8)).to(dev).to(torch.bfloat16)
    teacher = torch.nn.Sequential(torch.nn.Linear(64, 32), torch.nn.Linear(32, 8)).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        for p, tp in zip(model.parameters(), teacher.parameters()):
            tp.copy_(p)
            tp.view(-1)[::2] += 0.125
    l1, l0 = 1e-3, 1e-5
    base = torch.tensor(0.5, device=dev)
    fake = types.SimpleNamespace(l1_lambda=l1, l0_lambda=l0, l_norm_from=origin)
    got = T._regulariser(fake, model, teacher if origin == "init" else None, base)
    (got * 2.0).backward()
    got_g = [p.grad.clone() for p in model.parameters()]
    model.zero_grad()
    params = [(p - tp).view(-1) if origin == "init" else p.view(-1) for p, tp in zip(model.parameters(), teacher.parameters())]
    cat = torch.cat(params).float()
    want = base + l1 * torch.norm(cat, p=1) + l0 * torch.norm(cat, p=0)
    (want * 2.0).backward()
    assert got.item() == pytest.approx(want.item(), rel=1e-5)
    for p, g in zip(model.parameters(), got_g):
        assert torch.equal(g, p.grad), "L1 gradient = 2 * l1 * sign(p - ref), rounded to bf16"


def test_regulariser_errors_like_the_reference():
    from multimodal_unlearning_b200 import trainer as T
    dev = torch.device("cuda:0")
    model = torch.nn.Linear(8, 8).to(dev).to(torch.bfloat16)
    fake = types.SimpleNamespace(l1_lambda=1.0, l0_lambda=0.0, l_norm_from="init")
    with pytest.raises(ValueError, match="Teacher model is required"):
        T._regulariser(fake, model, None, torch.tensor(0.0, device=dev))
    nlp = types.SimpleNamespace(l1_lambda=1.0, l0_lambda=0.0)   # nlp trainer: no l_norm_from, oracle model required
    with pytest.raises(ValueError, match="required"):
        T._regulariser(nlp, model, None, torch.tensor(0.0, device=dev))
    with pytest.raises(RuntimeError, match="bf16"):
        T._regulariser(types.SimpleNamespace(l1_lambda=1.0, l0_lambda=0.0, l_norm_from="zero"), torch.nn.Linear(8, 8).to(dev),
                       None, torch.tensor(0.0, device=dev))
