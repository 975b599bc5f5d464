"""CPU: the oracle (oracle/head_oracle.py, oracle/salun_oracle.py) against the fixtures the reference
itself produced (oracle/gen_golden.py).  This is what pins the oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import head_oracle as HO
from oracle import salun_oracle as SO
from tests._golden import GOLDEN, MM_LOSSES, NLP_LOSSES, HeadCase, teacher_head_weight
from oracle import ref_harness as RH


def test_seeded_weights_match_fixture_checksums():
    z = np.load(GOLDEN / "head_mm_npo.npz")
    assert abs(RH.head_weight(1).double().abs().sum().item() - float(z["w_checksum"])) < 1e-6
    assert abs(teacher_head_weight().double().abs().sum().item() - float(z["wref_checksum"])) < 1e-6


def _check(case: HeadCase, fn, **kw):
    student, teacher, w, wr = case.oracle_inputs()
    loss, dh, dw = HO.loss_and_grads(fn, student, w, teacher=teacher, ref_weight=wr, beta=case.beta,
                                     loss_type=case.loss_type, **kw)
    if math.isnan(case.loss):
        assert math.isnan(loss.item())
        return
    assert abs(loss.item() - case.loss) <= 2e-5 * max(1.0, abs(case.loss)), (loss.item(), case.loss)
    for name, want in zip(case.student_splits, case.dhidden):
        got = dh[name]
        got = torch.zeros_like(want) if got is None else got
        scale = max(want.abs().max().item(), 1e-12)
        assert (got - want).abs().max().item() <= 2e-4 * scale + 1e-9, (case.loss_type, name)
    scale = max(case.dw_rows.abs().max().item(), 1e-12)
    assert (dw[case.probe_rows] - case.dw_rows).abs().max().item() <= 2e-4 * scale
    # column sums of dW vanish analytically (rows of softmax gradients sum to zero): absolute check
    assert (dw.double().sum(0) - case.dw_colsum).abs().max().item() <= 1e-3 * case.dw_sqnorm ** 0.5
    assert abs((dw.double() ** 2).sum().item() - case.dw_sqnorm) <= 1e-3 * case.dw_sqnorm


@pytest.mark.parametrize("loss_type", MM_LOSSES)
def test_mm_oracle_matches_reference(loss_type):
    _check(HeadCase("mm", loss_type), HO.mm_compute_loss)


def test_dpo_true_oracle_matches_the_expression_the_reference_discards():
    """head_mm_dpo_true.npz: the unmodified compute_loss("dpo") with F.logsigmoid recorded — value and gradients of
    -logsigmoid(0.1 * (pi - ref)).mean() (mm/trainer.py:390-394), which the reference overwrites before returning."""
    case = HeadCase("mm", "dpo_true")
    assert abs(case.beta - 0.1) < 1e-12
    _check(case, lambda student, weight, teacher=None, ref_weight=None, beta=1.0, loss_type=None, logits_dtype=None:
           HO.dpo_true(student, weight, teacher, ref_weight, beta=0.1, logits_dtype=logits_dtype))


@pytest.mark.parametrize("loss_type", NLP_LOSSES)
def test_nlp_oracle_matches_reference(loss_type):
    _check(HeadCase("nlp", loss_type), HO.nlp_compute_loss)


@pytest.mark.parametrize("kind", [HO.KIND_CE, HO.KIND_KL_FWD, HO.KIND_KL_REV, HO.KIND_NPO, HO.KIND_ENTROPY])
def test_rowloss_closed_forms_match_autograd(kind):
    """SURVEY.md App. A closed forms (what the CUDA row pass implements) against autograd, fp64."""
    g = torch.Generator().manual_seed(kind)
    z = torch.randn(7, 53, generator=g, dtype=torch.float64, requires_grad=True)
    zr = torch.randn(7, 53, generator=g, dtype=torch.float64)
    y = torch.randint(0, 53, (7,), generator=g)
    y[2] = -100
    beta = 0.7
    rows, dterm = HO.rowloss(kind, z, zr, y, beta)
    if kind == HO.KIND_CE:
        total = torch.nn.functional.cross_entropy(z, y, ignore_index=-100, reduction="sum")
    elif kind == HO.KIND_KL_FWD:
        total = HO.kl_batchmean(HO.logits2probs(z), HO.logits2probs(zr)) * 7
    elif kind == HO.KIND_KL_REV:
        total = HO.kl_batchmean(HO.logits2probs(zr), HO.logits2probs(z)) * 7
    elif kind == HO.KIND_NPO:
        total = HO.npo_mm(z, zr, beta) * (beta * 7 * 53 / 2)
    else:
        total = HO.neg_entropy_sum(z)
    # NPO: rows are log(1 + t); the gradient term s - p sum(s) is d/dz of (1/beta) sum log(1 + t)
    (total / beta if kind == HO.KIND_NPO else total).backward()
    assert abs(rows.sum().item() - total.item()) < 1e-10
    assert (dterm - z.grad).abs().max().item() < 1e-12


# ---- SalUn -----------------------------------------------------------------------------------------
def _salun_fixture():
    z = np.load(GOLDEN / "salun_tiny.npz")
    names = [str(n) for n in z["names"]]
    shapes = [tuple(int(x) for x in str(s).split(";")) for s in z["shapes"]]
    numels = [int(np.prod(s)) for s in shapes]
    grads = [torch.from_numpy(z[f"grad_{b}"].copy()).view(torch.bfloat16) for b in range(3)]
    return z, names, shapes, numels, grads


def test_salun_oracle_matches_reference_masks():
    z, names, shapes, numels, grads = _salun_fixture()
    acc = grads[0].clone()              # 0 + g
    for g in grads[1:]:
        acc = SO.accumulate(acc, g)
    offs = np.concatenate([[0], np.cumsum(numels)])
    named = {n: acc[offs[i]:offs[i + 1]].reshape(shapes[i]) for i, n in enumerate(names)}
    got = SO.masks(named)
    n_total = int(offs[-1])
    for th in SO.REFERENCE_THRESHOLDS:
        flat = np.concatenate([got[th][n].reshape(-1) for n in names])
        stable = np.unpackbits(z[f"mask_stable_{th}"])[:n_total].astype(np.int8)
        unstable = np.unpackbits(z[f"mask_unstable_{th}"])[:n_total].astype(np.int8)
        assert np.array_equal(flat, stable), f"th={th}: differs from the reference (stable sort)"
        # against the unmodified (unstable) reference: same popcount per tensor, differences only at ties
        for i, n in enumerate(names):
            sl = slice(offs[i], offs[i + 1])
            assert flat[sl].sum() == unstable[sl].sum() == int(th * numels[i])
            keys = SO.saliency_keys(named[n])
            tau, _ = SO.select(keys, th)
            diff = flat[sl] != unstable[sl]
            assert np.all(keys[diff] == tau)


def test_select_agrees_with_sort():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 40, size=1000).astype(np.uint16)
    for th in (0.0, 0.001, 0.1, 0.5, 0.999, 1.0):
        tau, ties = SO.select(keys, th)
        m = SO.mask_of_keys(keys, th)
        k = int(th * keys.size)
        assert m.sum() == k
        if k:
            assert (keys[m == 1] >= tau).all() and (keys[m == 0] <= tau).all()
            assert ((keys == tau) & (m == 1)).sum() == ties


def test_adamw_oracle_close_to_torch():
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(4096, generator=g).to(torch.bfloat16)
    mask = (torch.rand(4096, generator=g) < 0.5).to(torch.int8).numpy()
    p_ref = torch.nn.Parameter(p0.float().clone())
    opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, foreach=False)
    m = np.zeros(4096, np.float32)
    v = np.zeros(4096, np.float32)
    p = p0.float()
    for step in range(1, 4):
        grad = (0.01 * torch.randn(4096, generator=g)).to(torch.bfloat16)
        p_ref.grad = grad.float() * torch.from_numpy(mask).float()
        opt.step()
        # fp32 master trajectory of the oracle (re-rounding to bf16 each step would hide the comparison)
        newp, m, v = SO.adamw_step(p.to(torch.bfloat16) if step == 1 else p_bf16, grad, m, v, mask, step, lr=1e-3)
        p_bf16 = newp
        if step == 1:
            assert (newp.float() - p_ref.detach().to(torch.bfloat16).float()).abs().max().item() == 0.0
    st = opt.state[p_ref]
    assert np.allclose(m, st["exp_avg"].numpy(), rtol=1e-5, atol=1e-9)
    assert np.allclose(v, st["exp_avg_sq"].numpy(), rtol=1e-5, atol=1e-12)


def test_norm_oracles_match_the_torch_calls_the_reference_makes():
    """clip_grad_norm_ (HF Trainer default max_grad_norm = 1.0) and torch.norm(torch.cat(params), p) (mm/trainer.py:478-481)."""
    g = torch.Generator().manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(33, 7, generator=g)), torch.nn.Parameter(torch.randn(101, generator=g))]
    for p in ps:
        p.grad = (0.3 * torch.randn(p.shape, generator=g)).to(torch.bfloat16).float()
    before = [p.grad.to(torch.bfloat16) for p in ps]
    ss = SO.grad_sqnorm(before)
    coef = SO.clip_coef(ss, 1.0)
    total = torch.nn.utils.clip_grad_norm_(ps, 1.0)
    assert abs(ss ** 0.5 - total.item()) <= 1e-6 * total.item()
    assert float(coef) < 1.0
    for p, b in zip(ps, before):
        assert torch.allclose(SO.scale_grad(b, coef).float(), p.grad, atol=0, rtol=2 ** -8)
    assert SO.clip_coef(1e-4, 1.0) == 1.0
    a = [torch.randn(33, 7, generator=g).to(torch.bfloat16), torch.randn(9, generator=g).to(torch.bfloat16)]
    b = [x.clone() for x in a]
    b[0][3] += 0.5
    cat = torch.cat([(x - y).view(-1) for x, y in zip(a, b)])
    l1, l0 = SO.param_l1l0(a, b)
    assert l1 == torch.norm(cat.float(), p=1).item() and l0 == torch.norm(cat.float(), p=0).item() == 7.0
    l1z, l0z = SO.param_l1l0(a)
    assert abs(l1z - torch.norm(torch.cat([x.view(-1) for x in a]).float(), p=1).item()) <= 1e-5 * l1z
    assert l0z == sum(int((x != 0).sum()) for x in a)


def test_regulariser_oracle_matches_the_reference_compute_loss():
    """tests/golden/regulariser.npz = the reference's own compute_loss with and without l1_lambda / l0_lambda
    (oracle/gen_golden.py gen_regulariser): the oracle's L1 / L0 terms explain the loss difference, and the gradient
    difference is l1_lambda * sign(p - origin) — for l_norm_from = zero / init (mm) and the nlp variant (vs the oracle model)."""
    from oracle import ref_harness as RH
    z = np.load(GOLDEN / "regulariser.npz")
    l1, l0 = float(z["l1_lambda"]), float(z["l0_lambda"])
    cases = [("mm_zero", RH.tiny_llava, False), ("mm_init", RH.tiny_llava, True), ("nlp", RH.tiny_llama, True)]
    for tag, make, vs_teacher in cases:
        student = make(0)
        teacher = RH.perturbed_copy(student, 5)
        params = [p for p in student.parameters() if p.requires_grad]
        refs = [tp for p, tp in zip(student.parameters(), teacher.parameters()) if p.requires_grad] if vs_teacher else None
        n1, n0 = SO.param_l1l0(params, refs)
        want = float(z[f"{tag}_loss_reg"]) - float(z[f"{tag}_loss_plain"])
        got = l1 * n1 + l0 * n0
        assert abs(got - want) <= 2e-4 * abs(want), (tag, got, want)       # the reference sums 8.65 M fp32 terms in fp32
        named = dict(student.named_parameters())
        tnamed = dict(teacher.named_parameters())
        probes = [k for k in z.files if k.startswith(f"{tag}_dgrad::")]
        assert probes
        for k in probes:
            name = k.split("::", 1)[1]
            d = named[name].detach() - (tnamed[name].detach() if vs_teacher else 0)
            sign = torch.sign(d)
            sign = sign[:64] if sign.dim() == 2 else sign
            got_d = torch.from_numpy(z[k])
            assert torch.allclose(got_d, l1 * sign, atol=2e-6), (tag, name)   # fp32 gradient difference of two backward runs


def test_global_scope_oracle_is_one_ranking_over_all_tensors():
    """SO.masks_global (scope="global"; parity unpinned upstream, see its docstring): equals the pinned per-tensor rule applied
    to the concatenation; popcount = int(th * total); nested in th; differs from the per-tensor masks when tensors have
    different scales, coincides with them for a single tensor."""
    g = torch.Generator().manual_seed(3)
    named = {"a": (torch.randn(37, 11, generator=g) * 4).round().div(4).to(torch.bfloat16),          # ties
             "b": (1e-3 * torch.randn(1000, generator=g)).to(torch.bfloat16),                         # small scale
             "c": torch.zeros(13, dtype=torch.bfloat16), "d": torch.randn(5, generator=g).to(torch.bfloat16)}
    ths = (0.0, 0.1, 0.5, 0.77, 1.0)
    got = SO.masks_global(named, ths)
    total = sum(v.numel() for v in named.values())
    cat = np.concatenate([SO.saliency_keys(v) for v in named.values()])
    prev = None
    for th in ths:
        flat = np.concatenate([got[th][n].reshape(-1) for n in named])
        assert np.array_equal(flat, SO.mask_of_keys(cat, th))
        assert int(flat.sum()) == int(th * total)
        if prev is not None:
            assert np.all(flat >= prev)
        prev = flat
    per = SO.masks(named, (0.5,))[0.5]
    assert int(got[0.5]["b"].sum()) < int(per["b"].sum())            # the small-scale tensor loses under a global rank
    one = {"a": named["a"]}
    assert np.array_equal(SO.masks_global(one, (0.3,))[0.3]["a"], SO.masks(one, (0.3,))[0.3]["a"])


def test_adamw8_oracle_properties():
    """8-bit block-wise states (parity unpinned upstream: bitsandbytes is absent).  The oracle's restatement is internally
    consistent: codebooks sorted with 0 and 1; quantising exactly representable values (code * absmax) is the identity on the
    indices (idempotence of dequantise -> quantise); one 8-bit step stays close to the fp32-state step it approximates."""
    for signed in (True, False):
        code = SO.dynamic_map(signed)
        assert code.size == 256 and np.all(np.diff(code) > 0) and code[-1] == 1.0 and (code[0] == 0.0) != signed
    c1 = SO.dynamic_map(True)
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 256, 1024).astype(np.uint8)
    idx[::256] = 255                                             # every block holds its absmax exactly (code 1.0)
    absmax = np.asarray([0.5, 3.0, 1e-4, 7.0], np.float32)
    x = (c1[idx] * np.repeat(absmax, 256)).astype(np.float32)
    q, a = SO.adamw8_quantise(x, c1)
    assert np.array_equal(a, absmax) and np.array_equal(q, idx)
    n = 2048
    g = torch.Generator().manual_seed(1)
    p = torch.randn(n, generator=g).to(torch.bfloat16)
    gr = (0.01 * torch.randn(n, generator=g)).to(torch.bfloat16)
    z1 = int(np.nonzero(c1 == 0)[0][0])
    s1, s2 = np.full(n, z1, np.uint8), np.zeros(n, np.uint8)
    a1, a2 = np.zeros(8, np.float32), np.zeros(8, np.float32)
    p8, s1, s2, a1, a2 = SO.adamw8_step(p, gr, s1, s2, a1, a2, None, 1, lr=1e-3)
    p32, m, v = SO.adamw_step(p, gr, np.zeros(n, np.float32), np.zeros(n, np.float32), None, 1, lr=1e-3)
    # the states are quantised AFTER the update, so step 1 differs from the fp32-state step only through the reciprocal-multiply
    # spelling of the two divisions: at most the last bf16 bit of a parameter
    assert (p8.view(torch.int16).int() - p32.view(torch.int16).int()).abs().max().item() <= 1
    assert (p8 != p32).float().mean().item() <= 0.02
    m8 = SO.dynamic_map(True)[s1] * np.repeat(a1, 256)
    assert np.abs(m8 - m).max() <= 0.02 * np.abs(m).max()


def test_oracle_dpo_on_the_legacy_processor_fixture():
    """tests/golden/legacy_dpo.npz: the unmodified reference `compute_loss("dpo")` on a model that expands the image placeholder
    itself (T' = L + 575).  The oracle's `dpo` on the rows the reference's remove_image_tokens keeps gives the same loss and
    the same lm_head gradient."""
    z = np.load(GOLDEN / "legacy_dpo.npz")
    model = RH.LegacyExpandLM(0)
    idk = RH.legacy_split(4, 40, 2)
    captured = {}
    hook = model.lm_head.register_forward_pre_hook(lambda m, args: captured.__setitem__("h", args[0]))
    model(**idk)
    hook.remove()
    hidden = captured["h"].detach()                                       # [B, T', H]
    b, tp, h = hidden.shape
    l = idk["input_ids"].shape[1]
    rows = []
    for i in range(b):                                                     # remove_image_tokens as an index (mm/trainer_utils.py:48-70)
        pos = (idk["input_ids"][i] == RH.LEGACY_IMAGE_TOKEN).nonzero().flatten()
        s0 = int(pos[0]) if pos.numel() else 0
        rows.append(torch.cat([torch.arange(s0), torch.arange(tp - (l - s0), tp)]))
    cropped = torch.stack([hidden[i, r] for i, r in enumerate(rows)])
    w = model.lm_head.weight.detach()
    loss, dh, dw = HO.loss_and_grads(HO.mm_compute_loss, {"idk": HO.HeadInputs(cropped, idk["labels"])}, w, loss_type="dpo")
    assert abs(loss.item() - float(z["loss"])) <= 2e-5 * abs(float(z["loss"]))
    want_dw = torch.from_numpy(z["grad::lm_head.weight"])
    assert ((dw - want_dw).norm() / want_dw.norm()).item() <= 2e-4
