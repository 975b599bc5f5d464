"""GPU parity tests of the fused loss head (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): loss rel <= 1e-3 (fp32 accumulation), bf16 gradients rel <= 2e-2
(norm-wise per tensor; plus an element-wise bound of a few bf16 ulps of the tensor's max).
The oracle is evaluated on the same bf16-rounded inputs and rounds logits to bf16 like a bf16
checkpoint's lm_head does (straight-through), so that the comparison isolates the kernels' arithmetic.
"""
import math
import types

import pytest
import torch

from oracle import head_oracle as HO
from tests._golden import MM_LOSSES, NLP_LOSSES, HeadCase

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-3
GRAD_RTOL = 2e-2
KINDS = [HO.KIND_CE, HO.KIND_KL_FWD, HO.KIND_KL_REV, HO.KIND_NPO, HO.KIND_ENTROPY]


def _rel(got: torch.Tensor, want: torch.Tensor) -> float:
    got, want = got.double().cpu(), want.double().cpu()
    return ((got - want).norm() / want.norm().clamp_min(1e-30)).item()


def _assert_grad(got, want, what):
    got, want = got.double().cpu(), want.double().cpu()
    assert torch.isfinite(got).all(), what
    assert _rel(got, want) <= GRAD_RTOL, (what, _rel(got, want))
    # element-wise: bf16 storage (2^-8 relative) plus accumulated rounding, relative to the tensor's scale
    assert (got - want).abs().max().item() <= 3e-2 * want.abs().max().item() + 1e-12, what


def _inputs(n, h, v, seed, dev):
    g = torch.Generator().manual_seed(seed)
    hid = torch.randn(n, h, generator=g).to(torch.bfloat16)
    w = (0.05 * torch.randn(v, h, generator=g)).to(torch.bfloat16)
    rh = (hid.float() + 0.3 * torch.randn(n, h, generator=g)).to(torch.bfloat16)
    rw = (w.float() + 0.01 * torch.randn(v, h, generator=g)).to(torch.bfloat16)
    y = torch.randint(0, v, (n,), generator=g)
    y[::5] = -100
    return hid, w, rh, rw, y


def _oracle_fused(kind, hid, w, rh, rw, y, beta, coef, reduction="reference", n_seqs=None, row_weight=None):
    """loss/dH/dW of one head invocation from the oracle's primitives, fp32 on the bf16-rounded inputs."""
    h = hid.float().requires_grad_(True)
    W = w.float().requires_grad_(True)
    z = HO.project(h, W, torch.bfloat16)
    zr = HO.project(rh.float(), rw.float(), torch.bfloat16).detach() if kind in (1, 2, 3) else None
    n, v = z.shape
    if reduction == "reference":
        if kind == HO.KIND_CE:
            loss = torch.nn.functional.cross_entropy(z, y, ignore_index=-100, reduction="mean")
        elif kind == HO.KIND_KL_FWD:
            loss = HO.kl_batchmean(HO.logits2probs(z), HO.logits2probs(zr))
        elif kind == HO.KIND_KL_REV:
            loss = HO.kl_batchmean(HO.logits2probs(zr), HO.logits2probs(z))
        elif kind == HO.KIND_NPO:
            loss = HO.npo_mm(z, zr, beta)
        else:
            loss = HO.neg_entropy_sum(z)
    else:  # seq_mean over CE rows, optional row weights
        rows = torch.nn.functional.cross_entropy(z, y, ignore_index=-100, reduction="none")
        if row_weight is not None:
            rows = rows * row_weight
        loss = rows.sum() / n_seqs
    loss = coef * loss
    loss.backward()
    return loss.detach(), h.grad, W.grad


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("v", [1000, 32000, 32064])
def test_projection_and_lse(v):
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h = 200, 128
    hid, w, *_ = _inputs(n, h, v, 1, dev)
    hid, w = hid.to(dev), w.to(dev)
    ld = (v + 63) // 64 * 64
    z = torch.full((n, ld), 7.0, dtype=torch.bfloat16, device=dev)
    part = torch.zeros(n, head.vtiles(v), 2, dtype=torch.float32, device=dev)
    head.project(hid, w, z, part)
    torch.cuda.synchronize()
    ref = hid.float().cpu() @ w.float().cpu().t()
    got = z[:, :v].float().cpu()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    m = part[..., 0].max(dim=1, keepdim=True).values
    lse = (m[:, 0] + torch.log((part[..., 1] * torch.exp(part[..., 0] - m)).sum(1))).cpu()
    want = torch.logsumexp(got.double(), dim=-1)
    assert (lse.double() - want).abs().max().item() <= 1e-4
    # statistics of materialised logits (teacher handed over as a tensor) agree with the epilogue's
    part2 = torch.zeros_like(part)
    head._ref_lse(z, part2, v)
    m2 = part2[..., 0].max(dim=1, keepdim=True).values
    lse2 = (m2[:, 0] + torch.log((part2[..., 1] * torch.exp(part2[..., 0] - m2)).sum(1))).cpu()
    assert (lse2.double() - want).abs().max().item() <= 1e-4


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("v", [1000, 32064])
def test_rowloss_vs_closed_forms(kind, v):
    """clear_head_rowloss in isolation against SURVEY.md App. A closed forms (oracle.rowloss, fp64)."""
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h = 96, 128
    beta = 0.7 if kind == HO.KIND_NPO else 1.0
    hid, w, rh, rw, y = _inputs(n, h, v, 10 + kind, dev)
    ld = (v + 63) // 64 * 64
    nt = head.vtiles(v)
    z = torch.zeros((n, ld), dtype=torch.bfloat16, device=dev)
    zr = torch.zeros((n, ld), dtype=torch.bfloat16, device=dev)
    part = torch.empty(n, nt, 2, dtype=torch.float32, device=dev)
    rpart = torch.empty(n, nt, 2, dtype=torch.float32, device=dev)
    head.project(hid.to(dev), w.to(dev), z, part)
    head.project(rh.to(dev), rw.to(dev), zr, rpart)
    z64, zr64 = z[:, :v].double().cpu(), zr[:, :v].double().cpu()
    want_rows, want_d = HO.rowloss(kind, z64, zr64, y, beta)
    rows = torch.empty(n, dtype=torch.float32, device=dev)
    needs_ref = kind in (1, 2, 3)
    # forward only leaves the logits untouched
    before = z.clone()
    head.rowloss(kind, z, part, rows, v, ref_logits=zr if needs_ref else None, ref_lse_part=rpart if needs_ref else None,
                 labels=y.to(dev) if kind == 0 else None, beta=beta, write_grad=False)
    assert torch.equal(before, z)
    assert (rows.double().cpu() - want_rows).abs().max().item() <= LOSS_RTOL * want_rows.abs().max().item()
    head.rowloss(kind, z, part, rows, v, ref_logits=zr if needs_ref else None, ref_lse_part=rpart if needs_ref else None,
                 labels=y.to(dev) if kind == 0 else None, beta=beta, grad_scale=0.5, write_grad=True)
    torch.cuda.synchronize()
    assert (rows.double().cpu() - want_rows).abs().max().item() <= LOSS_RTOL * want_rows.abs().max().item()
    _assert_grad(z[:, :v], 0.5 * want_d, f"dlogits kind {kind}")
    assert (z[:, v:] == 0).all(), "padding columns must be zeroed"


@pytest.mark.parametrize("v,n", [(13, 70), (1003, 70), (1000, 70), (1000, 700), (32064, 70), (32064, 333), (24576, 200), (40000, 70)])
def test_npo_beta_one_kernel(v, n):
    """The shipped NPO recipe (loss_beta = 1, config/mm/forget.yaml:40) runs its own row kernel (batched reciprocal /
    logarithm, s cached in shared memory): ragged vocabularies (13, 1003), the BASELINE vocabulary, and one wider than
    the shared-memory cache and the single-round-trip statistics merge (40000 = 157 column tiles).  Vocabularies that are
    Row counts below and above the SM count, a row that fits the s-cache exactly (24576)."""
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    h = 128
    hid, w, rh, rw, y = _inputs(n, h, v, 77, dev)
    ld = (v + 63) // 64 * 64
    nt = head.vtiles(v)
    z = torch.zeros((n, ld), dtype=torch.bfloat16, device=dev)
    zr = torch.zeros((n, ld), dtype=torch.bfloat16, device=dev)
    part = torch.empty(n, nt, 2, dtype=torch.float32, device=dev)
    rpart = torch.empty(n, nt, 2, dtype=torch.float32, device=dev)
    head.project(hid.to(dev), w.to(dev), z, part)
    head.project(rh.to(dev), rw.to(dev), zr, rpart)
    want_rows, want_d = HO.rowloss(HO.KIND_NPO, z[:, :v].double().cpu(), zr[:, :v].double().cpu(), y, 1.0)
    rows = torch.empty(n, dtype=torch.float32, device=dev)
    before = z.clone()
    head.rowloss(HO.KIND_NPO, z, part, rows, v, ref_logits=zr, ref_lse_part=rpart, beta=1.0, write_grad=False)
    assert torch.equal(before, z)
    assert (rows.double().cpu() - want_rows).abs().max().item() <= LOSS_RTOL * want_rows.abs().max().item()
    head.rowloss(HO.KIND_NPO, z, part, rows, v, ref_logits=zr, ref_lse_part=rpart, beta=1.0, grad_scale=0.5, write_grad=True)
    torch.cuda.synchronize()
    assert (rows.double().cpu() - want_rows).abs().max().item() <= LOSS_RTOL * want_rows.abs().max().item()
    _assert_grad(z[:, :v], 0.5 * want_d, f"dlogits NPO beta=1 v={v}")
    assert (z[:, v:] == 0).all(), "padding columns must be zeroed"


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("shape", [(300, 136, 1000), (512, 128, 32000), (130, 256, 32064)])
def test_fused_head_vs_oracle(kind, shape):
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h, v = shape
    beta = 0.7 if kind == HO.KIND_NPO else 1.0
    coef = -0.7
    hid, w, rh, rw, y = _inputs(n, h, v, 20 + kind, dev)
    want_loss, want_dh, want_dw = _oracle_fused(kind, hid, w, rh, rw, y, beta, coef)
    for block_rows, overlap in ((None, False), (128, False), (128, True)):
        res = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), kind, labels=y.to(dev), ref_hidden=rh.to(dev),
                                     ref_weight=rw.to(dev), beta=beta, coef=coef, block_rows=block_rows, overlap=overlap)
        torch.cuda.synchronize()
        assert abs(res.loss.item() - want_loss.item()) <= LOSS_RTOL * abs(want_loss.item()), (res.loss.item(), want_loss.item())
        _assert_grad(res.dhidden, want_dh, f"dH kind {kind} block {block_rows}")
        _assert_grad(res.dweight, want_dw, f"dW kind {kind} block {block_rows}")
        assert res.dweight.dtype == torch.bfloat16 and res.dhidden.dtype == torch.bfloat16


def test_teacher_as_materialised_logits():
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h, v = 256, 128, 1000
    hid, w, rh, rw, y = _inputs(n, h, v, 31, dev)
    a = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_KL_FWD, ref_hidden=rh.to(dev), ref_weight=rw.to(dev))
    zr = (rh.to(dev).float() @ rw.to(dev).float().t()).to(torch.bfloat16)
    b = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_KL_FWD, ref_logits=zr.contiguous())
    assert abs(a.loss.item() - b.loss.item()) <= 1e-3 * abs(a.loss.item())
    assert _rel(b.dweight, a.dweight) < 1e-2


def test_dweight_accumulates_in_fp32_across_calls():
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h, v = 256, 128, 1000
    hid, w, rh, rw, y = _inputs(n, h, v, 32, dev)
    acc = torch.zeros(v, h, dtype=torch.float32, device=dev)
    one = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, labels=y.to(dev))
    for _ in range(2):
        head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, labels=y.to(dev), dweight_accum=acc, block_rows=128)
    assert _rel(acc, 2 * one.dweight.float()) < 1e-2


def test_seq_mean_and_row_weights():
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    b, t, h, v = 4, 40, 128, 1000
    hid, w, rh, rw, y = _inputs(b * t, h, v, 33, dev)
    rwt = torch.rand(b).repeat_interleave(t)
    want_loss, want_dh, want_dw = _oracle_fused(0, hid, w, rh, rw, y, 1.0, -1.0, "seq_mean", b, rwt)
    res = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, labels=y.to(dev), coef=-1.0, reduction="seq_mean",
                                 n_seqs=b, row_weight=rwt.to(dev))
    assert abs(res.loss.item() - want_loss.item()) <= LOSS_RTOL * abs(want_loss.item())
    _assert_grad(res.dhidden, want_dh, "dH seq_mean")
    _assert_grad(res.dweight, want_dw, "dW seq_mean")
    # forward-only per-sequence sums == get_batch_loss
    z = HO.project(hid.float(), w.float(), torch.bfloat16).view(b, t, v)
    labels = torch.cat([torch.full((b, 1), -100), y.view(b, t)[:, :-1]], dim=1)  # unshifted labels whose shift is y
    want = HO.get_batch_loss(z, labels)
    # get_batch_loss drops the last row; y's last column must then be ignored for equality
    y2 = y.view(b, t).clone()
    y2[:, -1] = -100
    seq = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, labels=y2.reshape(-1).to(dev), reduction="seq_mean",
                                 n_seqs=b, need_dhidden=False, need_dweight=False).seq_loss
    assert (seq.cpu() - want).abs().max().item() <= LOSS_RTOL * want.abs().max().item()


def test_edge_cases():
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h, v = 64, 128, 1000
    hid, w, rh, rw, y = _inputs(n, h, v, 34, dev)
    # all rows ignored: HF's mean CE is NaN (0/0); gradients are zero rows times inf -> the reference gives NaN too
    res = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, labels=torch.full((n,), -100, device=dev))
    assert math.isnan(res.loss.item())
    # one single valid row
    y1 = torch.full((n,), -100)
    y1[17] = 3
    want_loss, want_dh, want_dw = _oracle_fused(0, hid, w, rh, rw, y1, 1.0, 1.0)
    res = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, labels=y1.to(dev))
    assert abs(res.loss.item() - want_loss.item()) <= LOSS_RTOL * abs(want_loss.item())
    assert (res.dhidden[:17] == 0).all() and (res.dhidden[18:] == 0).all()
    _assert_grad(res.dhidden, want_dh, "dH single row")
    # a single row, tiny vocab
    res = head.head_loss_fwd_bwd(hid[:1].to(dev), w[:13].contiguous().to(dev), HO.KIND_ENTROPY)
    want_loss, _, _ = _oracle_fused(4, hid[:1], w[:13], rh[:1], rw[:13], y[:1], 1.0, 1.0)
    assert abs(res.loss.item() - want_loss.item()) <= LOSS_RTOL * abs(want_loss.item())
    # errors: no silent fallback
    with pytest.raises(RuntimeError):
        head.head_loss_fwd_bwd(hid, w, HO.KIND_CE, labels=y)  # CPU tensors
    with pytest.raises(RuntimeError):
        head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_NPO)  # teacher missing
    with pytest.raises(RuntimeError):
        head.head_loss_fwd_bwd(hid.to(dev).float(), w.to(dev), HO.KIND_CE, labels=y.to(dev))  # dtype


# --------------------------------------------------------------------------------------------------
# the reference's own outputs (tests/golden, produced by oracle/gen_golden.py) through the recipes
# --------------------------------------------------------------------------------------------------
def _run_recipe(case: HeadCase, dev):
    from multimodal_unlearning_b200 import trainer as T
    hid = {n: h.to(dev).to(torch.bfloat16).requires_grad_(True) for n, h in zip(case.student_splits, case.hidden)}
    lab = {n: y.to(dev) for n, y in zip(case.student_splits, case.labels)}
    ref = {n: h.to(dev).to(torch.bfloat16) for n, h in zip(case.teacher_splits, case.ref_hidden)}
    w = case.weight.to(dev).to(torch.bfloat16).requires_grad_(True)
    wr = case.ref_weight.to(dev).to(torch.bfloat16)

    class FakeModel(torch.nn.Module):
        """Trunk stand-in: returns the recorded hidden states of the split whose input_ids tag it carries."""

        def __init__(self, table, weight):
            super().__init__()
            self.table = table
            self.lm_head = torch.nn.Linear(weight.shape[1], weight.shape[0], bias=False, device=dev, dtype=torch.bfloat16)
            self.lm_head.weight = torch.nn.Parameter(weight.detach().clone())
            self.config = types.SimpleNamespace(image_token_index=None)

        def get_output_embeddings(self):
            return self.lm_head

        def forward(self, input_ids=None, attention_mask=None, **kw):
            name = self.names[int(input_ids.flatten()[0])]
            return types.SimpleNamespace(logits=self.lm_head(self.table[name]))

    names = sorted(set(case.student_splits) | set(case.teacher_splits))
    student = FakeModel(hid, w)
    student.names = names
    teacher = FakeModel(ref, wr)
    teacher.names = names
    for p in teacher.parameters():
        p.requires_grad_(False)

    def tag(name, like):
        return torch.full_like(like, names.index(name))

    if case.trainer == "mm":
        inputs = {}
        rnd = [n for n in case.student_splits if n.startswith("random")]
        for n in names:
            if n.startswith("random"):
                continue
            y = lab.get(n, torch.zeros(2, case.hidden[0].shape[1], dtype=torch.int64, device=dev))
            inputs[n] = {"input_ids": tag(n, y), "attention_mask": torch.ones_like(y), "labels": y}
        if "retain" not in inputs:  # the reference always moves inputs["retain"] to the device
            y = next(iter(lab.values()))
            inputs["retain"] = {"input_ids": tag(names[0], y), "attention_mask": torch.ones_like(y), "labels": y}
        if rnd:
            ys = torch.stack([lab[n] for n in rnd], dim=1)  # [B, K, L]
            ids = torch.stack([tag(n, lab[n]) for n in rnd], dim=1)
            inputs["random"] = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ys}
        true_dpo = case.loss_type == "dpo_true"   # the flagged extension: same `dpo` recipe, sigmoid log-ratio kept
        fake = types.SimpleNamespace(loss_type="dpo" if true_dpo else case.loss_type.lower(), teacher_model=teacher,
                                     loss_beta=case.beta, l1_lambda=0.0, l0_lambda=0.0, l_norm_from="zero", dpo_true=true_dpo,
                                     loss_args={"bad_weight": 2.5, "random_weight": 1.0, "normal_weight": 0.5})
        loss = T.MMTrainerForgetting.compute_loss(fake, student, inputs)
    else:
        def triple(n):
            y = lab[n]
            return (tag(n, y), y, torch.ones_like(y))
        if case.loss_type == "idk":
            # one forward over cat(forget, retain): the fixture holds it as a single 4-sequence split
            y = lab["forget+retain"]
            half = y.shape[0] // 2
            student.table["forget+retain"] = hid["forget+retain"]
            f = (tag("forget+retain", y[:half]), y[:half], torch.ones_like(y[:half]))
            r = (tag("forget+retain", y[half:]), y[half:], torch.ones_like(y[half:]))
            inputs = [f, r]
        else:
            first = case.student_splits[0]
            inputs = [triple("forget") if "forget" in lab else triple(first),
                      triple("retain") if "retain" in lab else triple(first)]
            if "idk" in lab:
                inputs.append(triple("idk"))
        fake = types.SimpleNamespace(loss_type=case.loss_type, oracle_model=teacher, loss_beta=case.beta, l1_lambda=0.0,
                                     l0_lambda=0.0)
        loss = T.CustomTrainerForgetting.compute_loss(fake, student, inputs)
    loss.backward()
    return loss, hid, student.lm_head.weight.grad


def _oracle_on_bf16(case: HeadCase):
    student, teacher, w, wr = case.oracle_inputs(cast=torch.bfloat16)
    fn = HO.mm_compute_loss if case.trainer == "mm" else HO.nlp_compute_loss
    if case.loss_type == "dpo_true":
        fn = lambda student, weight, teacher=None, ref_weight=None, beta=1.0, loss_type=None, logits_dtype=None: \
            HO.dpo_true(student, weight, teacher, ref_weight, beta=0.1, logits_dtype=logits_dtype)
    if case.trainer == "nlp" and case.loss_type == "npo":
        fn = HO.mm_compute_loss  # App. B #3: the nlp formula is NaN by construction; the product evaluates the mm one
    return HO.loss_and_grads(fn, student, w, teacher=teacher, ref_weight=wr, beta=case.beta, loss_type=case.loss_type,
                             logits_dtype=torch.bfloat16)


@pytest.mark.parametrize("trainer,loss_type", [("mm", l) for l in MM_LOSSES] + [("mm", "dpo_true")] + [("nlp", l) for l in NLP_LOSSES])
def test_recipes_against_reference_fixtures(trainer, loss_type):
    case = HeadCase(trainer, loss_type)
    dev = torch.device("cuda:0")
    loss, hid, dw = _run_recipe(case, dev)
    want_loss, want_dh, want_dw = _oracle_on_bf16(case)
    assert abs(loss.item() - want_loss.item()) <= LOSS_RTOL * max(abs(want_loss.item()), 1e-3), (loss.item(), want_loss.item())
    for n in case.student_splits:
        key = n
        if n.startswith("random"):
            key = n
        want = want_dh.get(key)
        got = hid[n].grad
        if want is None or want.abs().max() == 0:
            assert got is None or got.abs().max().item() == 0
            continue
        _assert_grad(got.reshape(want.shape), want, f"{loss_type} dH[{n}]")
    _assert_grad(dw, want_dw, f"{loss_type} dW")
    # and against what the reference itself returned in fp32 on the unrounded inputs (looser: bf16 inputs)
    if not math.isnan(case.loss):
        assert abs(loss.item() - case.loss) <= 2e-2 * max(abs(case.loss), 1.0), (loss.item(), case.loss)
        got_rows = dw[case.probe_rows.to(dev)].float().cpu()
        assert _rel(got_rows, case.dw_rows) <= 6e-2


# --------------------------------------------------------------------------------------------------
# BASELINE.json sizes: full-size checks against a chunked fp32 torch restatement on the GPU + properties
# --------------------------------------------------------------------------------------------------
def _torch_reference_fullsize(kind, hid, w, rh, rw, y, beta, chunk=1024, n_seqs=None, coef=1.0):
    """The reference's arithmetic (SURVEY.md App. A closed forms of mm/trainer.py:168-449) restated with cuBLAS fp32 on
    the GPU, in row chunks: (loss fp64, dH fp32, dW fp32).  n_seqs: `get_batch_loss` semantics (per-sequence sums, mean
    over sequences) instead of the HF mean over valid rows."""
    n, h = hid.shape
    v = w.shape[0]
    dH = torch.empty(n, h, dtype=torch.float32, device=hid.device)
    dW = torch.zeros(v, h, dtype=torch.float32, device=hid.device)
    wf, rwf = w.float(), rw.float()
    total = torch.zeros((), dtype=torch.float64, device=hid.device)
    m = (y != -100).sum()
    for r0 in range(0, n, chunk):
        hb = hid[r0:r0 + chunk].float()
        z = (hb @ wf.t()).to(torch.bfloat16).float()
        lp = torch.log_softmax(z, -1)
        p = lp.exp()
        if kind in (HO.KIND_KL_FWD, HO.KIND_KL_REV, HO.KIND_NPO):
            zr = (rh[r0:r0 + chunk].float() @ rwf.t()).to(torch.bfloat16).float()
            lq = torch.log_softmax(zr, -1)
            q = lq.exp()
        if kind == HO.KIND_CE:
            yy = y[r0:r0 + chunk]
            valid = (yy != -100)
            rows = -lp.gather(1, yy.clamp(min=0)[:, None])[:, 0] * valid
            d = p.clone()
            d[torch.arange(d.shape[0], device=d.device), yy.clamp(min=0)] -= 1
            denom = n_seqs if n_seqs else m
            d = d * valid[:, None] / denom
            total += rows.double().sum() / denom
        elif kind == HO.KIND_KL_FWD:   # kl_div(student_logp, teacher_logp, batchmean, log_target): all rows / N
            total += (q * (lq - lp)).double().sum() / n
            d = (p - q) / n
        elif kind == HO.KIND_KL_REV:   # kl_div(teacher_logp, student_logp, ...): sum p (lp - lq) / N, gradient through p
            t = lp - lq
            k = (p * t).sum(-1, keepdim=True)
            total += k.double().sum() / n
            d = p * (t - k) / n
        elif kind == HO.KIND_ENTROPY:  # sum p log p (raw sum, mm/trainer.py:195-196)
            e = (p * lp).sum(-1, keepdim=True)
            total += e.double().sum()
            d = p * (lp - e)
        else:  # NPO
            t = (p / (q + 1e-8)).pow(beta)
            s = t / (1 + t)
            total += torch.log1p(t).double().sum() * (2.0 / (beta * n * v))
            d = (s - p * s.sum(-1, keepdim=True)) * (2.0 / (n * v))
        d = d * coef
        dH[r0:r0 + chunk] = d @ wf
        dW += d.t() @ hb
    return total * coef, dH, dW


def _fullsize_inputs(n, h, v, seed=5):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    hid = torch.randn(n, h, generator=g, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(v, h, generator=g, device=dev)).to(torch.bfloat16)
    rh = (hid.float() + 0.05 * torch.randn(n, h, generator=g, device=dev)).to(torch.bfloat16)
    y = torch.randint(0, v, (n,), generator=g, device=dev)
    y[1023::1024] = -100
    return hid, w, rh, y


def _oracle_rows_check(kind, res, hid, w, rh, rw, y, beta, scale, n_rows=300, seed=9):
    """The oracle proper (CPU, fp64 closed forms of oracle/head_oracle.py) on `n_rows` random rows: the loss term of a row
    and its dH row only need that row's logits."""
    dev = hid.device
    g = torch.Generator().manual_seed(seed)
    idx = torch.unique(torch.cat([torch.tensor([0, 1, 511, hid.shape[0] - 1]), torch.randint(0, hid.shape[0], (n_rows,), generator=g)]))
    di = idx.to(dev)
    z = HO.project(hid[di].float().cpu(), w.float().cpu(), torch.bfloat16)
    zr = HO.project(rh[di].float().cpu(), rw.float().cpu(), torch.bfloat16)
    rows, dterm = HO.rowloss(kind, z.double(), zr.double(), y[di].cpu(), beta)
    got_rows = res.loss_rows[di].double().cpu()
    # the loss tolerance applies to the sum; a single KL row is a sum of signed terms q (lq - lp) that nearly cancel (2e-3 out
    # of terms of order 1e-1 here), so its own rounding error is bounded against a 5x wider band
    assert abs((got_rows - rows).sum().item()) <= LOSS_RTOL * abs(rows.sum().item())
    per_row = 5 * LOSS_RTOL if kind in (HO.KIND_KL_FWD, HO.KIND_KL_REV) else LOSS_RTOL
    assert (got_rows - rows).abs().max().item() <= per_row * rows.abs().max().item()
    want_rows_dh = (dterm * scale) @ w.double().cpu()
    assert _rel(res.dhidden[di], want_rows_dh) <= GRAD_RTOL, _rel(res.dhidden[di], want_rows_dh)
    return idx.numel()


FULLSIZE = [  # (kind, beta, V): every shipped row kind at N = 8192 rows, H = 4096
    (HO.KIND_NPO, 1.0, 32064),      # BASELINE configs[1]: the beta = 1 fast kernel
    (HO.KIND_NPO, 0.7, 32064),      # generic-beta kernel (XU-bound)
    (HO.KIND_CE, 1.0, 32000),       # configs[4]: Llama-2 vocabulary
    (HO.KIND_CE, 1.0, 32064),       # LLaVA vocabulary (192 x 167: ragged last 256-column tile)
    (HO.KIND_KL_FWD, 1.0, 32064),   # KL-retain split of configs[2]
    (HO.KIND_KL_REV, 1.0, 32064),   # scrub / llmu
    (HO.KIND_ENTROPY, 1.0, 32064),  # grad_diff_forget_entropy_*
]


@pytest.mark.parametrize("kind,beta,v", FULLSIZE)
def test_baseline_size_parity(kind, beta, v):
    """Every row kind at the BASELINE.json size (bs 8 x seq 1024 = 8192 rows, H = 4096, the LLaVA-1.5-7B / Llama-2-7B head):
    the whole invocation against the reference's arithmetic in fp32 (cuBLAS, on the GPU), and >= 256 random rows against the
    CPU oracle's fp64 closed forms."""
    from multimodal_unlearning_b200 import head
    n, h = 8192, 4096
    hid, w, rh, y = _fullsize_inputs(n, h, v)
    rw = w
    res = head.head_loss_fwd_bwd(hid, w, kind, labels=y, ref_hidden=rh, ref_weight=rw, beta=beta)
    want_loss, want_dh, want_dw = _torch_reference_fullsize(kind, hid, w, rh, rw, y, beta)
    assert abs(res.loss.item() - want_loss.item()) <= LOSS_RTOL * abs(want_loss.item()), (res.loss.item(), want_loss.item())
    assert _rel(res.dhidden, want_dh) <= GRAD_RTOL
    assert _rel(res.dweight, want_dw) <= GRAD_RTOL
    # size-independent properties: rows of a softmax-type gradient sum to zero, so dW's column sums vanish
    # and dH is orthogonal to nothing in particular but must be finite; ignored CE rows give exactly zero dH.
    col = res.dweight.float().sum(0)
    assert col.abs().max().item() <= 2e-2 * res.dweight.float().abs().sum(0).max().item()
    if kind == HO.KIND_CE:
        assert (res.dhidden[1023::1024] == 0).all()
    scale = {HO.KIND_CE: 1.0 / (y != -100).sum().item(), HO.KIND_KL_FWD: 1.0 / n, HO.KIND_KL_REV: 1.0 / n,
             HO.KIND_ENTROPY: 1.0}.get(kind, 2.0 / (n * v))
    assert _oracle_rows_check(kind, res, hid, w, rh, rw, y, beta, scale) >= 256


def test_baseline_size_dpo_idk_labels_compacted():
    """BASELINE configs[3]: `dpo` (shipped semantics: mean over sequences of the per-sequence CE sums on the IDK split) at
    bs 8 x seq 1024, H = 4096, V = 32064 with IDK-length answers (5-15 valid labels per sequence), through compact_ignored
    (device-side row compaction): against the fp32 restatement over ALL rows and the CPU oracle on the valid rows."""
    from multimodal_unlearning_b200 import head
    n, h, v, b = 8192, 4096, 32064, 8
    hid, w, rh, _ = _fullsize_inputs(n, h, v, seed=6)
    g = torch.Generator().manual_seed(3)
    y = torch.full((b, n // b), -100, dtype=torch.int64)
    for i in range(b):
        k = int(torch.randint(5, 16, (1,), generator=g))
        y[i, -k:] = torch.randint(0, v, (k,), generator=g)
    y = head.shift_labels(y).to(hid.device)
    res = head.head_loss_fwd_bwd(hid, w, HO.KIND_CE, labels=y, reduction="seq_mean", n_seqs=b, compact_ignored=True)
    want_loss, want_dh, want_dw = _torch_reference_fullsize(HO.KIND_CE, hid, w, rh, w, y, 1.0, n_seqs=b)
    assert abs(res.loss.item() - want_loss.item()) <= LOSS_RTOL * abs(want_loss.item())
    assert _rel(res.dhidden, want_dh) <= GRAD_RTOL and _rel(res.dweight, want_dw) <= GRAD_RTOL
    assert (res.dhidden[y == -100] == 0).all()
    seq = res.loss_rows.view(b, -1).sum(1)
    assert torch.allclose(res.seq_loss, seq, rtol=1e-5)
    valid = (y != -100).nonzero().flatten()
    z = HO.project(hid[valid].float().cpu(), w.float().cpu(), torch.bfloat16)
    rows, dterm = HO.rowloss(HO.KIND_CE, z.double(), z.double(), y[valid].cpu(), 1.0)
    assert (res.loss_rows[valid].double().cpu() - rows).abs().max().item() <= LOSS_RTOL * rows.abs().max().item()
    assert _rel(res.dhidden[valid], (dterm / b) @ w.double().cpu()) <= GRAD_RTOL


def test_baseline_size_two_split_fp32_accumulation():
    """BASELINE configs[2]: grad_diff_forget_ce_retain_kl at seq 2048 (bs 4: 8192 rows per split) — two head invocations per
    step whose dW contributions accumulate in ONE fp32 buffer inside the contractions' epilogues (dweight_accum):
    -beta * CE(forget) + KL_fwd(retain)."""
    from multimodal_unlearning_b200 import head
    n, h, v, beta = 8192, 4096, 32064, 1.0
    hf, w, _, yf = _fullsize_inputs(n, h, v, seed=7)
    yf.view(4, 2048)[:, :1300] = -100                      # image + question part ignored
    hr, _, rr, _ = _fullsize_inputs(n, h, v, seed=8)
    dw = torch.zeros(v, h, dtype=torch.float32, device=w.device)
    a = head.head_loss_fwd_bwd(hf, w, HO.KIND_CE, labels=yf, coef=-beta, dweight_accum=dw, compact_ignored=True)
    b = head.head_loss_fwd_bwd(hr, w, HO.KIND_KL_FWD, ref_hidden=rr, ref_weight=w, dweight_accum=dw)
    assert a.dweight is dw and b.dweight is dw
    la, dha, dwa = _torch_reference_fullsize(HO.KIND_CE, hf, w, hf, w, yf, 1.0, coef=-beta)
    lb, dhb, dwb = _torch_reference_fullsize(HO.KIND_KL_FWD, hr, w, rr, w, yf, 1.0)
    total, want = a.loss.item() + b.loss.item(), la.item() + lb.item()
    assert abs(total - want) <= LOSS_RTOL * abs(want)
    assert _rel(a.dhidden, dha) <= GRAD_RTOL and _rel(b.dhidden, dhb) <= GRAD_RTOL
    assert _rel(dw, dwa + dwb) <= GRAD_RTOL
    assert _rel(dw, dwa + dwb) <= 5e-3, "fp32 accumulation across splits: only the bf16 rounding of dlogits remains"


@pytest.mark.parametrize("k_splits", [1, 2, 3])
@pytest.mark.parametrize("layout", ["k_mn", "mn_mn"])
def test_gemm_split_k_reduce_add(k_splits, layout):
    """clear_gemm_bf16 with the fp32 reduce-add epilogue and K cut into work items (the dH / dW operand layouts)."""
    import ctypes as C

    from multimodal_unlearning_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    m, n, k = 384, 520, 1000
    g = torch.Generator().manual_seed(40 + k_splits)
    A = torch.randn(m, k, generator=g).to(torch.bfloat16)
    B = torch.randn(n, k, generator=g).to(torch.bfloat16)
    pad = lambda x: (x + 63) // 64 * 64
    if layout == "k_mn":   # A stored [m, k], B stored [k, n]
        a_st = torch.zeros(m, pad(k), dtype=torch.bfloat16); a_st[:, :k] = A
        b_st = torch.zeros(k, pad(n), dtype=torch.bfloat16); b_st[:, :n] = B.t()
        a_major, b_major = 0, 1
    else:                  # A stored [k, m], B stored [k, n]
        a_st = torch.zeros(k, pad(m), dtype=torch.bfloat16); a_st[:, :m] = A.t()
        b_st = torch.zeros(k, pad(n), dtype=torch.bfloat16); b_st[:, :n] = B.t()
        a_major, b_major = 1, 1
    a_st, b_st = a_st.to(dev), b_st.to(dev)
    D = torch.full((m, pad(n)), 0.5, dtype=torch.float32, device=dev)   # accumulate on top of existing contents
    args = _lib.GemmArgs(a=a_st.data_ptr(), b=b_st.data_ptr(), d=D.data_ptr(), lse_part=None, lda=a_st.stride(0),
                         ldb=b_st.stride(0), ldd=D.stride(0), m=m, n=n, k=k, a_major=a_major, b_major=b_major, d_f32=1,
                         accumulate=1, cta_group=0, raster_m_fast=-1, k_splits=k_splits)
    _lib.check(lib.clear_gemm_bf16(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "clear_gemm_bf16")
    torch.cuda.synchronize()
    want = A.float() @ B.float().t() + 0.5
    got = D[:, :n].cpu()
    assert (got - want).abs().max().item() <= 2e-4 * want.abs().max().item()
    assert (D[:, n:] == 0.5).all()


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("layout", ["k_k", "k_mn", "mn_mn"])
@pytest.mark.parametrize("raster", [-1, 0, 1])
def test_gemm_layouts_groups_and_tile_orders(cta_group, layout, raster):
    """Every operand layout the head uses, on one CTA and on CTA pairs (the claimed-work ring is mirrored into the peer
    CTA only in the paired case), under the three tile orders (traffic model, N-fast, M-fast).  Shapes leave ragged
    edge tiles in both dimensions and more work items than CTA groups would get from one wave."""
    import ctypes as C

    from multimodal_unlearning_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    m, n, k = 2100, 5000, 200          # 9 x 20 (pairs) / 17 x 20 (single) output tiles
    g = torch.Generator().manual_seed(7)
    A = torch.randn(m, k, generator=g).to(torch.bfloat16)
    B = torch.randn(n, k, generator=g).to(torch.bfloat16)
    pad = lambda x: (x + 63) // 64 * 64

    def store(x, major):  # x is [mn, k]
        if major == 0:
            st = torch.zeros(x.shape[0], pad(k), dtype=torch.bfloat16)
            st[:, :k] = x
        else:
            st = torch.zeros(k, pad(x.shape[0]), dtype=torch.bfloat16)
            st[:, :x.shape[0]] = x.t()
        return st.to(dev)

    a_major, b_major = {"k_k": (0, 0), "k_mn": (0, 1), "mn_mn": (1, 1)}[layout]
    a_st, b_st = store(A, a_major), store(B, b_major)
    D = torch.full((m, pad(n)), 7.0, dtype=torch.bfloat16, device=dev)
    args = _lib.GemmArgs(a=a_st.data_ptr(), b=b_st.data_ptr(), d=D.data_ptr(), lse_part=None, lda=a_st.stride(0),
                         ldb=b_st.stride(0), ldd=D.stride(0), m=m, n=n, k=k, a_major=a_major, b_major=b_major, d_f32=0,
                         accumulate=0, cta_group=cta_group, raster_m_fast=raster, k_splits=1)
    for _ in range(2):   # twice: the second launch takes the next work counter of the ring
        _lib.check(lib.clear_gemm_bf16(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "clear_gemm_bf16")
    torch.cuda.synchronize()
    want = A.float() @ B.float().t()
    got = D[:, :n].float().cpu()
    assert (got - want).abs().max().item() <= 2 ** -7 * want.abs().max().item()
    assert (D[:, n:] == 7.0).all(), "columns beyond N must not be written"


@pytest.mark.parametrize("k", [64, 576, 2048])
@pytest.mark.parametrize("layout", ["k_k", "mn_mn"])
def test_gemm_scheduler_ring_wraps_many_times(k, layout):
    """The claimed-work ring (8 slots of decoded items handed over by st.async, two TMA threads, fence-free hand-backs) under
    boundary-dominated work: 2048 output tiles of 1 / 9 / 32 K blocks (about 27 items per CTA pair: the ring wraps three times,
    items shorter and longer than the hand-on distance), five launches back to back on one stream (the work counters re-arm
    themselves), every launch checked against cuBLAS in fp32."""
    import ctypes as C

    from multimodal_unlearning_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    m, n = 8192, 16384
    g = torch.Generator(device=dev).manual_seed(100 + k)
    A = torch.randn(m, k, generator=g, device=dev).to(torch.bfloat16)
    B = torch.randn(n, k, generator=g, device=dev).to(torch.bfloat16)
    major = 0 if layout == "k_k" else 1
    a_st = A.contiguous() if major == 0 else A.t().contiguous()
    b_st = B.contiguous() if major == 0 else B.t().contiguous()
    want = A.float() @ B.float().t()
    scale = want.abs().max().item()
    outs = [torch.full((m, n), 3.0, dtype=torch.bfloat16, device=dev) for _ in range(5)]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for D in outs:
        args = _lib.GemmArgs(a=a_st.data_ptr(), b=b_st.data_ptr(), d=D.data_ptr(), lse_part=None, lda=a_st.stride(0),
                             ldb=b_st.stride(0), ldd=D.stride(0), m=m, n=n, k=k, a_major=major, b_major=major, d_f32=0,
                             accumulate=0, cta_group=0, raster_m_fast=-1, k_splits=1)
        _lib.check(lib.clear_gemm_bf16(C.byref(args), stream), "clear_gemm_bf16")
    torch.cuda.synchronize()
    for D in outs:
        assert (D.float() - want).abs().max().item() <= 2 ** -7 * scale
        assert torch.equal(D, outs[0]), "every launch computes the same bits"


@pytest.mark.parametrize("reduction", ["reference", "seq_mean", "sum"])
@pytest.mark.parametrize("valid_frac", [0.05, 0.5, 1.0, 0.0])
def test_ce_row_compaction_is_exact(reduction, valid_frac):
    """compact_ignored: the head evaluated on the non-ignored rows only gives the loss, dH (zeros on ignored rows) and dW
    of the evaluation over every row — compared with the uncompacted kernels and, for the loss, with the oracle."""
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    b, t, h, v = 4, 96, 128, 1000
    n = b * t
    hid, w, _, _, y = _inputs(n, h, v, 91, dev)
    g = torch.Generator().manual_seed(5)
    keep = torch.rand(n, generator=g) < valid_frac
    y = torch.where(keep, y, torch.full_like(y, -100))
    rw = torch.rand(n, generator=g) if reduction == "seq_mean" else None
    kw = dict(labels=y.to(dev), coef=-0.7, reduction=reduction, n_seqs=b if reduction == "seq_mean" else None,
              row_weight=rw.to(dev) if rw is not None else None)
    full = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, **kw)
    comp = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, compact_ignored=True, **kw)
    torch.cuda.synchronize()
    if valid_frac == 0.0 and reduction == "reference":
        assert torch.isnan(comp.loss).all() and torch.isnan(full.loss).all()      # mean over no rows, as HF
        return
    assert abs(comp.loss.item() - full.loss.item()) <= 1e-5 * max(abs(full.loss.item()), 1e-6)
    assert torch.allclose(comp.loss_rows, full.loss_rows, rtol=1e-6, atol=0)
    ignored = (~keep).to(dev)
    assert (comp.dhidden[ignored] == 0).all() and (full.dhidden[ignored] == 0).all()
    assert torch.equal(comp.dhidden, full.dhidden)        # row-wise results do not depend on which rows share a launch
    if keep.any():
        rel = ((comp.dweight.float() - full.dweight.float()).norm() / full.dweight.float().norm()).item()
        assert rel <= 1e-2, rel                             # same products, different K blocking of the bf16 inputs
    if reduction == "seq_mean":
        assert torch.allclose(comp.seq_loss, full.seq_loss, rtol=1e-5, atol=1e-6)
    # forward only
    f_only = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, compact_ignored=True, need_dhidden=False,
                                    need_dweight=False, **kw)
    assert f_only.dhidden is None and f_only.dweight is None
    assert abs(f_only.loss.item() - full.loss.item()) <= 1e-5 * max(abs(full.loss.item()), 1e-6)


@pytest.mark.parametrize("kind", KINDS)
def test_one_call_entry_equals_the_staged_entry_points(kind):
    """clear_head_loss_fwd_bwd (one C call, workspace carved in the library) and the same stages driven one by one from
    Python (what bench.py instruments with per-launch events) are the same computation, bit for bit."""
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h, v = 300, 128, 1000
    hid, w, rh, rw, y = _inputs(n, h, v, 33 + kind, dev)
    y[::5] = -100
    kw = dict(labels=y.to(dev), ref_hidden=rh.to(dev), ref_weight=rw.to(dev), beta=0.7 if kind == HO.KIND_NPO else 1.0, coef=-0.7,
              block_rows=128)
    one = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), kind, **kw)
    with head.StageTimer(staged=True) as st:
        staged = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), kind, **kw)
    summ = st.summary()
    assert any(k.startswith("gemm") for k in summ) and "rowloss" in summ
    # default timer: the one-call entry itself, events recorded by the library between its stages (3 blocks of 128 rows)
    with head.StageTimer() as st2:
        timed = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), kind, **kw)
    summ2 = st2.summary()
    assert summ2["rowloss"][0] == 3 and summ2["gemm_backward"][0] == 3 and all(v[1] > 0 for v in summ2.values())
    assert torch.equal(timed.loss, one.loss) and torch.equal(timed.dweight, one.dweight)
    assert torch.equal(one.loss, staged.loss) and torch.equal(one.loss_rows, staged.loss_rows)
    assert torch.equal(one.dhidden, staged.dhidden) and torch.equal(one.dweight, staged.dweight)
    assert one.launches == staged.launches


def test_ce_compaction_makes_no_host_synchronisation():
    """The valid-row count of compact_ignored never visits the host: with torch's sync debug mode set to "error" any
    .item() / blocking copy inside the call would raise.  Several valid fractions reuse the same launch configuration."""
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h, v = 1024, 256, 4000
    hid, w, _, _, y = _inputs(n, h, v, 17, dev)
    hid, w = hid.to(dev), w.to(dev)
    ys = []
    for frac in (0.02, 0.4, 1.0):
        g = torch.Generator().manual_seed(int(frac * 100))
        keep = torch.rand(n, generator=g) < frac
        ys.append(torch.where(keep, y.clamp(min=0), torch.full_like(y, -100)).to(dev))
    full = [head.head_loss_fwd_bwd(hid, w, HO.KIND_CE, labels=yy) for yy in ys]
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        comp = [head.head_loss_fwd_bwd(hid, w, HO.KIND_CE, labels=yy, compact_ignored=True, block_rows=256) for yy in ys]
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    for f, c, yy in zip(full, comp, ys):
        assert torch.allclose(c.loss, f.loss, rtol=1e-5)
        assert torch.allclose(c.loss_rows, f.loss_rows, rtol=1e-6, atol=0)
        assert (c.dhidden[yy == -100] == 0).all()
        assert _rel(c.dhidden, f.dhidden.float()) <= 1e-2 and _rel(c.dweight, f.dweight.float()) <= 1e-2


def test_row_map_against_the_reference_fixture():
    """`remove_image_tokens` as a kernel input (SURVEY.md §8a A6d / §8b row_index).  tests/golden/row_map.npz was produced by
    the UNMODIFIED reference functions (remove_image_tokens + get_batch_loss, mm/trainer_utils.py:48-81) on a legacy-processor
    batch (T' = L + 575: image first, in the middle, absent): per-sequence losses and d(sum)/d hidden must match with the
    hidden rows selected by `row_index` inside the library (no cropped copy of logits or hidden is made by the caller)."""
    import numpy as np

    from multimodal_unlearning_b200 import head
    from multimodal_unlearning_b200.trainer import flat_row_index, image_row_index
    from tests._golden import GOLDEN
    dev = torch.device("cuda:0")
    z = np.load(GOLDEN / "row_map.npz")
    ids, labels = torch.from_numpy(z["input_ids"]).to(dev), torch.from_numpy(z["labels"]).to(dev)
    hidden = torch.from_numpy(z["hidden_bits"].copy()).view(torch.bfloat16).to(dev)
    w = torch.from_numpy(z["weight_bits"].copy()).view(torch.bfloat16).to(dev)
    b, tp, h = hidden.shape
    rows = flat_row_index(image_row_index(ids, tp, int(z["image_token"])), tp)
    want_seq = torch.from_numpy(z["seq_loss"])
    got = head.sequence_nll(hidden, w, labels, row_index=rows)          # forward only, compacted
    assert (got.double().cpu() - want_seq).abs().max().item() <= LOSS_RTOL * want_seq.abs().max().item()
    for compact in (False, True):
        res = head.head_loss_fwd_bwd(hidden.reshape(b * tp, h), w, HO.KIND_CE, labels=head.shift_labels(labels), reduction="seq_mean",
                                     n_seqs=b, coef=float(b), row_index=rows, compact_ignored=compact)
        assert (res.seq_loss.double().cpu() - want_seq).abs().max().item() <= LOSS_RTOL * want_seq.abs().max().item()
        assert abs(res.loss.item() - want_seq.sum().item()) <= LOSS_RTOL * want_seq.sum().item()
        want_dh = torch.from_numpy(z["dhidden"]).reshape(b * tp, h)
        assert res.dhidden.shape == (b * tp, h)
        _assert_grad(res.dhidden, want_dh, f"dH through the row map (compact={compact})")
        assert (res.dhidden[(want_dh == 0).all(dim=1).to(dev)] == 0).all()      # rows outside the map: exactly zero
    # a teacher-based kind through the same map equals the call on rows gathered by the caller, bit for bit
    rh = (hidden.float() + 0.2 * torch.randn(hidden.shape, device=dev)).to(torch.bfloat16).reshape(b * tp, h)
    a = head.head_loss_fwd_bwd(hidden.reshape(b * tp, h), w, HO.KIND_KL_FWD, ref_hidden=rh, ref_weight=w, row_index=rows)
    sel = rows.long()
    c = head.head_loss_fwd_bwd(hidden.reshape(b * tp, h)[sel].contiguous(), w, HO.KIND_KL_FWD, ref_hidden=rh[sel].contiguous(), ref_weight=w)
    assert torch.equal(a.loss, c.loss) and torch.equal(a.dweight, c.dweight)
    assert torch.equal(a.dhidden[sel], c.dhidden)


@pytest.mark.parametrize("kind,beta", [(HO.KIND_KL_FWD, 1.0), (HO.KIND_KL_REV, 1.0), (HO.KIND_NPO, 1.0), (HO.KIND_NPO, 0.7)])
@pytest.mark.parametrize("v", [1003, 32064])
def test_fp32_teacher_logits(kind, beta, v):
    """ref_logits in fp32 ("bf16-or-fp32"): consumed at fp32 precision, not rounded to bf16."""
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h = 70, 128
    hid, w, rh, rw, y = _inputs(n, h, v, 50 + kind, dev)
    zr = (rh.float() @ rw.float().t())                                   # fp32 teacher logits, NOT bf16-representable
    pad = (-v) % 8
    zr_dev = torch.nn.functional.pad(zr, (0, pad)).to(dev) if pad else zr.to(dev)
    res = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), kind, ref_logits=zr_dev, beta=beta, coef=0.9, block_rows=32)
    hh = hid.float().requires_grad_(True)
    ww = w.float().requires_grad_(True)
    zs = HO.project(hh, ww, torch.bfloat16)
    if kind == HO.KIND_KL_FWD:
        want = HO.kl_batchmean(HO.logits2probs(zs), HO.logits2probs(zr))
    elif kind == HO.KIND_KL_REV:
        want = HO.kl_batchmean(HO.logits2probs(zr), HO.logits2probs(zs))
    else:
        want = HO.npo_mm(zs, zr, beta)
    (0.9 * want).backward()
    assert abs(res.loss.item() - 0.9 * want.item()) <= LOSS_RTOL * abs(want.item())
    _assert_grad(res.dhidden, hh.grad, "dH")
    _assert_grad(res.dweight, ww.grad, "dW")
    # and it is NOT what the bf16-rounded teacher gives (the fp32 path is really taken) at a loss-level resolution
    res16 = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), kind, ref_logits=zr_dev.to(torch.bfloat16), beta=beta, coef=0.9)
    assert res16.loss.item() != res.loss.item()


def test_out_of_range_label_raises_the_status_flag():
    """torch's CrossEntropyLoss asserts on a label outside [0, V); a kernel cannot raise, so the row's loss term is NaN
    (the loss is poisoned) and bit 0 of the caller's status word is set (ADVICE r1 / VERDICT weak #14)."""
    from multimodal_unlearning_b200 import head
    dev = torch.device("cuda:0")
    n, h, v = 64, 128, 1000
    hid, w, _, _, y = _inputs(n, h, v, 77, dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    ok = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, labels=y.to(dev), status_flags=flags)
    assert int(flags.item()) == 0 and torch.isfinite(ok.loss).all()
    y[3] = v + 5
    for compact in (False, True):
        flags.zero_()
        bad = head.head_loss_fwd_bwd(hid.to(dev), w.to(dev), HO.KIND_CE, labels=y.to(dev), status_flags=flags, compact_ignored=compact)
        assert int(flags.item()) & 1 and torch.isnan(bad.loss).all()
        assert torch.isfinite(bad.dhidden.float()).all() and (bad.dhidden[3] == 0).all()
