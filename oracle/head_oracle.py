"""CPU restatement of the CLEAR unlearning loss head (reference: somvy/multimodal_unlearning).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg, never by the product package (multimodal_unlearning_b200/ has no CPU fallback).

Parity status: PINNED.  The reference ships no tests or golden vectors, so the pins are
outputs of the reference itself: oracle/gen_golden.py runs the unmodified
`MMTrainerForgetting.compute_loss` (mm/trainer.py:168-483) and `CustomTrainerForgetting.compute_loss`
(nlp/dataloader.py:144-485) on the tiny LLaVA / Llama of BASELINE.json configs[0] and stores loss, dH and
dW probes under tests/golden/; tests/test_oracle_golden.py checks this file against them.

The reference evaluates the path as eager PyTorch on `[B, T, V]` logits; this file restates the same
arithmetic as functions of (hidden states, lm_head weight, labels, teacher hidden states, teacher
weight) — the inputs of the fused CUDA head — so the two can be compared on identical inputs.
Third-party arithmetic on the path that is not in /root/reference: HF `ForCausalLMLoss`
(transformers==4.45.1 pinned by req.txt:250; 5.5.0 here, transformers/loss/loss_utils.py:28-67), restated in
`hf_causal_ce`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

IGNORE = -100

# enum ClearLossKind (include/clear_b200.h)
KIND_CE, KIND_KL_FWD, KIND_KL_REV, KIND_NPO, KIND_ENTROPY = range(5)

# Order of the student / teacher forwards inside the reference's compute_loss, per loss type
# (mm/trainer.py:179-446).  gen_golden.py uses it to label captured lm_head inputs.
MM_CALL_ORDER = {
    "retain_ft": (["retain"], []),
    "grad_ascent": (["forget"], []),
    "grad_diff_forget_ce_retain_ce": (["forget", "retain"], []),
    "grad_diff_forget_ce_retain_kl": (["forget", "retain"], ["retain"]),
    "grad_diff_forget_kl_retain_ce": (["forget", "retain"], ["forget"]),
    "grad_diff_forget_kl_retain_kl": (["forget", "retain"], ["forget", "retain"]),
    "grad_diff_forget_entropy_retain_ce": (["forget", "retain"], []),
    "grad_diff_forget_entropy_retain_kl": (["forget", "retain"], ["retain"]),
    "scrub": (["forget", "retain"], ["forget", "retain"]),
    "kl": (["forget", "retain"], ["retain"]),
    "llmu": (["forget", "idk", "retain"], ["retain"]),
    "npo": (["forget"], ["forget"]),
    "idk": (["retain", "idk"], []),
    "eco_ft": (["forget", "retain"], []),
    "dpo": (["idk", "forget"], ["idk", "forget"]),
    "sku": (["forget", "random0", "random1", "random2", "retain"], ["retain"]),
}
# nlp/dataloader.py:151-457 (case-sensitive names; `idk` is one forward over cat(forget, retain))
NLP_CALL_ORDER = {
    "retain_ft": (["retain"], []),
    "grad_ascent": (["forget"], []),
    "grad_diff_forget_ce_retain_ce": (["forget", "retain"], []),
    "grad_diff_forget_ce_retain_KL": (["forget", "retain"], ["retain"]),
    "grad_diff_forget_KL_retain_ce": (["forget", "retain"], ["forget"]),
    "grad_diff_forget_KL_retain_KL": (["forget", "retain"], ["forget", "retain"]),
    "grad_diff_forget_entropy_retain_ce": (["forget", "retain"], []),
    "scrub": (["forget", "retain"], ["forget", "retain"]),
    "KL": (["forget", "retain"], ["retain"]),
    "LLMU": (["forget", "idk", "retain"], ["retain"]),
    "npo": (["forget"], ["forget"]),
    "idk": (["forget+retain"], []),
    "eco_ft": (["forget", "retain"], []),
    "dpo": (["idk", "forget"], ["idk", "forget"]),
}


# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------
def project(hidden: torch.Tensor, weight: torch.Tensor, logits_dtype=None) -> torch.Tensor:
    """lm_head: nn.Linear(H, V, bias=False) inside `model(**inputs)` (mm/trainer.py:180 etc.).
    `logits_dtype=torch.bfloat16` reproduces a bf16 checkpoint, whose lm_head emits bf16 logits."""
    z = F.linear(hidden, weight)
    if logits_dtype is not None:
        # straight-through rounding: value rounded as the bf16 model would, gradient of the fp32 product
        z = z + (z.detach().to(logits_dtype).to(z.dtype) - z.detach())
    return z


def shift_labels(labels: torch.Tensor) -> torch.Tensor:
    """HF ForCausalLMLoss label handling (loss_utils.py:55-57): pad one ignore on the right, drop the
    first — row (b, t) is scored against labels[b, t+1]; the last row of each sequence is ignored."""
    return F.pad(labels, (0, 1), value=IGNORE)[..., 1:].contiguous()


def hf_causal_ce(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """`outputs.loss` of the HF causal LM (A2): fp32 upcast, shift, mean over non-ignored tokens."""
    z = logits.float()
    y = shift_labels(labels)
    return F.cross_entropy(z.view(-1, z.size(-1)), y.view(-1), ignore_index=IGNORE, reduction="mean")


def logits2probs(logits: torch.Tensor, log: bool = True) -> torch.Tensor:
    """mm/trainer_utils.py:16-21 / nlp/dataloader.py:49-54."""
    p = F.log_softmax(logits, dim=-1) if log else F.softmax(logits, dim=-1)
    return p.view(-1, logits.shape[-1])


def kl_batchmean(x_logp: torch.Tensor, t_logp: torch.Tensor) -> torch.Tensor:
    """F.kl_div(x, t, reduction='batchmean', log_target=True) = sum exp(t) (t - x) / x.size(0);
    x.size(0) = B*T after logits2probs' view (mm/trainer.py:203-208)."""
    return (t_logp.exp() * (t_logp - x_logp)).sum() / x_logp.size(0)


def get_batch_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """mm/trainer_utils.py:73-81: per-sequence SUM of shifted token CE -> [B]."""
    y = labels[..., 1:].contiguous()
    z = logits[..., :-1, :].contiguous()
    return F.cross_entropy(z.transpose(-1, -2), y, ignore_index=IGNORE, reduction="none").sum(dim=-1)


def npo_mm(z: torch.Tensor, z_ref: torch.Tensor, beta: float) -> torch.Tensor:
    """mm/trainer.py:348-353: (2/beta) * mean over ALL N*V elements of log(1 + (p/(q+1e-8))^beta)."""
    q = logits2probs(z_ref, log=False)
    p = logits2probs(z, log=False)
    ratios = p / (q + 1e-8)
    return (2 / beta) * torch.mean(torch.log(1 + torch.pow(ratios, beta)))


def npo_nlp(z: torch.Tensor, z_ref: torch.Tensor, beta: float) -> torch.Tensor:
    """nlp/dataloader.py:396-401 as written: log(1 + (log(p/q))^beta) — NaN whenever log(p/q) < -1
    (SURVEY.md App. B #3).  Restated for completeness; the product routes nlp `npo` to npo_mm."""
    q = logits2probs(z_ref, log=False)
    p = logits2probs(z, log=False)
    ratios = torch.log(p / q)
    return 2 / beta * torch.mean(torch.log(1 + ratios**beta))


def neg_entropy_sum(z: torch.Tensor) -> torch.Tensor:
    """mm/trainer.py:195-196: sum p log p over every row and vocab entry (raw sum)."""
    p = logits2probs(z, log=False)
    return torch.sum(p * torch.log(p))


# --------------------------------------------------------------------------------------------------
# the trainers' compute_loss as a function of head inputs
# --------------------------------------------------------------------------------------------------
class HeadInputs:
    """What one `model(**split)` call contributes to the head: hidden [B,T,H] and labels [B,T]."""

    def __init__(self, hidden, labels):
        self.hidden, self.labels = hidden, labels


def mm_compute_loss(loss_type: str, student: dict, weight, teacher: dict | None = None, ref_weight=None,
                    beta: float = 1.0, loss_args: dict | None = None, logits_dtype=None) -> torch.Tensor:
    """MMTrainerForgetting.compute_loss (mm/trainer.py:168-449), branch by branch.
    student: split name -> HeadInputs ("random" -> list of K HeadInputs); teacher: split -> hidden."""
    loss_type = loss_type.lower()  # mm/trainer.py:114
    loss_args = loss_args or {"bad_weight": 2.5, "random_weight": 1.0, "normal_weight": 0.5}

    def fwd(split):  # student logits
        return project(student[split].hidden, weight, logits_dtype)

    def ce(split):  # outputs.loss
        return hf_causal_ce(fwd(split), student[split].labels)

    def tfwd(split):  # teacher logits under no_grad
        with torch.no_grad():
            return project(teacher[split], ref_weight, logits_dtype)

    def kl_fwd(split):  # kl_div(student_logp, teacher_logp)
        return kl_batchmean(logits2probs(fwd(split)), logits2probs(tfwd(split)))

    def kl_rev(split):  # kl_div(teacher_logp, student_logp): arguments swapped
        return kl_batchmean(logits2probs(tfwd(split)), logits2probs(fwd(split)))

    if loss_type == "retain_ft":  # :179-181
        return ce("retain")
    if loss_type == "grad_ascent":  # :183-185
        return -1 * ce("forget")
    if loss_type.startswith("grad_diff"):  # :187-229
        if "forget_ce" in loss_type:
            forget_loss = ce("forget")
        elif "forget_entropy" in loss_type:
            forget_loss = neg_entropy_sum(fwd("forget"))
        elif "forget_kl" in loss_type:
            forget_loss = kl_fwd("forget")
        else:
            raise ValueError(f"Invalid loss type on forget {loss_type}")
        if "retain_ce" in loss_type:
            retain_loss = ce("retain")
        elif "retain_kl" in loss_type:
            retain_loss = kl_fwd("retain")
        else:
            raise ValueError(f"Invalid loss type on retain {loss_type}")
        return -1 * beta * forget_loss + retain_loss
    if loss_type == "scrub":  # :231-257
        return -1 * beta * kl_rev("forget") + kl_rev("retain") + ce("retain")
    if loss_type == "kl":  # :259-277
        return -1 * ce("forget") + kl_fwd("retain")
    if loss_type == "llmu":  # :279-300
        return -1 * ce("forget") + kl_rev("retain") + ce("idk")
    if loss_type == "npo":  # :343-353
        return npo_mm(fwd("forget"), tfwd("forget"), beta)
    if loss_type == "idk":  # :355-358
        return ce("retain") + ce("idk")
    if loss_type == "eco_ft":  # :359-362
        return beta * ce("forget") + ce("retain")
    if loss_type == "dpo":  # :364-398 — the value returned is -idk_loss_current.mean() (App. B #4)
        idk_loss_current = -1 * get_batch_loss(fwd("idk"), student["idk"].labels)
        return -idk_loss_current.mean()
    if loss_type == "sku":  # :412-446
        bad_loss = ce("forget")
        random_loss = 0.0
        for r in student["random"]:
            random_loss = random_loss + hf_causal_ce(project(r.hidden, weight, logits_dtype), r.labels)
        normal_loss = -1 * kl_fwd("retain")
        return (loss_args["bad_weight"] * bad_loss + loss_args["random_weight"] * random_loss
                + loss_args["normal_weight"] * normal_loss)
    raise ValueError(f"Invalid loss type {loss_type}")


def dpo_true(student: dict, weight, teacher: dict, ref_weight, beta: float = 0.1, logits_dtype=None) -> torch.Tensor:
    """The sigmoid log-ratio the reference computes and then discards (mm/trainer.py:390-394);
    offered by the product behind an explicit flag (`dpo_true`)."""
    def seq(hidden, w, labels):
        return -1 * get_batch_loss(project(hidden, w, logits_dtype), labels)

    idk_cur = seq(student["idk"].hidden, weight, student["idk"].labels)
    fgt_cur = seq(student["forget"].hidden, weight, student["forget"].labels)
    with torch.no_grad():
        idk_ref = seq(teacher["idk"], ref_weight, student["idk"].labels)
        fgt_ref = seq(teacher["forget"], ref_weight, student["forget"].labels)
    pi_logratios = idk_cur - fgt_cur
    ref_logratios = idk_ref - fgt_ref
    return -F.logsigmoid(beta * (pi_logratios - ref_logratios)).mean()


def nlp_compute_loss(loss_type: str, student: dict, weight, teacher: dict | None = None, ref_weight=None,
                     beta: float = 1.0, logits_dtype=None) -> torch.Tensor:
    """CustomTrainerForgetting.compute_loss (nlp/dataloader.py:144-459). Names are case-sensitive."""
    def fwd(split):
        return project(student[split].hidden, weight, logits_dtype)

    def ce(split):
        return hf_causal_ce(fwd(split), student[split].labels)

    def tfwd(split):
        with torch.no_grad():
            return project(teacher[split], ref_weight, logits_dtype)

    def kl_fwd(split):
        return kl_batchmean(logits2probs(fwd(split)), logits2probs(tfwd(split)))

    def kl_rev(split):
        return kl_batchmean(logits2probs(tfwd(split)), logits2probs(fwd(split)))

    if loss_type == "retain_ft":
        return ce("retain")
    if loss_type == "grad_ascent":
        return -1 * ce("forget")
    if loss_type.startswith("grad_diff"):
        if "forget_ce" in loss_type:
            forget_loss = ce("forget")
        elif "forget_entropy" in loss_type:
            forget_loss = neg_entropy_sum(fwd("forget"))
        elif "forget_KL" in loss_type:
            forget_loss = kl_fwd("forget")
        else:
            raise ValueError(f"Invalid loss type on forget {loss_type}")
        if "retain_ce" in loss_type:
            retain_loss = ce("retain")
        elif "retain_KL" in loss_type:
            retain_loss = kl_fwd("retain")
        else:
            raise ValueError(f"Invalid loss type on retain {loss_type}")
        return -1 * beta * forget_loss + retain_loss
    if loss_type == "scrub":
        return -1 * beta * kl_rev("forget") + kl_rev("retain") + ce("retain")
    if loss_type == "KL":
        return -1 * ce("forget") + kl_fwd("retain")
    if loss_type == "LLMU":
        return -1 * ce("forget") + kl_rev("retain") + ce("idk")
    if loss_type == "npo":
        return npo_nlp(fwd("forget"), tfwd("forget"), beta)
    if loss_type == "idk":  # :403-410 one forward over cat(forget, retain), one mean
        return ce("forget+retain")
    if loss_type == "eco_ft":
        return beta * ce("forget") + ce("retain")
    if loss_type == "dpo":  # :425-457
        idk_loss_current = -1 * get_batch_loss(fwd("idk"), student["idk"].labels)
        return -idk_loss_current.mean()
    raise ValueError(f"Invalid loss type {loss_type}")


def loss_and_grads(fn, student: dict, weight, **kw):
    """Runs `fn` (mm_compute_loss / nlp_compute_loss / dpo_true) with autograd on every student hidden
    tensor and on W.  Returns (loss, {split: dH}, dW)."""
    leaves = {}
    st = {}
    for name, hi in student.items():
        if isinstance(hi, list):
            st[name] = []
            for i, x in enumerate(hi):
                h = x.hidden.detach().clone().requires_grad_(True)
                leaves[f"{name}{i}"] = h
                st[name].append(HeadInputs(h, x.labels))
        else:
            h = hi.hidden.detach().clone().requires_grad_(True)
            leaves[name] = h
            st[name] = HeadInputs(h, hi.labels)
    w = weight.detach().clone().requires_grad_(True)
    loss = fn(student=st, weight=w, **kw)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in leaves.items()}, w.grad


# --------------------------------------------------------------------------------------------------
# row-level restatement with the kernel's interface (ClearRowLossArgs): closed forms of SURVEY.md
# App. A, used to test clear_head_rowloss in isolation.  All rows, fp64-capable.
# --------------------------------------------------------------------------------------------------
def rowloss(kind: int, z: torch.Tensor, z_ref: torch.Tensor | None = None, labels: torch.Tensor | None = None,
            beta: float = 1.0):
    """z, z_ref: [N, V]; labels: [N] already shifted.  Returns (loss_rows [N], dterm [N, V]) with the
    normalisations of include/clear_b200.h (no 1/N, 1/M, 2/(N V) factors)."""
    lp = F.log_softmax(z, dim=-1)
    p = lp.exp()
    if kind == KIND_CE:
        valid = labels != IGNORE
        y = labels.clamp(min=0)
        nll = -lp.gather(1, y[:, None])[:, 0]
        onehot = F.one_hot(y, z.size(-1)).to(z.dtype)
        return torch.where(valid, nll, torch.zeros_like(nll)), (p - onehot) * valid[:, None].to(z.dtype)
    if kind == KIND_ENTROPY:
        e = (p * lp).sum(-1)
        return e, p * (lp - e[:, None])
    lq = F.log_softmax(z_ref, dim=-1)
    q = lq.exp()
    if kind == KIND_KL_FWD:
        return (q * (lq - lp)).sum(-1), p - q
    if kind == KIND_KL_REV:
        k = (p * (lp - lq)).sum(-1)
        return k, p * ((lp - lq) - k[:, None])
    if kind == KIND_NPO:
        r = p / (q + 1e-8)
        t = r.pow(beta)
        s = t / (1 + t)
        return torch.log(1 + t).sum(-1), s - p * s.sum(-1, keepdim=True)
    raise ValueError(kind)
