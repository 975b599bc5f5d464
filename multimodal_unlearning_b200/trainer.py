"""Drop-in unlearning trainers: the reference's `compute_loss` / `forget_loss` surface over the fused head.

    MMTrainerForgetting      <- mm/trainer.py:112-483        (dict-of-splits inputs, lower-cased loss names)
    CustomTrainerForgetting  <- nlp/dataloader.py:94-485     (tuple inputs, case-sensitive loss names)

Same constructor kwargs (`forget_loss, teacher_model | oracle_model, loss_beta, l1_lambda, l0_lambda,
l_norm_from, loss_args`), same `compute_loss(model, inputs, return_outputs=False)` contract, same loss
values.  What changes is underneath: instead of `model(**inputs)`
producing `[B, T, V]` logits that a chain of eager softmax / kl_div / CE ops then consumes, the model's
output embedding is swapped for a capture module for the duration of the call, the trunk runs as
before, and the captured hidden states go to `head.fused_head_loss`, which returns the loss with dH
(back into the trunk's autograd graph) and d lm_head.weight attached.

Deliberate differences from the reference:
  * `model.device = ...` (mm/trainer.py:169) is not assigned (#7);
  * `return_outputs=True` returns `None` outputs instead of raising NameError (#8);
  * nlp `npo` evaluates the mm formula: the nlp one is NaN by construction (#3);
  * shipped `dpo` returns the IDK per-sequence CE like the reference but skips the three forwards whose
    results the reference discards; `dpo_true=True` evaluates the sigmoid log-ratio instead (#4);
  * `rmu` has no vocabulary head and is not provided.

Work the reference does and this does not, with identical results: CE-type terms run the head on the rows whose shifted
label is not -100 only (`compact_ignored`; ignored rows contribute nothing to the loss or to any gradient).
Optional extensions (constructor kwargs, all off by default): `block_rows`, `dpo_true`, `teacher_cache_gb` (replay of the
frozen teacher's hidden states, teacher_cache.py); `enable_fused_head_allreduce(model)` hands the data-parallel mean of
d lm_head.weight to the all-reduce fused into the dW contraction.
"""
from __future__ import annotations

import contextlib

import torch
from torch import nn

from . import head, norms
from .teacher_cache import TeacherCache
from .head import KIND_CE, KIND_ENTROPY, KIND_KL_FWD, KIND_KL_REV, KIND_NPO

try:  # transformers.Trainer needs `accelerate` to be *constructed*, not to be subclassed
    from transformers import Trainer as _TrainerBase
except Exception:  # pragma: no cover - transformers absent
    _TrainerBase = object

LOSSES_WITH_TEACHER = ("DPO", "SCRUB", "RMU", "LLMU", "SKU", "NPO")  # mm/trainer_utils.py:5
DPO_BETA = 0.1  # hard-coded at mm/trainer.py:393, nlp/dataloader.py:452 (not loss_beta)


def loss_needs_teacher(loss_type: str) -> bool:
    """mm/trainer_utils.py:8-9, made case-insensitive (App. B #6: the reference lower-cases the name
    before testing for "KL" and therefore never matches)."""
    return "KL" in loss_type.upper() or loss_type.upper() in LOSSES_WITH_TEACHER


def has_lora_adapter(model) -> bool:
    return any("lora" in name for name, _ in model.named_parameters())


# --------------------------------------------------------------------------------------------------
# getting the trunk's hidden states without running the stock head
# --------------------------------------------------------------------------------------------------
class _HeadCapture(nn.Module):
    """Stands in for lm_head during one forward: keeps its input, emits a 1-wide placeholder."""

    def __init__(self, real: nn.Module):
        super().__init__()
        self.__dict__["real"] = real  # not registered: must not show up in parameters()/state_dict()
        self.hidden = None

    @property
    def weight(self):
        return self.__dict__["real"].weight

    def forward(self, x):
        self.hidden = x
        # The 1-wide placeholder depends on the weight (times zero): under DistributedDataParallel with
        # find_unused_parameters=True — the reference's own setting, mm/forget.py — the reducer searches the graph of the
        # wrapped forward's OUTPUTS for the parameters that will receive a gradient; lm_head.weight gets its gradient later,
        # through the fused head, and must not be marked unused in the meantime.
        w = self.__dict__["real"].weight
        return x[..., :1] + (w.reshape(-1)[:1].sum() * 0).to(x.dtype)


def _unwrap(model):
    seen = 0
    while hasattr(model, "module") and isinstance(getattr(model, "module"), nn.Module) and seen < 8:
        model = model.module
        seen += 1
    if hasattr(model, "get_base_model"):  # peft
        with contextlib.suppress(Exception):
            model = model.get_base_model()
    return model


def _locate_head(model):
    base = _unwrap(model)
    lm_head = base.get_output_embeddings()
    if lm_head is None or not hasattr(lm_head, "weight"):
        raise RuntimeError("model has no output embedding (lm_head): the fused loss head needs its weight")
    if getattr(lm_head, "bias", None) is not None:
        raise RuntimeError("lm_head with a bias is not supported (LLaVA / Llama heads have none)")
    for parent in base.modules():
        for name, child in parent.named_children():
            if child is lm_head:
                return parent, name, lm_head
    raise RuntimeError("could not locate lm_head inside the model")


def trunk_hidden(model, model_inputs: dict):
    """Runs `model(**model_inputs)` (no labels) with lm_head swapped out.  Returns (hidden [B, T', H] still
    attached to the trunk's graph, lm_head.weight [V, H])."""
    parent, name, lm_head = _locate_head(model)
    cap = _HeadCapture(lm_head)
    setattr(parent, name, cap)
    try:
        kw = {k: v for k, v in model_inputs.items() if k != "labels"}
        model(**kw)
    finally:
        setattr(parent, name, lm_head)
    if cap.hidden is None:
        raise RuntimeError("lm_head was not called by the model's forward")
    return cap.hidden, lm_head.weight


def flat_row_index(idx: torch.Tensor, n_rows: int) -> torch.Tensor:
    """[B, L] per-sample row numbers (image_row_index) -> flat int32 [B * L] rows of hidden.reshape(B * n_rows, H): the
    `row_index` input of the fused head."""
    b = idx.shape[0]
    return (idx + torch.arange(b, device=idx.device)[:, None] * n_rows).reshape(-1).to(torch.int32)


def expand_labels(labels: torch.Tensor, input_ids: torch.Tensor, n_rows: int, image_token_id: int) -> torch.Tensor:
    """Labels [B, L] -> [B, T'] when a legacy processor expanded the image placeholder inside the model (T' > L): every text
    position keeps its label at its new row, the image rows (and the placeholder itself) are ignored — what HF's legacy
    `_merge_input_ids_with_image_features` does to `labels` before its CE when the reference calls `model(**x).loss`."""
    idx = image_row_index(input_ids, n_rows, image_token_id)
    if idx is None:
        return labels
    keep = torch.where(input_ids == image_token_id, torch.full_like(labels, head.IGNORE_INDEX), labels)
    out = torch.full((labels.shape[0], n_rows), head.IGNORE_INDEX, dtype=labels.dtype, device=labels.device)
    return out.scatter(1, idx, keep)


def image_row_index(input_ids: torch.Tensor, n_rows: int, image_token_id: int) -> torch.Tensor | None:
    """`remove_image_tokens` (mm/trainer_utils.py:48-70) as a row map: for each sample the indices of
    the L output rows that line up with the L input positions when the model expanded one image
    placeholder into patch rows (legacy processors: T' > L).  Rows [0, s) and the last L - s rows, s =
    first image-token position; the last L rows when the sample has no image.  None when T' == L.
    No host synchronisation (the reference calls .tolist())."""
    b, l = input_ids.shape
    if n_rows == l:
        return None
    is_img = input_ids == image_token_id
    has = is_img.any(dim=1, keepdim=True)
    s = torch.where(has, is_img.int().argmax(dim=1, keepdim=True), torch.zeros_like(has, dtype=torch.int64))
    pos = torch.arange(l, device=input_ids.device)[None, :]
    return torch.where(pos < s, pos, pos + (n_rows - l))


# --------------------------------------------------------------------------------------------------
# the loss recipes
# --------------------------------------------------------------------------------------------------
class _Recipes:
    """compute_loss body shared by the two trainers.  `fwd(split)` runs the student trunk on a split and
    returns (hidden, weight, labels); `tfwd(split)` the teacher's (hidden, weight) under no_grad."""

    def __init__(self, fwd, tfwd, beta, loss_args, block_rows, kl_tag: str):
        self._fwd, self._tfwd, self.beta, self.loss_args, self.block_rows = fwd, tfwd, beta, loss_args, block_rows
        self.kl_tag = kl_tag  # "kl" (mm, lower-cased names) or "KL" (nlp)
        self._cache, self._tcache = {}, {}
        self.expand = lambda split, labels, n_rows: labels   # mm: expand_labels for legacy processors (T' > L)

    # One trunk forward per (model, split) and step: `scrub` reads both .logits and .loss of the same
    # retain forward (mm/trainer.py:244-257), `dpo` scores the same forwards twice.
    def fwd(self, split):
        if split not in self._cache:
            self._cache[split] = self._fwd(split)
        return self._cache[split]

    def tfwd(self, split):
        if split not in self._tcache:
            self._tcache[split] = self._tfwd(split)
        return self._tcache[split]

    # -- building blocks -------------------------------------------------------------------------
    def ce(self, split, coef=1.0):
        """`model(**split).loss` — HF causal-LM CE, mean over non-ignored shifted labels (A2)."""
        h, w, labels, _ = self.fwd(split)
        if labels.shape[1] != h.shape[1]:   # legacy processor: image placeholder expanded inside the model
            labels = self.expand(split, labels, h.shape[1])
        return head.fused_head_loss(h, w, KIND_CE, labels=head.shift_labels(labels), coef=coef,
                                    block_rows=self.block_rows, compact_ignored=True)

    def _with_teacher(self, split, kind, coef, beta=1.0):
        h, w, _, _ = self.fwd(split)
        rh, rw = self.tfwd(split)
        if rh.shape != h.shape:
            raise RuntimeError(f"teacher and student produced different shapes on '{split}': {rh.shape} vs {h.shape}")
        return head.fused_head_loss(h, w, kind, ref_hidden=rh, ref_weight=rw, beta=beta, coef=coef,
                                    block_rows=self.block_rows)

    def kl_fwd(self, split, coef=1.0):
        """kl_div(student_logp, teacher_logp, batchmean, log_target) over ALL rows / N (A4)."""
        return self._with_teacher(split, KIND_KL_FWD, coef)

    def kl_rev(self, split, coef=1.0):
        """kl_div(teacher_logp, student_logp, ...): arguments swapped, gradient through the target (A5)."""
        return self._with_teacher(split, KIND_KL_REV, coef)

    def entropy(self, split, coef=1.0):
        h, w, _, _ = self.fwd(split)
        return head.fused_head_loss(h, w, KIND_ENTROPY, coef=coef, block_rows=self.block_rows)

    def npo(self, split):
        return self._with_teacher(split, KIND_NPO, 1.0, beta=self.beta)

    def seq_ce(self, split, coef=1.0, row_weight=None):
        """get_batch_loss: per-sequence sums -> (mean over B scaled by coef, [B] sums) (A6c)."""
        h, w, labels, rows = self.fwd(split)
        return head.fused_head_loss(h, w, KIND_CE, labels=head.shift_labels(labels), coef=coef, reduction="seq_mean",
                                    n_seqs=h.shape[0], row_weight=row_weight, block_rows=self.block_rows, compact_ignored=True,
                                    row_index=rows)

    def teacher_seq_logp(self, split, labels, rows=None):
        rh, rw = self.tfwd(split)
        return -head.sequence_nll(rh, rw, labels, block_rows=self.block_rows, row_index=rows)

    # -- dispatch ----------------------------------------------------------------------------------
    def grad_diff(self, t):
        tag = self.kl_tag
        if "forget_ce" in t:
            forget = self.ce("forget", -self.beta)
        elif "forget_entropy" in t:
            forget = self.entropy("forget", -self.beta)
        elif f"forget_{tag}" in t:
            forget = self.kl_fwd("forget", -self.beta)
        else:
            raise ValueError(f"Invalid loss type on forget {t}")
        if "retain_ce" in t:
            retain = self.ce("retain")
        elif f"retain_{tag}" in t:
            retain = self.kl_fwd("retain")
        else:
            raise ValueError(f"Invalid loss type on retain {t}")
        return forget + retain  # -1 * beta * forget_loss + retain_loss (the coefficient rides in the kernel)

    def dpo_true(self):
        """-mean_b logsigmoid(0.1 * ((lp_idk - lp_fgt) - (lp_idk_ref - lp_fgt_ref)))  (mm/trainer.py:390-394).
        Two-phase: forward-only per-sequence log-probs fix the per-sequence weights, then the
        forward+backward runs with those weights on the rows."""
        hi, wi, li, ri = self.fwd("idk")
        hf, wf, lf, rf = self.fwd("forget")
        with torch.no_grad():
            pi = (-head.sequence_nll(hi.detach(), wi.detach(), li, self.block_rows, row_index=ri)
                  + head.sequence_nll(hf.detach(), wf.detach(), lf, self.block_rows, row_index=rf))
            ref = self.teacher_seq_logp("idk", li, ri) - self.teacher_seq_logp("forget", lf, rf)
            x = DPO_BETA * (pi - ref)
            wgt = DPO_BETA * torch.sigmoid(-x)  # -d logsigmoid(x)/dx = -sigmoid(-x): +w on idk CE rows, -w on forget CE rows
            value = -torch.nn.functional.logsigmoid(x).mean()

        def weighted(h, w, labels, rows, coef):
            return head.fused_head_loss(h, w, KIND_CE, labels=head.shift_labels(labels), coef=coef, reduction="seq_mean",
                                        n_seqs=h.shape[0], row_weight=wgt.repeat_interleave(labels.shape[1]),
                                        block_rows=self.block_rows, compact_ignored=True, row_index=rows)[0]

        surrogate = weighted(hi, wi, li, ri, 1.0) + weighted(hf, wf, lf, rf, -1.0)  # gradient of the true loss
        return surrogate + (value - surrogate).detach()


def enable_fused_head_allreduce(model, group=None):
    """Data-parallel runs (one process per GPU, one NVSwitch box): average d loss / d lm_head.weight with the all-reduce
    that is fused into the dW contraction (distributed.FusedGradReducer) instead of DDP's bucketed NCCL all-reduce.
    Call once, on the bare model, BEFORE it is wrapped by DistributedDataParallel (HF Trainer / accelerate wrap it
    inside `train()`): the parameter is put on DDP's ignore list and every head invocation then delivers the mean
    gradient itself.  Returns the reducer (None when there is a single rank)."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    from . import distributed as D

    if not dist.is_initialized() or dist.get_world_size(group) < 2:
        return None
    base = _unwrap(model)
    if getattr(getattr(base, "config", None), "tie_word_embeddings", False):
        raise RuntimeError("lm_head.weight is tied to the input embedding: its gradient has a second contribution that "
                           "DDP must reduce; the fused head all-reduce cannot take the parameter over")
    lm_head = _locate_head(model)[2]
    names = [n for n, p in model.named_parameters() if p is lm_head.weight]
    if not names:
        raise RuntimeError("lm_head.weight is not a parameter of the model")
    DDP._set_params_and_buffers_to_ignore_for_model(model, names)
    # 8 rotating buffers (2.1 GB at V x H = 32064 x 4096): recipes with up to six head invocations per step (sku: 5) keep every
    # reduced gradient in place until backward, none is copied
    reducer = D.FusedGradReducer(tuple(lm_head.weight.shape), lm_head.weight.device, group=group, average=True, n_buffers=8)
    head.set_dweight_reducer(reducer)
    return reducer


def _teacher_cache(self) -> TeacherCache | None:
    """Extension: `teacher_cache_gb=<float>` ctor kwarg (or a `teacher_cache` attribute) enables
    replay of the frozen teacher's hidden states; absent / 0 = the reference's behaviour (teacher forward every step)."""
    cache = getattr(self, "teacher_cache", None)
    if cache is None:
        gb = getattr(self, "teacher_cache_gb", 0) or 0
        if gb > 0:
            cache = TeacherCache(int(gb * (1 << 30)))
            self.teacher_cache = cache
    return cache


def _regulariser(self, model, teacher, loss):
    """L1 / L0 weight regulariser (mm/trainer.py:451-481; nlp/dataloader.py:461-483).  The norm of the
    concatenation is the sum of per-tensor norms, so the 7 B-element torch.cat is not formed."""
    l1 = getattr(self, "l1_lambda", None) or 0
    l0 = getattr(self, "l0_lambda", None) or 0
    if l1 == 0 and l0 == 0:
        return loss
    origin = getattr(self, "l_norm_from", "init")  # the nlp trainer has no l_norm_from: always distance from the oracle
    params = [p for p in model.parameters() if p.requires_grad]
    refs = None
    if origin == "init" and not has_lora_adapter(model):
        if teacher is None:
            raise ValueError("Teacher model is required for L1/L0 regularization during training without LORA!")
        refs = [tp.to(p.device) for p, tp in zip(model.parameters(), teacher.parameters()) if p.requires_grad]
        for p, r in zip(params, refs):
            assert p.shape == r.shape
    elif origin == "zero" and has_lora_adapter(model):
        raise ValueError("L1/L0 regularization is not supported for LORA!")
    return loss + norms.weight_regulariser(params, refs, l1, l0).to(loss.dtype)


class MMTrainerForgetting(_TrainerBase):
    """mm/trainer.py:112-483.  Construct exactly like the reference's class."""

    def __init__(self, *args, **kwargs):
        self.loss_type = kwargs.pop("forget_loss").lower()
        self.teacher_model = kwargs.pop("teacher_model")
        self.loss_beta = kwargs.pop("loss_beta")
        self.l1_lambda = kwargs.pop("l1_lambda")
        self.l0_lambda = kwargs.pop("l0_lambda")
        self.l_norm_from = kwargs.pop("l_norm_from")
        self.loss_args = kwargs.pop("loss_args", {})
        self.block_rows = kwargs.pop("block_rows", None)   # extension: rows staged per pass of the head
        self.dpo_true = kwargs.pop("dpo_true", False)      # extension: App. B #4
        self.teacher_cache_gb = kwargs.pop("teacher_cache_gb", 0)  # extension: replay of the frozen teacher (teacher_cache.py)
        if self.l1_lambda and self.l1_lambda != 0 and self.l_norm_from not in ("zero", "init"):
            raise ValueError(f"Invalid l_norm_from {self.l_norm_from}, should be 'zero' or 'init'")
        super().__init__(*args, **kwargs)
        if self.teacher_model is not None:
            self.teacher_model.eval()
            for p in self.teacher_model.parameters():
                p.requires_grad_(False)

    @staticmethod
    def to_device(device, inputs):
        return {k: v.to(device) for k, v in inputs.items()}

    def compute_loss(self, model, inputs, return_outputs=False, *args, **kwargs):
        device = next(model.parameters()).device
        splits = {k: MMTrainerForgetting.to_device(device, v) for k, v in inputs.items()}
        if "random" in splits:  # [B, K, ...] -> K splits (mm/trainer.py:420-421)
            rnd = splits.pop("random")
            k_rand = rnd["input_ids"].shape[1]
            for i in range(k_rand):
                splits[f"random{i}"] = {k: v.transpose(0, 1)[i] for k, v in rnd.items()}
        else:
            k_rand = 0
        teacher = getattr(self, "teacher_model", None)
        cfg = _unwrap(model).config
        image_token = getattr(cfg, "image_token_index", getattr(cfg, "image_token_id", None))

        # `remove_image_tokens` (mm/trainer_utils.py:48-70; the reference applies it in `dpo` only, mm/trainer.py:373-387): with
        # a legacy processor the model expands the image placeholder into patch rows (T' > L).  No copy is made: the L rows
        # that line up with the input positions are named by a row map that the fused head takes as a kernel input.
        def row_map(hidden, split):
            ids = splits[split]["input_ids"]
            if image_token is None or hidden.shape[1] == ids.shape[1]:
                return None
            return flat_row_index(image_row_index(ids, hidden.shape[1], image_token), hidden.shape[1])

        # KL / NPO / entropy consume every output row, as in the reference; CE terms see the labels the way HF's legacy merge
        # hands them to its CE (expand_labels); `dpo` scores the cropped rows (crop=True).
        def fwd(split, crop=False):
            h, w = trunk_hidden(model, splits[split])
            return h, w, splits[split]["labels"], (row_map(h, split) if crop else None)

        def tfwd(split, crop=False):
            if teacher is None:
                raise RuntimeError(f"loss '{self.loss_type}' needs teacher_model")
            cache = _teacher_cache(self)
            keys = None
            if cache is not None:
                keys = TeacherCache.sample_keys(splits[split], image_token)
                h = cache.lookup(keys)
                if h is not None:
                    return h.to(device), _locate_head(teacher)[2].weight.to(device)
            with torch.no_grad():
                h, w = trunk_hidden(teacher, MMTrainerForgetting.to_device(next(teacher.parameters()).device, splits[split]))
            if cache is not None:
                cache.store(keys, h)
            return h.to(device), w.to(device)

        r = _Recipes(fwd, tfwd, self.loss_beta, getattr(self, "loss_args", None) or {},
                     getattr(self, "block_rows", None), "kl")
        if image_token is not None:
            r.expand = lambda split, labels, n_rows: expand_labels(labels, splits[split]["input_ids"], n_rows, image_token)
        t = self.loss_type
        if t == "retain_ft":
            loss = r.ce("retain")
        elif t == "grad_ascent":
            loss = r.ce("forget", -1.0)
        elif t.startswith("grad_diff"):
            loss = r.grad_diff(t)
        elif t == "scrub":
            loss = r.kl_rev("forget", -self.loss_beta) + r.kl_rev("retain") + r.ce("retain")
        elif t == "kl":
            loss = r.ce("forget", -1.0) + r.kl_fwd("retain")
        elif t == "llmu":
            loss = r.ce("forget", -1.0) + r.ce("idk") + r.kl_rev("retain")
        elif t == "npo":
            loss = r.npo("forget")
        elif t == "idk":
            loss = r.ce("retain") + r.ce("idk")
        elif t == "eco_ft":
            loss = r.ce("forget", self.loss_beta) + r.ce("retain")
        elif t == "dpo":
            r._fwd = lambda s: fwd(s, crop=True)
            r._tfwd = lambda s: tfwd(s, crop=True)
            loss = r.dpo_true() if getattr(self, "dpo_true", False) else r.seq_ce("idk")[0]
        elif t == "sku":
            la = r.loss_args
            loss = r.ce("forget", la["bad_weight"])
            for i in range(k_rand):
                loss = loss + r.ce(f"random{i}", la["random_weight"])
            loss = loss + r.kl_fwd("retain", -la["normal_weight"])
        elif t == "rmu":
            raise NotImplementedError("rmu is an activation-MSE loss without a vocabulary head (mm/trainer.py:302-326); "
                                      "it is outside the fused loss head")
        else:
            raise ValueError(f"Invalid loss type {self.loss_type}")
        loss = _regulariser(self, model, teacher, loss)
        return (loss, None) if return_outputs else loss

    def prediction_step(self, model, inputs, prediction_loss_only: bool, ignore_keys=None):
        input_ids, labels, attention_mask = inputs
        with torch.no_grad():
            outputs = model(input_ids, labels=labels, attention_mask=attention_mask)
        return (outputs.loss, outputs.logits, labels)


class CustomTrainerForgetting(_TrainerBase):
    """nlp/dataloader.py:94-485.  Inputs are [forget, retain, (idk)] triples of (input_ids, labels,
    attention_mask); loss names are case-sensitive (`KL`, `LLMU`, `grad_diff_forget_KL_retain_ce`, ...)."""

    def __init__(self, *args, **kwargs):
        self.loss_type = kwargs.pop("forget_loss")
        self.oracle_model = kwargs.pop("oracle_model")
        self.loss_beta = kwargs.pop("loss_beta")
        self.l1_lambda = kwargs.pop("l1_lambda")
        self.l0_lambda = kwargs.pop("l0_lambda")
        self.block_rows = kwargs.pop("block_rows", None)
        self.dpo_true = kwargs.pop("dpo_true", False)
        self.teacher_cache_gb = kwargs.pop("teacher_cache_gb", 0)
        super().__init__(*args, **kwargs)
        if self.oracle_model is not None:
            self.oracle_model.eval()
            for p in self.oracle_model.parameters():
                p.requires_grad_(False)

    @staticmethod
    def to_device(device, inputs):
        return (inputs[0].to(device), inputs[1].to(device), inputs[2].to(device))

    def compute_loss(self, model, inputs, return_outputs=False, *args, **kwargs):
        device = next(model.parameters()).device
        forget_inputs, retain_inputs, *idk_inputs = inputs
        splits = {"forget": CustomTrainerForgetting.to_device(device, forget_inputs),
                  "retain": CustomTrainerForgetting.to_device(device, retain_inputs)}
        if idk_inputs:
            splits["idk"] = CustomTrainerForgetting.to_device(device, idk_inputs[0])
        teacher = getattr(self, "oracle_model", None)

        def fwd(split):
            if split == "forget+retain" and split not in splits:
                # nlp `idk`: one forward over cat(forget, retain), one mean (nlp/dataloader.py:403-410); built on demand
                splits[split] = tuple(torch.cat((a, b), dim=0) for a, b in zip(splits["forget"], splits["retain"]))
            ids, labels, att = splits[split]
            h, w = trunk_hidden(model, {"input_ids": ids, "attention_mask": att})
            return h, w, labels, None

        def tfwd(split):
            if teacher is None:
                raise RuntimeError(f"loss '{self.loss_type}' needs oracle_model")
            ids, _, att = splits[split]
            tdev = next(teacher.parameters()).device
            cache = _teacher_cache(self)
            keys = None
            if cache is not None:
                keys = TeacherCache.sample_keys({"input_ids": ids, "attention_mask": att}, None)
                h = cache.lookup(keys)
                if h is not None:
                    return h.to(device), _locate_head(teacher)[2].weight.to(device)
            with torch.no_grad():
                h, w = trunk_hidden(teacher, {"input_ids": ids.to(tdev), "attention_mask": att.to(tdev)})
            if cache is not None:
                cache.store(keys, h)
            return h.to(device), w.to(device)

        r = _Recipes(fwd, tfwd, self.loss_beta, {}, getattr(self, "block_rows", None), "KL")
        t = self.loss_type
        if t == "retain_ft":
            loss = r.ce("retain")
        elif t == "grad_ascent":
            loss = r.ce("forget", -1.0)
        elif t.startswith("grad_diff"):
            loss = r.grad_diff(t)
        elif t == "scrub":
            loss = r.kl_rev("forget", -self.loss_beta) + r.kl_rev("retain") + r.ce("retain")
        elif t == "KL":
            loss = r.ce("forget", -1.0) + r.kl_fwd("retain")
        elif t == "LLMU":
            loss = r.ce("forget", -1.0) + r.ce("idk") + r.kl_rev("retain")
        elif t == "npo":
            loss = r.npo("forget")  # the mm formula (App. B #3)
        elif t == "idk":
            loss = r.ce("forget+retain")
        elif t == "eco_ft":
            loss = r.ce("forget", self.loss_beta) + r.ce("retain")
        elif t == "dpo":
            loss = r.dpo_true() if getattr(self, "dpo_true", False) else r.seq_ce("idk")[0]
        elif t == "RMU":
            raise NotImplementedError("RMU is an activation-MSE loss without a vocabulary head; outside the fused loss head")
        else:
            raise ValueError(f"Invalid loss type {self.loss_type}")
        loss = _regulariser(self, model, teacher, loss)
        return (loss, None) if return_outputs else loss

    def prediction_step(self, model, inputs, prediction_loss_only: bool, ignore_keys=None):
        input_ids, labels, attention_mask = inputs
        with torch.no_grad():
            outputs = model(input_ids, labels=labels, attention_mask=attention_mask)
        return (outputs.loss, outputs.logits, labels)
