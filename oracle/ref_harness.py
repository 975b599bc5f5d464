"""Runs the UNMODIFIED reference (somvy/multimodal_unlearning, mounted at /root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) validate the CPU restatement in
`oracle/head_oracle.py` / `oracle/salun_oracle.py` against the reference's own code and (b) generate the
golden fixtures under tests/golden/ (see oracle/gen_golden.py).  It needs /root/reference, which only
exists in the build container — nothing that runs on the GPU box imports it.

Import recipe: the reference imports `deepspeed`, `natsort` and `hydra` at module
level; none is installed here, and none is used on the paths we call, so empty stub modules suffice.
`Trainer.__init__` needs `accelerate` (absent): `compute_loss` is therefore called unbound with a
SimpleNamespace standing in for `self`.
"""
from __future__ import annotations

import sys
import types
from dataclasses import dataclass, field

import torch

REFERENCE_ROOT = "/root/reference"


def import_reference():
    for name in ("deepspeed", "natsort", "hydra"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["natsort"].natsorted = sorted
    sys.modules["hydra"].main = lambda **kw: (lambda f: f)
    for p in (REFERENCE_ROOT, REFERENCE_ROOT + "/nlp", REFERENCE_ROOT + "/methods/salun"):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        import mm.trainer as MM  # MMTrainerForgetting            mm/trainer.py:112
        import mm.trainer_utils as MMU  # logits2probs, get_batch_loss      mm/trainer_utils.py
        import dataloader as NLP  # CustomTrainerForgetting        nlp/dataloader.py:94
        import generate_mask as SALUN  # save_gradient_ratio            methods/salun/generate_mask.py:17
    return MM, MMU, NLP, SALUN


# --------------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: tiny random-init LLaVA (2-layer Llama + tiny CLIP, vocab 32000)
# --------------------------------------------------------------------------------------------------
VOCAB = 32000
HIDDEN = 128
IMAGE_TOKEN = 31999
N_IMAGE_TOKENS = 4  # (28 / 14)^2 patches, CLS dropped by the "default" feature-select strategy


def tiny_llava(seed: int = 0):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig, LlavaForConditionalGeneration

    cfg = LlavaConfig(
        text_config=LlamaConfig(vocab_size=VOCAB, hidden_size=HIDDEN, intermediate_size=256, num_hidden_layers=2,
                                num_attention_heads=4, num_key_value_heads=4),
        vision_config=CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                       num_attention_heads=4, image_size=28, patch_size=14),
        image_token_index=IMAGE_TOKEN,
    )
    torch.manual_seed(seed)
    model = LlavaForConditionalGeneration(cfg)
    # lm_head is re-drawn from a dedicated generator so that fixtures can regenerate W from a seed
    # instead of storing 16 MB (tests check its checksum).
    with torch.no_grad():
        model.lm_head.weight.copy_(head_weight(1))
    return model


def head_weight(seed: int, vocab: int = VOCAB, hidden: int = HIDDEN) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return 0.02 * torch.randn(vocab, hidden, generator=g)


def tiny_llama(seed: int = 0):
    from transformers import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(seed)
    model = LlamaForCausalLM(LlamaConfig(vocab_size=VOCAB, hidden_size=HIDDEN, intermediate_size=256,
                                         num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4))
    with torch.no_grad():
        model.lm_head.weight.copy_(head_weight(1))
    return model


def perturbed_copy(model, seed: int, std: float = 0.01):
    """Teacher = student + noise, so that log-ratios are non-trivial."""
    import copy

    teacher = copy.deepcopy(model)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in teacher.parameters():
            p.add_(std * torch.randn(p.shape, generator=g))
    teacher.eval()
    for p in teacher.parameters():
        p.requires_grad_(False)
    return teacher


def mm_split(batch: int, length: int, seed: int, k_random: int = 0) -> dict:
    """One collated split in the layout of mm/dataset.py:291-354: left padding, one image per sample,
    labels = -100 on padding + image + question tokens."""
    g = torch.Generator().manual_seed(seed)
    lead = (batch, k_random) if k_random else (batch,)
    n = batch * max(k_random, 1)
    ids = torch.randint(5, VOCAB - 10, (n, length), generator=g)
    att = torch.ones(n, length, dtype=torch.long)
    lab = ids.clone()
    for i in range(n):
        pad = int(torch.randint(0, max(2, length // 8), (1,), generator=g))
        ans = int(torch.randint(max(2, length // 8), max(3, length // 3), (1,), generator=g))
        ids[i, :pad] = 0
        att[i, :pad] = 0
        ids[i, pad:pad + N_IMAGE_TOKENS] = IMAGE_TOKEN
        lab[i, : length - ans] = -100
    pix = torch.randn(n, 3, 28, 28, generator=g)
    out = {"input_ids": ids, "attention_mask": att, "labels": lab, "pixel_values": pix}
    return {k: v.reshape(*lead, *v.shape[1:]) for k, v in out.items()}


def nlp_triple(batch: int, length: int, seed: int):
    """[input_ids, labels, attention_mask] as produced by nlp/data_module.py:10-43 (right padding with
    EOS, first pad position labelled, the rest -100; question tokens -100)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(5, VOCAB - 10, (batch, length), generator=g)
    att = torch.ones(batch, length, dtype=torch.long)
    lab = ids.clone()
    for i in range(batch):
        used = int(torch.randint(length // 2, length - 2, (1,), generator=g))
        q = int(torch.randint(2, used // 2, (1,), generator=g))
        ids[i, used:] = 2
        att[i, used:] = 0
        lab[i, used + 1:] = -100
        lab[i, :q] = -100
    return [ids, lab, att]


class Wrap(torch.nn.Module):
    """Stands in for DDP/PEFT: makes `model.device = ...` (mm/trainer.py:169) legal on a bare HF model."""

    def __init__(self, m):
        super().__init__()
        self.module, self.config = m, m.config

    def forward(self, *a, **kw):
        return self.module(*a, **kw)


@dataclass
class HeadCapture:
    """Records every lm_head invocation: its input (the trunk's hidden states) and, after backward,
    the gradient that flowed into it."""

    hidden: list = field(default_factory=list)

    def attach(self, model):
        def pre(_mod, args):
            h = args[0]
            if h.requires_grad:
                h.retain_grad()
            self.hidden.append(h)

        return model.lm_head.register_forward_pre_hook(pre)


def run_mm(loss_type: str, student, teacher, inputs: dict, beta: float = 1.0, loss_args=None, l1_lambda: float = 0.0,
           l0_lambda: float = 0.0, l_norm_from: str = "zero"):
    """Reference MMTrainerForgetting.compute_loss + backward; returns loss, captured hidden / dH per
    student forward (in call order), teacher hidden per teacher forward, and dW."""
    MM, _, _, _ = import_reference()
    cap_s, cap_t = HeadCapture(), HeadCapture()
    hs = cap_s.attach(student)
    ht = cap_t.attach(teacher) if teacher is not None else None
    fake_self = types.SimpleNamespace(
        loss_type=loss_type.lower(), teacher_model=teacher, loss_beta=beta, l1_lambda=l1_lambda, l0_lambda=l0_lambda,
        l_norm_from=l_norm_from, loss_args=loss_args or {"bad_weight": 2.5, "random_weight": 1.0, "normal_weight": 0.5},
        is_deepspeed_enabled=False, to_device=MM.MMTrainerForgetting.to_device)
    student.zero_grad(set_to_none=True)
    student.train()
    try:
        loss = MM.MMTrainerForgetting.compute_loss(fake_self, Wrap(student), inputs)
        loss.backward()
    finally:
        hs.remove()
        if ht is not None:
            ht.remove()
    return {
        "loss": loss.detach(),
        "hidden": [h.detach() for h in cap_s.hidden],
        "dhidden": [h.grad.detach() if h.grad is not None else None for h in cap_s.hidden],
        "ref_hidden": [h.detach() for h in cap_t.hidden],
        "dweight": student.lm_head.weight.grad.detach().clone(),
    }


def run_nlp(loss_type: str, student, oracle_model, inputs: list, beta: float = 1.0, l1_lambda: float = 0.0,
            l0_lambda: float = 0.0):
    _, _, NLP, _ = import_reference()
    cap_s, cap_t = HeadCapture(), HeadCapture()
    hs = cap_s.attach(student)
    ht = cap_t.attach(oracle_model) if oracle_model is not None else None
    fake_self = types.SimpleNamespace(loss_type=loss_type, oracle_model=oracle_model, loss_beta=beta, l1_lambda=l1_lambda,
                                      l0_lambda=l0_lambda, is_deepspeed_enabled=False,
                                      to_device=NLP.CustomTrainerForgetting.to_device)
    student.zero_grad(set_to_none=True)
    student.train()
    try:
        loss = NLP.CustomTrainerForgetting.compute_loss(fake_self, student, inputs)
        loss.backward()
    finally:
        hs.remove()
        if ht is not None:
            ht.remove()
    return {
        "loss": loss.detach(),
        "hidden": [h.detach() for h in cap_s.hidden],
        "dhidden": [h.grad.detach() if h.grad is not None else None for h in cap_s.hidden],
        "ref_hidden": [h.detach() for h in cap_t.hidden],
        "dweight": student.lm_head.weight.grad.detach().clone(),
    }


def run_mm_dpo_true(student, teacher, inputs: dict):
    """The DPO sigmoid log-ratio the reference evaluates and then overwrites (mm/trainer.py:390-394): the unmodified
    compute_loss("dpo") runs with torch.nn.functional.logsigmoid recording its result; the recorded tensor is still part
    of the reference's own autograd graph, so -mean of it and its backward are the reference's value and gradients of
    that expression.  Same return layout as run_mm."""
    import torch.nn.functional as F

    MM, _, _, _ = import_reference()
    cap_s, cap_t = HeadCapture(), HeadCapture()
    hs = cap_s.attach(student)
    ht = cap_t.attach(teacher)
    fake_self = types.SimpleNamespace(
        loss_type="dpo", teacher_model=teacher, loss_beta=1.0, l1_lambda=0.0, l0_lambda=0.0, l_norm_from="zero", loss_args={},
        is_deepspeed_enabled=False, to_device=MM.MMTrainerForgetting.to_device)
    student.zero_grad(set_to_none=True)
    student.train()
    recorded = []
    real = F.logsigmoid

    def spy(x, *a, **kw):
        out = real(x, *a, **kw)
        recorded.append(out)
        return out

    F.logsigmoid = spy
    try:
        MM.MMTrainerForgetting.compute_loss(fake_self, Wrap(student), inputs)
        assert len(recorded) == 1, "the reference calls logsigmoid exactly once in the dpo branch"
        loss = -recorded[0].mean()
        loss.backward()
    finally:
        F.logsigmoid = real
        hs.remove()
        ht.remove()
    return {
        "loss": loss.detach(),
        "hidden": [h.detach() for h in cap_s.hidden],
        "dhidden": [h.grad.detach() if h.grad is not None else None for h in cap_s.hidden],
        "ref_hidden": [h.detach() for h in cap_t.hidden],
        "dweight": student.lm_head.weight.grad.detach().clone(),
    }



# --------------------------------------------------------------------------------------------------
# A LLaVA checkpoint as a LEGACY processor drives it (transformers 4.45.1, the reference's pin): input_ids carry ONE image
# placeholder per sample and the model expands it into 576 patch rows inside forward, so the logits have T' = L + 575 rows
# and `dpo` crops them with remove_image_tokens (mm/trainer.py:373-387).  transformers 5.x pre-expands the placeholders, so
# the tiny-LLaVA fixtures never take that branch; this stand-in does, with the reference's own assumptions (at most one
# image per sample, samples without an image left-padded to the common length).
# --------------------------------------------------------------------------------------------------
LEGACY_VOCAB, LEGACY_HIDDEN, LEGACY_PATCHES, LEGACY_IMAGE_TOKEN = 1000, 64, 576, 999


class LegacyExpandLM(torch.nn.Module):
    def __init__(self, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.config = types.SimpleNamespace(image_token_index=LEGACY_IMAGE_TOKEN, tie_word_embeddings=False)
        self.embed = torch.nn.Embedding(LEGACY_VOCAB, LEGACY_HIDDEN)
        self.patch = torch.nn.Parameter(torch.empty(LEGACY_PATCHES, LEGACY_HIDDEN))
        self.mix = torch.nn.Linear(LEGACY_HIDDEN, LEGACY_HIDDEN)
        self.lm_head = torch.nn.Linear(LEGACY_HIDDEN, LEGACY_VOCAB, bias=False)
        with torch.no_grad():
            self.embed.weight.copy_(torch.randn(LEGACY_VOCAB, LEGACY_HIDDEN, generator=g))
            self.patch.copy_(0.5 * torch.randn(LEGACY_PATCHES, LEGACY_HIDDEN, generator=g))
            self.mix.weight.copy_(torch.randn(LEGACY_HIDDEN, LEGACY_HIDDEN, generator=g) / LEGACY_HIDDEN ** 0.5)
            self.mix.bias.zero_()
            self.lm_head.weight.copy_(0.2 * torch.randn(LEGACY_VOCAB, LEGACY_HIDDEN, generator=g))

    def get_output_embeddings(self):
        return self.lm_head

    def forward(self, input_ids, attention_mask=None, labels=None, pixel_values=None):
        x = self.embed(input_ids)
        rows = []
        for i in range(input_ids.shape[0]):
            pos = (input_ids[i] == LEGACY_IMAGE_TOKEN).nonzero().flatten()
            if pos.numel() == 0:   # no image: left padding up to the common length
                rows.append(torch.cat([x.new_zeros(LEGACY_PATCHES - 1, x.shape[-1]), x[i]]))
            else:
                s0 = int(pos[0])
                rows.append(torch.cat([x[i, :s0], self.patch.to(x.dtype), x[i, s0 + 1:]]))
        h = torch.tanh(self.mix(torch.stack(rows)))
        return types.SimpleNamespace(logits=self.lm_head(h), loss=None)


def legacy_split(batch: int, length: int, seed: int, image_at=(5, 0, None, 30)):
    """{input_ids, attention_mask, labels} with one image placeholder per sample at `image_at[i]` (None: text only)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, LEGACY_IMAGE_TOKEN - 50, (batch, length), generator=g)
    for i in range(batch):
        at = image_at[i % len(image_at)]
        if at is not None:
            ids[i, at] = LEGACY_IMAGE_TOKEN
    labels = torch.randint(0, LEGACY_VOCAB, (batch, length), generator=g)
    labels[:, : length - 8] = -100
    return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": labels}
