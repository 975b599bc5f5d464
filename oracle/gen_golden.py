"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY; runs in the build container (the reference is not present on the GPU box).
    python -m oracle.gen_golden            # writes every fixture (about a minute)

Head fixtures (one file per loss type and trainer):
    reference `compute_loss` + backward on the tiny LLaVA / Llama of BASELINE.json configs[0]
    (2-layer Llama, hidden 128, vocab 32000; fp32; student = seeded init, teacher = student + noise),
    with forward pre-hooks on lm_head capturing what reaches the head.  Stored: per student forward
    the hidden states [B, T, 128] and labels, per teacher forward the hidden states, the loss, dH per
    student forward, and probes of dW (selected rows, column sums, squared norm).  lm_head weights are
    regenerated from seeds (oracle.ref_harness.head_weight); their checksums are stored.
Row-map fixture (row_map.npz): the reference's `remove_image_tokens` + `get_batch_loss` on a synthetic legacy-processor
    batch (T' = L + 575), see gen_row_map.
SalUn fixture:
    reference `save_gradient_ratio` on a 2-layer Llama (hidden 32, vocab 512): per-batch gradients (bf16
    bit patterns), and the ten masks, from the unmodified code and from the same code with
    `stable=True` added to its torch.sort call.
"""
from __future__ import annotations

import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

from . import head_oracle as HO
from . import ref_harness as RH

GOLDEN = Path(__file__).resolve().parents[1] / "tests" / "golden"
TEACHER_HEAD_SEED = 2
N_PROBE = 96

MM_LOSSES = list(HO.MM_CALL_ORDER)
NLP_LOSSES = list(HO.NLP_CALL_ORDER)


def teacher_head_weight():
    g = torch.Generator().manual_seed(TEACHER_HEAD_SEED)
    return RH.head_weight(1) + 0.01 * torch.randn(RH.VOCAB, RH.HIDDEN, generator=g)


def probe_rows(labels: torch.Tensor, seed: int) -> np.ndarray:
    ids = torch.unique(labels[labels >= 0])[: N_PROBE // 2]
    g = torch.Generator().manual_seed(seed)
    rnd = torch.randint(0, RH.VOCAB, (N_PROBE - ids.numel(),), generator=g)
    return torch.unique(torch.cat([ids, rnd])).numpy()


def pack(res: dict, labels: list, names_s: list, names_t: list, rows: np.ndarray, beta: float) -> dict:
    out = {"loss": res["loss"].double().numpy(), "beta": np.float64(beta),
           "student_splits": np.array(names_s), "teacher_splits": np.array(names_t), "probe_rows": rows}
    for i, n in enumerate(names_s):
        out[f"hidden_{i}"] = res["hidden"][i].float().numpy()
        dh = res["dhidden"][i]  # None: the forward's result was discarded by the reference (dpo's forget pass)
        out[f"dhidden_{i}"] = (dh if dh is not None else torch.zeros_like(res["hidden"][i])).float().numpy()
        out[f"labels_{i}"] = labels[i].numpy()
    for i, n in enumerate(names_t):
        out[f"ref_hidden_{i}"] = res["ref_hidden"][i].float().numpy()
    dw = res["dweight"].double()
    out["dw_rows"] = dw[rows].float().numpy()
    out["dw_colsum"] = dw.sum(0).numpy()
    out["dw_sqnorm"] = (dw * dw).sum().numpy()
    out["w_checksum"] = RH.head_weight(1).double().abs().sum().numpy()
    out["wref_checksum"] = teacher_head_weight().double().abs().sum().numpy()
    return out


def gen_mm(length_default: int = 32):
    student = RH.tiny_llava(0)
    teacher = RH.perturbed_copy(student, 5)
    with torch.no_grad():
        teacher.lm_head.weight.copy_(teacher_head_weight())
    for lt in MM_LOSSES:
        length = 256 if lt == "grad_diff_forget_ce_retain_ce" else length_default  # configs[0]: bs 2, seq 256
        beta = 1.0 if lt in ("grad_diff_forget_ce_retain_ce", "grad_ascent") else 0.7
        inputs = {"retain": RH.mm_split(2, length, 11), "forget": RH.mm_split(2, length, 12),
                  "idk": RH.mm_split(2, length, 13)}
        if lt == "sku":
            inputs["random"] = RH.mm_split(2, length, 14, k_random=3)
        names_s, names_t = HO.MM_CALL_ORDER[lt]
        res = RH.run_mm(lt, student, teacher, {k: dict(v) for k, v in inputs.items()}, beta=beta)
        assert len(res["hidden"]) == len(names_s), (lt, len(res["hidden"]), names_s)
        assert len(res["ref_hidden"]) == len(names_t), (lt, len(res["ref_hidden"]), names_t)

        def lab(name):
            if name.startswith("random"):
                return inputs["random"]["labels"][:, int(name[6:])]
            return inputs[name]["labels"]

        labels = [lab(n) for n in names_s]
        out = pack(res, labels, names_s, names_t, probe_rows(labels[0], 3), beta)
        np.savez(GOLDEN / f"head_mm_{lt}.npz", **out)
        print(f"mm  {lt:40s} loss {float(res['loss']):+.6f}")


def gen_mm_dpo_true(length: int = 32):
    """head_mm_dpo_true.npz: value and gradients of the DPO expression the reference computes and overwrites."""
    student = RH.tiny_llava(0)
    teacher = RH.perturbed_copy(student, 5)
    with torch.no_grad():
        teacher.lm_head.weight.copy_(teacher_head_weight())
    inputs = {"retain": RH.mm_split(2, length, 11), "forget": RH.mm_split(2, length, 12), "idk": RH.mm_split(2, length, 13)}
    names_s, names_t = HO.MM_CALL_ORDER["dpo"]
    res = RH.run_mm_dpo_true(student, teacher, {k: dict(v) for k, v in inputs.items()})
    assert len(res["hidden"]) == len(names_s) and len(res["ref_hidden"]) == len(names_t)
    labels = [inputs[n]["labels"] for n in names_s]
    out = pack(res, labels, names_s, names_t, probe_rows(labels[0], 3), 0.1)
    np.savez(GOLDEN / "head_mm_dpo_true.npz", **out)
    print(f"mm  {'dpo_true':40s} loss {float(res['loss']):+.6f}")


def gen_nlp(length: int = 32):
    student = RH.tiny_llama(0)
    teacher = RH.perturbed_copy(student, 5)
    with torch.no_grad():
        teacher.lm_head.weight.copy_(teacher_head_weight())
    for lt in NLP_LOSSES:
        beta = 0.7 if lt not in ("npo",) else 1.0
        triples = {"forget": RH.nlp_triple(2, length, 21), "retain": RH.nlp_triple(2, length, 22),
                   "idk": RH.nlp_triple(2, length, 23)}
        inputs = [triples["forget"], triples["retain"], triples["idk"]]
        names_s, names_t = HO.NLP_CALL_ORDER[lt]
        res = RH.run_nlp(lt, student, teacher, inputs, beta=beta)
        assert len(res["hidden"]) == len(names_s), (lt, len(res["hidden"]), names_s)

        def lab(name):
            if name == "forget+retain":
                return torch.cat([triples["forget"][1], triples["retain"][1]], 0)
            return triples[name][1]

        labels = [lab(n) for n in names_s]
        out = pack(res, labels, names_s, names_t, probe_rows(labels[0], 3), beta)
        np.savez(GOLDEN / f"head_nlp_{lt}.npz", **out)
        print(f"nlp {lt:40s} loss {float(res['loss']):+.6f}")


def gen_salun():
    from transformers import LlamaConfig, LlamaForCausalLM

    _, _, _, SALUN = RH.import_reference()
    torch.manual_seed(3)
    vocab = 512
    model = LlamaForCausalLM(LlamaConfig(vocab_size=vocab, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                         num_attention_heads=2, num_key_value_heads=2)).to(torch.bfloat16)
    g = torch.Generator().manual_seed(4)
    batches = []
    for _ in range(3):
        ids = torch.randint(3, vocab, (4, 24), generator=g)
        lab = ids.clone()
        lab[:, :8] = -100
        batches.append((ids, lab, torch.ones_like(ids), torch.arange(4)))

    # per-batch gradients, exactly as save_gradient_ratio obtains them (:36-44)
    names = [n for n, _ in model.named_parameters()]
    per_batch = []
    model.eval()
    for ids, lab, att, _ in batches:
        model.zero_grad(set_to_none=True)
        model(input_ids=ids, attention_mask=att, labels=lab).loss.backward()
        per_batch.append({n: p.grad.detach().clone() for n, p in model.named_parameters()})

    def run_reference(stable: bool):
        real_sort = torch.sort
        if stable:
            torch.sort = lambda *a, **kw: real_sort(*a, stable=True, **kw)
        try:
            with tempfile.TemporaryDirectory() as d:
                cfg = types.SimpleNamespace(unlearn_lr=1e-5, weight_decay=0.0, save_dir=d)
                SALUN.save_gradient_ratio(batches, model, None, cfg)
                return {th: torch.load(f"{d}/with_{th}.pt") for th in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0)}
        finally:
            torch.sort = real_sort

    ref_unstable = run_reference(False)
    ref_stable = run_reference(True)
    out = {"names": np.array(names), "shapes": np.array([";".join(map(str, per_batch[0][n].shape)) for n in names])}
    for b, grads in enumerate(per_batch):
        out[f"grad_{b}"] = np.concatenate([grads[n].reshape(-1).view(torch.int16).numpy() for n in names])
    for th in ref_stable:
        flat_s = np.concatenate([ref_stable[th][n].reshape(-1).numpy() for n in names])
        flat_u = np.concatenate([ref_unstable[th][n].reshape(-1).numpy() for n in names])
        out[f"mask_stable_{th}"] = np.packbits(flat_s.astype(np.uint8))
        out[f"mask_unstable_{th}"] = np.packbits(flat_u.astype(np.uint8))
    np.savez_compressed(GOLDEN / "salun_tiny.npz", **out)
    n_total = sum(per_batch[0][n].numel() for n in names)
    print(f"salun: {len(names)} tensors, {n_total} elements, 3 batches, 10 thresholds")


REG_L1, REG_L0 = 1e-4, 1e-6
REG_PROBES = ("model.language_model.norm.weight", "model.norm.weight", "lm_head.weight")


def gen_regulariser():
    """L1 / L0 weight regulariser (mm/trainer.py:451-481, nlp/dataloader.py:461-483) through the reference's own
    compute_loss: the same step with and without the lambdas, for l_norm_from = zero / init (mm) and the nlp variant
    (always relative to the oracle model).  Stored: both losses and, for a few parameters, grad(with) - grad(without)."""
    def grads(model):
        return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    out = {"l1_lambda": np.float64(REG_L1), "l0_lambda": np.float64(REG_L0)}
    student = RH.tiny_llava(0)
    teacher = RH.perturbed_copy(student, 5)
    inputs = {"retain": RH.mm_split(2, 32, 11), "forget": RH.mm_split(2, 32, 12), "idk": RH.mm_split(2, 32, 13)}
    plain = RH.run_mm("grad_ascent", student, teacher, {k: dict(v) for k, v in inputs.items()})
    g0 = grads(student)
    for origin in ("zero", "init"):
        res = RH.run_mm("grad_ascent", student, teacher, {k: dict(v) for k, v in inputs.items()}, l1_lambda=REG_L1,
                        l0_lambda=REG_L0, l_norm_from=origin)
        g1 = grads(student)
        out[f"mm_{origin}_loss_plain"] = plain["loss"].double().numpy()
        out[f"mm_{origin}_loss_reg"] = res["loss"].double().numpy()
        for n in g1:
            if n in REG_PROBES:
                d = (g1[n] - g0[n]).float()
                out[f"mm_{origin}_dgrad::{n}"] = (d[:64] if d.dim() == 2 else d).numpy()
    student = RH.tiny_llama(0)
    teacher = RH.perturbed_copy(student, 5)
    triples = [RH.nlp_triple(2, 32, 21), RH.nlp_triple(2, 32, 22), RH.nlp_triple(2, 32, 23)]
    plain = RH.run_nlp("grad_ascent", student, teacher, triples)
    g0 = grads(student)
    res = RH.run_nlp("grad_ascent", student, teacher, triples, l1_lambda=REG_L1, l0_lambda=REG_L0)
    g1 = grads(student)
    out["nlp_loss_plain"] = plain["loss"].double().numpy()
    out["nlp_loss_reg"] = res["loss"].double().numpy()
    for n in g1:
        if n in REG_PROBES:
            d = (g1[n] - g0[n]).float()
            out[f"nlp_dgrad::{n}"] = (d[:64] if d.dim() == 2 else d).numpy()
    np.savez_compressed(GOLDEN / "regulariser.npz", **out)
    print("regulariser:", {k: float(v) for k, v in out.items() if "loss" in k})


def gen_row_map():
    """Legacy-processor case of `dpo` (mm/trainer.py:373-387): the model expands ONE image placeholder per sample into 576
    patch rows (T' = L + 575), `remove_image_tokens` (mm/trainer_utils.py:48-70) crops the logits back to L rows and
    `get_batch_loss` (:73-81) scores them.  Transformers 5.x pre-expands the placeholders, so the tiny-LLaVA fixtures never
    take this branch; here the reference's two functions run, unmodified, on logits formed from stored hidden states:
    inputs (hidden [B, T', H] bf16 bits, W [V, H] bf16 bits, input_ids, labels) and their outputs (the row of `logits` each
    output row came from — recovered by tagging, the per-sequence losses, and d(sum of losses)/d hidden)."""
    _, MMU, _, _ = RH.import_reference()
    b, l, h, v, img_tok, n_patch = 4, 48, 64, 1000, 999, 576
    tp = l + n_patch - 1
    g = torch.Generator().manual_seed(11)
    ids = torch.randint(0, 900, (b, l), generator=g)
    starts = [5, 0, None, 30]                      # image first-occurrence index; sample 2 has no image (left padding)
    for i, s0 in enumerate(starts):
        if s0 is not None:
            ids[i, s0] = img_tok
    labels = torch.randint(0, v, (b, l), generator=g)
    labels[:, :34] = -100                          # question / image part ignored
    labels[1, 40:] = -100
    hidden = torch.randn(b, tp, h, generator=g).to(torch.bfloat16)
    w = (0.1 * torch.randn(v, h, generator=g)).to(torch.bfloat16)
    hid = hidden.float().requires_grad_(True)
    exact = torch.nn.functional.linear(hid, w.float())
    logits = exact + (exact.to(torch.bfloat16).float() - exact).detach()   # a bf16 lm_head's output, straight-through
    cropped = MMU.remove_image_tokens(ids, logits, img_tok)
    loss = MMU.get_batch_loss(cropped, labels)
    loss.sum().backward()
    # which row of `logits` each cropped row is: crop a tensor of row numbers with the same function
    tag = torch.arange(tp, dtype=torch.float32)[None, :, None].expand(b, tp, 1).contiguous()
    rows = MMU.remove_image_tokens(ids, tag, img_tok)[..., 0].long()
    np.savez_compressed(GOLDEN / "row_map.npz", input_ids=ids.numpy(), labels=labels.numpy(),
                        hidden_bits=hidden.view(torch.int16).numpy(), weight_bits=w.view(torch.int16).numpy(),
                        image_token=np.int64(img_tok), rows=rows.numpy(), seq_loss=loss.detach().double().numpy(),
                        dhidden=hid.grad.float().numpy())
    print("row_map: seq_loss", loss.detach().tolist())


def gen_legacy_dpo():
    """The UNMODIFIED `MMTrainerForgetting.compute_loss` with loss_type "dpo" on the legacy-processor stand-in
    (ref_harness.LegacyExpandLM: T' = L + 575 rows per sample), i.e. through the reference's own remove_image_tokens /
    get_batch_loss calls (mm/trainer.py:364-398): the loss and the gradients it leaves on the model."""
    MM, _, _, _ = RH.import_reference()
    student, teacher = RH.LegacyExpandLM(0), RH.LegacyExpandLM(1)
    inputs = {"retain": RH.legacy_split(4, 40, 1), "idk": RH.legacy_split(4, 40, 2), "forget": RH.legacy_split(4, 40, 3)}
    fake = types.SimpleNamespace(loss_type="dpo", teacher_model=teacher, loss_beta=1.0, l1_lambda=0.0, l0_lambda=0.0,
                                 l_norm_from="zero", loss_args={}, is_deepspeed_enabled=False,
                                 to_device=MM.MMTrainerForgetting.to_device)
    loss = MM.MMTrainerForgetting.compute_loss(fake, student, inputs)
    loss.backward()
    out = {"loss": loss.detach().double().numpy()}
    for n, p in student.named_parameters():
        out[f"grad::{n}"] = p.grad.float().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    np.savez_compressed(GOLDEN / "legacy_dpo.npz", **out)
    print("legacy_dpo: loss", float(loss))


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    which = set(sys.argv[1:]) or {"mm", "nlp", "salun", "reg", "row_map"}
    if "row_map" in which:
        gen_row_map()
        gen_legacy_dpo()
    if "reg" in which:
        gen_regulariser()
    if "mm" in which:
        gen_mm()
    if "mm" in which or "dpo_true" in which:
        gen_mm_dpo_true()
    if "nlp" in which:
        gen_nlp()
    if "salun" in which:
        gen_salun()


if __name__ == "__main__":
    main()
