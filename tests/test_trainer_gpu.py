"""GPU: the drop-in trainers on a real (tiny, random-init) HF LLaVA / Llama in bf16 — BASELINE configs[0]
shapes — against the same recipes evaluated the reference's way (stock `model(**inputs)` logits + eager
torch ops, restated in oracle/head_oracle.py) on the same device and weights."""
import types

import pytest
import torch

from oracle import head_oracle as HO
from oracle import ref_harness as RH

pytestmark = pytest.mark.gpu


def _models(kind):
    dev = torch.device("cuda:0")
    student = RH.tiny_llava(0) if kind == "mm" else RH.tiny_llama(0)   # fp32 on the host, seeded
    teacher = RH.perturbed_copy(student, 5).to(dev).to(torch.bfloat16)
    student = student.to(dev).to(torch.bfloat16)
    student.train()
    return student, teacher, dev


def _stock(model, split, dev, grad=True):
    """What the reference does: model(**inputs) -> .loss / .logits (fp32 upcast of the bf16 logits)."""
    kw = {k: v.to(dev) for k, v in split.items()}
    ctx = torch.enable_grad() if grad else torch.no_grad()
    with ctx:
        out = model(**kw)
    return out


def _reference_mm(loss_type, student, teacher, inputs, beta, dev):
    f = lambda s: _stock(student, inputs[s], dev)
    t = lambda s: _stock(teacher, inputs[s], dev, grad=False).logits.float()
    lp = lambda z: HO.logits2probs(z.float())
    if loss_type == "grad_diff_forget_ce_retain_ce":
        return -beta * f("forget").loss + f("retain").loss
    if loss_type == "npo":
        return HO.npo_mm(f("forget").logits.float(), t("forget"), beta)
    if loss_type == "grad_diff_forget_kl_retain_kl":
        return (-beta * HO.kl_batchmean(lp(f("forget").logits), lp(t("forget")))
                + HO.kl_batchmean(lp(f("retain").logits), lp(t("retain"))))
    if loss_type == "scrub":
        r = f("retain")
        return (-beta * HO.kl_batchmean(lp(t("forget")), lp(f("forget").logits))
                + HO.kl_batchmean(lp(t("retain")), lp(r.logits)) + r.loss)
    if loss_type == "dpo":
        return HO.get_batch_loss(f("idk").logits.float(), inputs["idk"]["labels"].to(dev)).mean()
    if loss_type == "grad_diff_forget_entropy_retain_ce":
        return -beta * HO.neg_entropy_sum(f("forget").logits.float()) + f("retain").loss
    raise KeyError(loss_type)


def _grads(model):
    """lm_head (dW), two trunk weights and the image projector (reached through dH and the image rows)."""
    params = dict(model.named_parameters())

    def pick(*frags):
        cand = [k for k in params if all(f in k for f in frags)]
        assert cand, frags
        return cand[0]

    names = [pick("lm_head.weight"), pick("language_model", "layers.1.self_attn.q_proj.weight"),
             pick("language_model", "layers.0.mlp.down_proj.weight"), pick("multi_modal_projector", "linear_1.weight")]
    return {n: (None if params[n].grad is None else params[n].grad.detach().float().clone()) for n in names}


@pytest.mark.parametrize("loss_type", ["grad_diff_forget_ce_retain_ce", "npo", "grad_diff_forget_kl_retain_kl", "scrub", "dpo",
                                       "grad_diff_forget_entropy_retain_ce"])
def test_mm_trainer_on_tiny_llava(loss_type):
    from multimodal_unlearning_b200 import trainer as T
    student, teacher, dev = _models("mm")
    length = 256 if loss_type == "grad_diff_forget_ce_retain_ce" else 64   # configs[0]: bs 2, seq 256
    beta = 0.7
    inputs = {"retain": RH.mm_split(2, length, 11), "forget": RH.mm_split(2, length, 12), "idk": RH.mm_split(2, length, 13)}
    student.zero_grad(set_to_none=True)
    want = _reference_mm(loss_type, student, teacher, inputs, beta, dev)
    want.backward()
    want_g = _grads(student)
    student.zero_grad(set_to_none=True)
    fake = types.SimpleNamespace(loss_type=loss_type, teacher_model=teacher, loss_beta=beta, l1_lambda=0.0, l0_lambda=0.0,
                                 l_norm_from="zero", loss_args={})
    got = T.MMTrainerForgetting.compute_loss(fake, student, inputs)
    assert got.dim() == 0 and got.device.type == "cuda"
    got.backward()
    got_g = _grads(student)
    assert abs(got.item() - want.item()) <= 2e-3 * max(abs(want.item()), 1e-2), (got.item(), want.item())
    for n, w in want_g.items():
        g = got_g[n]
        assert g is not None and torch.isfinite(g).all(), n
        rel = ((g - w).norm() / w.norm().clamp_min(1e-30)).item()
        # both sides are bf16 end to end (the stock path rounds its dlogits and GEMM outputs too)
        assert rel <= 5e-2, (loss_type, n, rel)
    # the head was restored and no capture module leaked into the model
    assert type(student.lm_head).__name__ == "Linear"


@pytest.mark.parametrize("loss_type", ["grad_ascent", "KL", "idk"])
def test_nlp_trainer_on_tiny_llama(loss_type):
    from multimodal_unlearning_b200 import trainer as T
    student, teacher, dev = _models("nlp")
    beta = 1.0
    tr = [RH.nlp_triple(2, 64, 21), RH.nlp_triple(2, 64, 22), RH.nlp_triple(2, 64, 23)]

    def stock(model, t, grad=True):
        with (torch.enable_grad() if grad else torch.no_grad()):
            return model(t[0].to(dev), labels=t[1].to(dev), attention_mask=t[2].to(dev))

    student.zero_grad(set_to_none=True)
    if loss_type == "grad_ascent":
        want = -stock(student, tr[0]).loss
    elif loss_type == "KL":
        want = -stock(student, tr[0]).loss + HO.kl_batchmean(HO.logits2probs(stock(student, tr[1]).logits.float()),
                                                              HO.logits2probs(stock(teacher, tr[1], False).logits.float()))
    else:
        cat = [torch.cat((a, b), 0) for a, b in zip(tr[0], tr[1])]
        want = stock(student, cat).loss
    want.backward()
    want_w = student.lm_head.weight.grad.float().clone()
    student.zero_grad(set_to_none=True)
    fake = types.SimpleNamespace(loss_type=loss_type, oracle_model=teacher, loss_beta=beta, l1_lambda=0.0, l0_lambda=0.0)
    got = T.CustomTrainerForgetting.compute_loss(fake, student, tr)
    got.backward()
    assert abs(got.item() - want.item()) <= 2e-3 * abs(want.item())
    rel = ((student.lm_head.weight.grad.float() - want_w).norm() / want_w.norm()).item()
    assert rel <= 5e-2, rel


def test_save_gradient_ratio_dropin(tmp_path):
    """multimodal_unlearning_b200.salun.save_gradient_ratio == reference semantics (oracle) on the same gradients."""
    import numpy as np
    from transformers import LlamaConfig, LlamaForCausalLM

    from multimodal_unlearning_b200 import salun
    from oracle import salun_oracle as SO
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = LlamaForCausalLM(LlamaConfig(vocab_size=512, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                         num_attention_heads=2, num_key_value_heads=2)).to(torch.bfloat16).to(dev)
    g = torch.Generator().manual_seed(4)
    batches = []
    for _ in range(3):
        ids = torch.randint(3, 512, (4, 24), generator=g)
        lab = ids.clone()
        lab[:, :8] = -100
        batches.append((ids, lab, torch.ones_like(ids), torch.arange(4)))
    # the oracle's view: accumulate the same per-batch gradients the drop-in sees
    acc = {}
    model.eval()
    for ids, lab, att, _ in batches:
        model.zero_grad(set_to_none=True)
        model(input_ids=ids.to(dev), attention_mask=att.to(dev), labels=lab.to(dev)).loss.backward()
        for n, p in model.named_parameters():
            gr = p.grad.detach().cpu()
            acc[n] = gr.clone() if n not in acc else SO.accumulate(acc[n], gr)
    want = SO.masks(acc)
    cfg = types.SimpleNamespace(unlearn_lr=1e-5, weight_decay=0.0, save_dir=str(tmp_path))
    salun.save_gradient_ratio(batches, model, None, cfg)
    for th in SO.REFERENCE_THRESHOLDS:
        hard = torch.load(tmp_path / f"with_{th}.pt")
        assert set(hard) == set(acc)
        for n in acc:
            assert hard[n].dtype == torch.int8 and tuple(hard[n].shape) == tuple(acc[n].shape)
            assert np.array_equal(hard[n].numpy(), want[th][n]), (th, n)


# ---- frozen-teacher replay ----------------------------------------------------------
class _CountCalls:
    def __init__(self, module):
        self.n = 0
        self.h = module.register_forward_hook(lambda *a: setattr(self, "n", self.n + 1))


@pytest.mark.parametrize("kind,loss_type", [("mm", "npo"), ("mm", "grad_diff_forget_kl_retain_kl"), ("nlp", "KL")])
def test_teacher_cache_replays_the_same_loss_without_teacher_forwards(kind, loss_type):
    from multimodal_unlearning_b200 import trainer as T
    student, teacher, dev = _models(kind)
    if kind == "mm":
        inputs = {"retain": RH.mm_split(2, 64, 11), "forget": RH.mm_split(2, 64, 12), "idk": RH.mm_split(2, 64, 13)}
        mk = lambda gb: types.SimpleNamespace(loss_type=loss_type, teacher_model=teacher, loss_beta=0.7, l1_lambda=0.0,
                                              l0_lambda=0.0, l_norm_from="zero", loss_args={}, teacher_cache_gb=gb)
        call = lambda fake, inp: T.MMTrainerForgetting.compute_loss(fake, student, inp)
    else:
        inputs = [RH.nlp_triple(2, 64, 21), RH.nlp_triple(2, 64, 22), RH.nlp_triple(2, 64, 23)]
        mk = lambda gb: types.SimpleNamespace(loss_type=loss_type, oracle_model=teacher, loss_beta=1.0, l1_lambda=0.0,
                                              l0_lambda=0.0, teacher_cache_gb=gb)
        call = lambda fake, inp: T.CustomTrainerForgetting.compute_loss(fake, student, inp)
    counter = _CountCalls(teacher)
    plain = call(mk(0), inputs)
    n_plain = counter.n
    assert n_plain >= 1
    cached = mk(1.0)
    first = call(cached, inputs)           # fills the cache (teacher still runs)
    assert counter.n == 2 * n_plain
    student.zero_grad(set_to_none=True)
    second = call(cached, inputs)          # replay: the teacher is not called at all
    assert counter.n == 2 * n_plain
    assert first.item() == plain.item() and second.item() == plain.item()
    second.backward()
    st = cached.teacher_cache.stats()
    assert st["hits"] == n_plain and st["entries"] == 2 * n_plain and st["bytes"] > 0
    # a batch with one changed token is a miss for that sample -> the teacher runs again, result unchanged semantics
    if kind == "mm":
        other = {k: dict(v) for k, v in inputs.items()}
        ids = other["forget"]["input_ids"].clone()
        ids[0, -1] = (ids[0, -1] + 1) % 1000
        other["forget"]["input_ids"] = ids
    else:
        other = [tuple(x.clone() for x in t) for t in inputs]
        other[1][0][0, 5] = (other[1][0][0, 5] + 1) % 1000
    before = counter.n
    changed = call(cached, other)
    assert counter.n > before
    assert changed.item() == call(mk(0), other).item()
    counter.h.remove()


def test_mm_dpo_with_a_legacy_processor_against_the_reference_fixture():
    """A6d end to end: `dpo` on a model that expands the image placeholder inside forward (legacy processor, T' = L + 575
    rows).  tests/golden/legacy_dpo.npz holds the loss and every parameter gradient of the UNMODIFIED reference trainer
    (remove_image_tokens + get_batch_loss on [B, T', V] logits, fp32, CPU); here the trunk's hidden rows are selected by the
    fused head's row_index input (compacted to the labelled rows on the device) — no cropped copy of logits or hidden."""
    import numpy as np

    from multimodal_unlearning_b200 import trainer as T
    from tests._golden import GOLDEN
    dev = torch.device("cuda:0")
    z = np.load(GOLDEN / "legacy_dpo.npz")
    student, teacher = RH.LegacyExpandLM(0).to(dev), RH.LegacyExpandLM(1).to(dev)
    inputs = {"retain": RH.legacy_split(4, 40, 1), "idk": RH.legacy_split(4, 40, 2), "forget": RH.legacy_split(4, 40, 3)}
    fake = types.SimpleNamespace(loss_type="dpo", teacher_model=teacher, loss_beta=1.0, l1_lambda=0.0, l0_lambda=0.0,
                                 l_norm_from="zero", loss_args={})
    loss = T.MMTrainerForgetting.compute_loss(fake, student, inputs)
    loss.backward()
    want = float(z["loss"])
    assert abs(loss.item() - want) <= 2e-3 * abs(want), (loss.item(), want)
    for n, p in student.named_parameters():
        w = torch.from_numpy(z[f"grad::{n}"])
        g = p.grad.detach().float().cpu()
        rel = ((g - w).norm() / w.norm().clamp_min(1e-30)).item()
        assert rel <= 3e-2, (n, rel)                       # the head runs in bf16 (logits, dlogits), the reference in fp32
    assert type(student.lm_head).__name__ == "Linear"
    # the flagged sigmoid log-ratio takes the same four row maps (student / teacher x idk / forget)
    student.zero_grad(set_to_none=True)
    fake.dpo_true = True
    lt = T.MMTrainerForgetting.compute_loss(fake, student, inputs)
    lt.backward()
    assert torch.isfinite(lt) and all(torch.isfinite(p.grad).all() for p in student.parameters())
    # CE recipes on the same model: labels follow their tokens onto the expanded rows, image rows are ignored
    student.zero_grad(set_to_none=True)
    fake.dpo_true, fake.loss_type = False, "grad_ascent"
    got = T.MMTrainerForgetting.compute_loss(fake, student, inputs)
    f = {k: v.to(dev) for k, v in inputs["forget"].items()}
    logits = student(**f).logits.float()
    tp, l = logits.shape[1], f["input_ids"].shape[1]
    exp = torch.full((f["input_ids"].shape[0], tp), -100, dtype=torch.long, device=dev)
    for i in range(exp.shape[0]):
        pos = (f["input_ids"][i] == RH.LEGACY_IMAGE_TOKEN).nonzero().flatten()
        if pos.numel() == 0:
            exp[i, tp - l:] = f["labels"][i]
        else:
            s0 = int(pos[0])
            exp[i, :s0] = f["labels"][i, :s0]
            exp[i, s0 + RH.LEGACY_PATCHES:] = f["labels"][i, s0 + 1:]
    want_ce = -torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), exp[:, 1:].reshape(-1), ignore_index=-100)
    assert abs(got.item() - want_ce.item()) <= 2e-3 * abs(want_ce.item()), (got.item(), want_ce.item())
