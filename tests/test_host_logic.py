"""CPU: host-side logic of the drop-in layer (no kernels)."""
import os
import types

import pytest
import torch

from oracle import head_oracle as HO


def test_shift_labels_matches_hf_causal_shift():
    from multimodal_unlearning_b200.head import shift_labels
    y = torch.tensor([[5, -100, 7, 8], [1, 2, -100, -100]])
    s = shift_labels(y)
    assert s.tolist() == [-100, 7, 8, -100, 2, -100, -100, -100]
    assert torch.equal(s, HO.shift_labels(y).reshape(-1))


def test_normalisations_follow_the_reference():
    from multimodal_unlearning_b200 import head
    n, v = 512, 32000
    assert head._normalisation(head.KIND_CE, "reference", n, v, 1.0, None) == (1.0, 1.0, True)
    assert head._normalisation(head.KIND_KL_FWD, "reference", n, v, 1.0, None) == (1 / n, 1 / n, False)
    ls, gs, _ = head._normalisation(head.KIND_NPO, "reference", n, v, 0.5, None)
    assert ls == pytest.approx(2 / (0.5 * n * v)) and gs == pytest.approx(2 / (n * v))
    assert head._normalisation(head.KIND_ENTROPY, "reference", n, v, 1.0, None) == (1.0, 1.0, False)
    assert head._normalisation(head.KIND_CE, "seq_mean", n, v, 1.0, 4) == (0.25, 0.25, False)
    with pytest.raises(ValueError):
        head._normalisation(head.KIND_CE, "seq_mean", n, v, 1.0, None)


def _reference_remove_image_tokens(input_ids, logits, image_token_id):
    """mm/trainer_utils.py:48-70 restated (the oracle of image_row_index)."""
    out = []
    b, l = input_ids.shape
    for i in range(b):
        pos = (input_ids[i] == image_token_id).nonzero().flatten()
        if pos.numel() == 0:
            out.append(logits[i, -l:, :])
            continue
        s = int(pos[0])
        out.append(torch.cat([logits[i, :s, :], logits[i, -(l - s):, :]]))
    return torch.stack(out)


def test_image_row_index_is_remove_image_tokens():
    from multimodal_unlearning_b200.trainer import image_row_index
    b, l, extra, tok = 4, 12, 9, 99
    ids = torch.randint(0, 50, (b, l))
    ids[0, 3] = tok
    ids[1, 0] = tok
    ids[2, l - 1] = tok            # image last; sample 3 has no image
    t_out = l + extra
    logits = torch.randn(b, t_out, 5)
    idx = image_row_index(ids, t_out, tok)
    got = torch.gather(logits, 1, idx[..., None].expand(-1, -1, 5))
    assert torch.equal(got, _reference_remove_image_tokens(ids, logits, tok))
    assert image_row_index(ids, l, tok) is None
    ref_path = "/root/reference/mm/trainer_utils.py"
    if os.path.exists(ref_path):  # the build container: check the restatement against the reference itself
        from oracle.ref_harness import import_reference
        _, MMU, _, _ = import_reference()
        assert torch.equal(got, MMU.remove_image_tokens(ids, logits, tok))


def test_row_map_against_the_reference_fixture():
    """tests/golden/row_map.npz: the rows the UNMODIFIED remove_image_tokens kept on a legacy-processor batch (T' = L + 575,
    image first / in the middle / absent) are the rows image_row_index names; expand_labels puts every text label on its row
    and ignores the image rows."""
    import numpy as np

    from multimodal_unlearning_b200.trainer import expand_labels, flat_row_index, image_row_index
    from tests._golden import GOLDEN
    z = np.load(GOLDEN / "row_map.npz")
    ids, labels, rows = torch.from_numpy(z["input_ids"]), torch.from_numpy(z["labels"]), torch.from_numpy(z["rows"])
    tok = int(z["image_token"])
    tp = z["hidden_bits"].shape[1]
    idx = image_row_index(ids, tp, tok)
    assert torch.equal(idx, rows)
    flat = flat_row_index(idx, tp)
    assert flat.dtype == torch.int32 and torch.equal(flat.view(ids.shape).long(), rows + torch.arange(ids.shape[0])[:, None] * tp)
    exp = expand_labels(labels, ids, tp, tok)
    assert exp.shape == (ids.shape[0], tp)
    text = ids != tok
    assert torch.equal(exp.gather(1, rows)[text], labels[text])
    assert int((exp != -100).sum()) == int(((labels != -100) & text).sum())


def test_partition_range_and_slice_layout():
    from multimodal_unlearning_b200 import distributed as D
    for n, w in ((10, 3), (2, 4), (0, 2), (823, 8)):
        r = D.partition_range(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_loss_needs_teacher_is_case_insensitive():
    from multimodal_unlearning_b200.trainer import loss_needs_teacher
    for t in ("npo", "dpo", "scrub", "llmu", "sku", "KL", "kl", "grad_diff_forget_ce_retain_kl", "grad_diff_forget_KL_retain_ce"):
        assert loss_needs_teacher(t), t
    for t in ("retain_ft", "grad_ascent", "idk", "eco_ft", "grad_diff_forget_ce_retain_ce"):
        assert not loss_needs_teacher(t), t


def test_head_capture_swaps_and_restores_lm_head():
    from transformers import LlamaConfig, LlamaForCausalLM

    from multimodal_unlearning_b200.trainer import trunk_hidden
    torch.manual_seed(0)
    m = LlamaForCausalLM(LlamaConfig(vocab_size=64, hidden_size=16, intermediate_size=32, num_hidden_layers=1,
                                     num_attention_heads=2, num_key_value_heads=2))
    ids = torch.randint(0, 64, (2, 6))
    head_before = m.lm_head
    n_params = sum(1 for _ in m.parameters())
    h, w = trunk_hidden(m, {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ids})
    assert m.lm_head is head_before and sum(1 for _ in m.parameters()) == n_params
    assert h.shape == (2, 6, 16) and h.requires_grad and w is m.lm_head.weight
    out = m(input_ids=ids, labels=ids)
    # captured hidden + weight + shifted labels reproduce HF's outputs.loss
    want = HO.hf_causal_ce(torch.nn.functional.linear(h, w), ids)
    assert torch.allclose(out.loss, want, atol=1e-6)

    class Wrapped(torch.nn.Module):  # DDP-like wrapper
        def __init__(self, mod):
            super().__init__()
            self.module = mod

        def forward(self, **kw):
            return self.module(**kw)

    h2, _ = trunk_hidden(Wrapped(m), {"input_ids": ids})
    assert torch.allclose(h2, h)
    # a failing forward still restores the head
    with pytest.raises(Exception):
        trunk_hidden(m, {"input_ids": ids, "bogus_kwarg_xyz": 1, "position_ids": "bad"})
    assert m.lm_head is head_before


def test_trainer_constructor_surface():
    from multimodal_unlearning_b200 import trainer as T
    import inspect
    src = inspect.getsource(T.MMTrainerForgetting.__init__)
    for k in ("forget_loss", "teacher_model", "loss_beta", "l1_lambda", "l0_lambda", "l_norm_from", "loss_args"):
        assert f'"{k}"' in src
    src = inspect.getsource(T.CustomTrainerForgetting.__init__)
    for k in ("forget_loss", "oracle_model", "loss_beta", "l1_lambda", "l0_lambda"):
        assert f'"{k}"' in src
    sig = inspect.signature(T.MMTrainerForgetting.compute_loss)
    assert list(sig.parameters)[:4] == ["self", "model", "inputs", "return_outputs"]
    fake = types.SimpleNamespace(loss_type="bogus", teacher_model=None, loss_beta=1.0, l1_lambda=0, l0_lambda=0)
    m = torch.nn.Linear(2, 2)
    m.config = types.SimpleNamespace()
    with pytest.raises(ValueError, match="Invalid loss type"):
        T.MMTrainerForgetting.compute_loss(fake, m, {"retain": {"input_ids": torch.zeros(1, 2, dtype=torch.long)}})


def test_flat_layout_pads_to_eight_and_counts_chunks():
    from multimodal_unlearning_b200 import salun
    lay = salun.FlatLayout([("a", (3, 5)), ("b", (8,)), ("c", (0,)), ("d", (20000,))])
    assert lay.offsets == [0, 16, 24, 24, 20024] and lay.total == 20024
    chunk = 8192
    assert lay.chunk_prefix == [0, 1, 2, 2, 2 + (20000 + chunk - 1) // chunk]
    flat = torch.arange(lay.total)
    assert lay.view(flat, 0).shape == (3, 5) and lay.view(flat, 3)[0].item() == 24


def test_partition_units_is_balanced_and_deterministic():
    from multimodal_unlearning_b200.distributed import partition_units
    sizes = [4096 * 11008] * 6 + [4096 * 4096] * 8 + [4096] * 5 + [32000 * 4096]
    parts = partition_units(sizes, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) <= 1.35 * (sum(sizes) / 4)
    assert parts == partition_units(sizes, 4)


def test_teacher_cache_lru_eviction_and_keys():
    from multimodal_unlearning_b200.teacher_cache import TeacherCache
    dev = torch.device("cpu")   # keys and LRU bookkeeping are host logic; rows live wherever the caller puts them
    c = TeacherCache(max_bytes=3 * 4 * 8 * 2)          # room for three [4, 8] bf16 rows
    mk = lambda i: {"input_ids": torch.full((1, 4), i), "attention_mask": torch.ones(1, 4, dtype=torch.long)}
    keys = [TeacherCache.sample_keys(mk(i), None)[0] for i in range(4)]
    assert len(set(keys)) == 4 and all(len(k) == 16 for k in keys)
    rows = [torch.full((1, 4, 8), float(i), dtype=torch.bfloat16, device=dev) for i in range(4)]
    for i in range(3):
        c.store([keys[i]], rows[i])
    assert c.lookup([keys[0]]) is not None              # touch 0: 1 becomes least recently used
    c.store([keys[3]], rows[3])
    assert c.lookup([keys[1]]) is None and c.lookup([keys[0], keys[3]]).shape == (2, 4, 8)
    assert c.stats()["bytes"] == 3 * 4 * 8 * 2
    # attention mask and image bytes are part of the key
    a = mk(7)
    b = {**a, "attention_mask": torch.tensor([[0, 1, 1, 1]])}
    assert TeacherCache.sample_keys(a, None) != TeacherCache.sample_keys(b, None)
    img = {"input_ids": torch.tensor([[5, 9, 9, 2], [5, 1, 1, 2]]), "attention_mask": torch.ones(2, 4, dtype=torch.long),
           "pixel_values": torch.zeros(1, 3, 2, 2)}
    k0 = TeacherCache.sample_keys(img, image_token_id=9)
    img2 = {**img, "pixel_values": torch.ones(1, 3, 2, 2)}
    k1 = TeacherCache.sample_keys(img2, image_token_id=9)
    assert k0[0] != k1[0] and k0[1] == k1[1]            # only the sample that owns the image changes


def test_head_loss_operator_is_registered_with_a_fake_kernel():
    """torch.ops.clear_b200.head_loss: the PyTorch custom-op layer over the C ABI.  Shape / dtype propagation runs without
    a GPU (fake tensors); the real kernel refuses CPU tensors (no fallback)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from multimodal_unlearning_b200 import head  # noqa: F401  (registers the operator)
    op = torch.ops.clear_b200.head_loss
    with FakeTensorMode():
        h = torch.empty(64, 128, dtype=torch.bfloat16, device="cuda")
        w = torch.empty(1000, 128, dtype=torch.bfloat16, device="cuda")
        y = torch.empty(64, dtype=torch.int64, device="cuda")
        loss, seq, dh, dw, rows = op(h, w, 0, y, None, None, None, None, 1.0, 1.0, "seq_mean", 4, 0, True, False, True, None)
        assert loss.shape == (1,) and seq.shape == (4,) and dh.shape == (64, 128) and dw.numel() == 0 and rows.shape == (64,)
        assert dh.dtype == torch.bfloat16 and loss.dtype == torch.float32
        # with a row map: 32 evaluated rows out of the 64 rows of hidden; dH keeps hidden's shape
        idx = torch.empty(32, dtype=torch.int32, device="cuda")
        loss, seq, dh, dw, rows = op(h, w, 0, y[:32], None, None, None, None, 1.0, 1.0, "reference", 0, 0, True, True, False, idx)
        assert dh.shape == (64, 128) and dw.shape == (1000, 128) and rows.shape == (32,)
    with pytest.raises(RuntimeError, match="CUDA"):
        op(torch.zeros(8, 16, dtype=torch.bfloat16), torch.zeros(32, 16, dtype=torch.bfloat16), 0, torch.zeros(8, dtype=torch.int64),
           None, None, None, None, 1.0, 1.0, "reference", 0, 0, True, True, False, None)


def test_enable_fused_head_allreduce_without_a_process_group_is_a_no_op():
    """Single process (no torch.distributed group): nothing to reduce, nothing is touched."""
    from multimodal_unlearning_b200 import head, trainer as T
    model = torch.nn.Sequential(torch.nn.Linear(8, 8))
    assert T.enable_fused_head_allreduce(model) is None
    assert head._DW_REDUCER is None
    assert not getattr(model, "_ddp_params_and_buffers_to_ignore", [])


def test_operator_schema_and_defaults():
    from multimodal_unlearning_b200 import head
    assert head.DEFAULT_BLOCK_ROWS in (2048, 4096, 8192) and head.MAX_BF16_ACCUM_BLOCKS >= 1
    # the operator schema carries the compaction flag and the row map
    schema = str(torch.ops.clear_b200.head_loss.default._schema)
    assert "bool compact_ignored" in schema and "str reduction" in schema and "Tensor? row_index" in schema
