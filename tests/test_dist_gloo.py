"""CPU, world_size 2 over gloo: the data-parallel host logic (gradient averaging equals the single-process
gradient of the concatenated batch; sharded global-threshold SalUn equals the single-process mask)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from multimodal_unlearning_b200 import distributed as D
    from oracle import head_oracle as HO
    from oracle import salun_oracle as SO
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)

    # ---- DP equivalence of the head: mean over ranks of per-rank dW == dW of the concatenated batch for
    # losses normalised per row count (KL: 1/N each) -------------------------------------------------------
    g = torch.Generator().manual_seed(0)
    n, h, v = 32, 16, 50
    hid = torch.randn(world * n, h, generator=g)
    ref = torch.randn(world * n, h, generator=g)
    wgt = 0.1 * torch.randn(v, h, generator=g)

    def kl(hh, rr, ww):
        return HO.kl_batchmean(HO.logits2probs(hh @ ww.t()), HO.logits2probs(rr @ ww.t()))

    w_local = wgt.clone().requires_grad_(True)
    kl(hid[rank * n:(rank + 1) * n], ref[rank * n:(rank + 1) * n], w_local).backward()
    dw = w_local.grad.clone()
    red = D.GradReducer(average=True)
    red.wait(red.launch(dw))
    w_all = wgt.clone().requires_grad_(True)
    kl(hid, ref, w_all).backward()
    assert torch.allclose(dw, w_all.grad, atol=1e-6), "DDP mean of per-rank gradients != full-batch gradient"

    # ---- global-threshold SalUn over rank-sliced flat gradients ---------------------------------------
    total = 4000
    gg = torch.Generator().manual_seed(1)
    acc = ((torch.randn(total, generator=gg) * 3).round() / 3).to(torch.bfloat16)   # many ties
    keys = SO.saliency_keys(acc)
    lo, hi = rank * total // world, (rank + 1) * total // world
    hist = torch.from_numpy(np.bincount(keys[lo:hi], minlength=32768).astype(np.int64))
    D.reduce_histogram_(hist)
    assert np.array_equal(hist.numpy(), np.bincount(keys, minlength=32768))
    ths = (0.1, 0.5, 0.9)
    full = {th: SO.mask_of_keys(keys, th) for th in ths}
    taus, budgets, local_ties = [], [], []
    for th in ths:
        k = int(th * total)
        above = 0
        for b in range(32767, -1, -1):
            if k <= above + hist[b].item():
                taus.append(b); budgets.append(k - above); break
            above += hist[b].item()
        local_ties.append(int((keys[lo:hi] == taus[-1]).sum()))
    mine = D.tie_budget_for_rank(torch.tensor(local_ties), torch.tensor(budgets))
    for j, th in enumerate(ths):
        kk = keys[lo:hi]
        m = (kk > taus[j]).astype(np.int8)
        tie_pos = np.nonzero(kk == taus[j])[0][: int(mine[j])]
        m[tie_pos] = 1
        assert np.array_equal(m, full[th][lo:hi]), (rank, th)
    (Path(out_dir) / f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def _ddp_capture_worker(rank, world, port, out_dir):
    """DistributedDataParallel(find_unused_parameters=True) — the reference's setting (mm/forget.py) — around a model whose
    lm_head is swapped for the capture module during the wrapped forward: lm_head.weight gets its gradient later, from a
    loss formed OUTSIDE the wrapped forward (here: plain torch ops standing in for the fused head, which needs a GPU).
    Two wrapped forwards per backward, as in every grad_diff / KL recipe.  Without the capture module's weight-dependent
    placeholder, DDP marks the weight unused and its hook then raises "Expected to mark a variable ready only once"."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from torch.nn.parallel import DistributedDataParallel as DDP
    from multimodal_unlearning_b200 import distributed as D
    from multimodal_unlearning_b200 import trainer as T
    from oracle import head_oracle as HO
    from oracle import ref_harness as RH
    D.init_from_env("gloo")
    model = RH.tiny_llama(0)
    ddp = DDP(model, find_unused_parameters=True)
    batches = [RH.nlp_triple(2, 24, 10 + rank), RH.nlp_triple(2, 24, 20 + rank)]
    for step in range(2):                      # twice: the reducer must be re-armed correctly for the next iteration
        ddp.zero_grad(set_to_none=True)
        loss = 0.0
        for ids, lab, att in batches:
            h, w = T.trunk_hidden(ddp, {"input_ids": ids, "attention_mask": att})
            assert h.requires_grad and w is model.lm_head.weight
            loss = loss + HO.hf_causal_ce(torch.nn.functional.linear(h, w), lab)
        loss.backward()
        g = model.lm_head.weight.grad
        assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
        both = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(both, g.contiguous())
        assert torch.equal(both[0], both[1]), "DDP averaged lm_head.weight.grad over the ranks"
    # the same gradient as the stock head under the same wrapper (mean over ranks of per-rank gradients)
    ref = RH.tiny_llama(0)
    dref = DDP(ref, find_unused_parameters=True)
    loss = sum(dref(input_ids=ids, attention_mask=att, labels=lab).loss for ids, lab, att in batches)
    loss.backward()
    assert torch.allclose(model.lm_head.weight.grad, ref.lm_head.weight.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(model.model.layers[0].mlp.up_proj.weight.grad, ref.model.layers[0].mlp.up_proj.weight.grad, rtol=1e-3, atol=1e-7)
    (Path(out_dir) / f"ddp{rank}").write_text("ok")
    dist.destroy_process_group()


def test_capture_head_under_ddp_find_unused_parameters(tmp_path):
    port = _free_port()
    mp.spawn(_ddp_capture_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ddp0").exists() and (tmp_path / "ddp1").exists()
