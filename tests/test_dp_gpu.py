"""GPU, 2 ranks over NCCL (skipped on a 1-GPU box): DDP-style mean of per-rank dW equals the gradient of the
concatenated batch; per-rank dH equals the matching rows; the overlapped all-reduce is ordered correctly."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from multimodal_unlearning_b200 import distributed as D
    from multimodal_unlearning_b200 import head
    from multimodal_unlearning_b200._lib import KIND_KL_FWD
    r, w_, local = D.init_from_env("nccl")
    dev = torch.device("cuda", local)
    n, h, v = 512, 256, 32064
    g = torch.Generator().manual_seed(0)
    hid = torch.randn(world * n, h, generator=g).to(torch.bfloat16).to(dev)
    ref = (hid.float().cpu() + 0.3 * torch.randn(world * n, h, generator=g)).to(torch.bfloat16).to(dev)
    wgt = (0.05 * torch.randn(v, h, generator=g)).to(torch.bfloat16).to(dev)
    sl = slice(rank * n, (rank + 1) * n)
    res = head.head_loss_fwd_bwd(hid[sl], wgt, KIND_KL_FWD, ref_hidden=ref[sl], ref_weight=wgt)
    red = D.GradReducer(average=True)
    hnd = red.launch(res.dweight)
    full = head.head_loss_fwd_bwd(hid, wgt, KIND_KL_FWD, ref_hidden=ref, ref_weight=wgt)   # overlaps the all-reduce
    red.wait(hnd)
    torch.cuda.synchronize()
    rel = ((res.dweight.float() - full.dweight.float()).norm() / full.dweight.float().norm()).item()
    assert rel < 2e-2, rel
    # per-rank loss is the mean over its own rows; dH of the full batch carries 1/(R n) instead of 1/n
    relh = ((res.dhidden.float() / world - full.dhidden[sl].float()).norm() / full.dhidden[sl].float().norm()).item()
    assert relh < 2e-2, relh
    # the same mean, fused into the contraction: reduce-scatter in the dW epilogue over NVLink + multicast all-gather
    fused = D.FusedGradReducer((v, h), dev, average=True)
    nccl_dw = res.dweight.clone()
    for it in range(5):                                        # every rotating buffer, a buffer reuse, and (it == 3) the peer-store all-gather
        os.environ["CLEAR_NO_MULTICAST"] = "1" if it == 3 else "0"
        tgt = fused.begin()
        r2 = head.head_loss_fwd_bwd(hid[sl], wgt, KIND_KL_FWD, ref_hidden=ref[sl], ref_weight=wgt, dweight_target=tgt,
                                    block_rows=256 if it == 1 else None)   # it == 1: two row blocks accumulate at the owners
        assert r2.dweight is None
        ev, dw = fused.finish(tgt)
        torch.cuda.current_stream().wait_event(ev)
        torch.cuda.synchronize()
        rel = ((dw.float() - full.dweight.float()).norm() / full.dweight.float().norm()).item()
        assert rel < 2e-2, (it, rel)
        rel2 = ((dw.float() - nccl_dw.float()).norm() / nccl_dw.float().norm()).item()
        assert rel2 < 1e-2, (it, rel2)
        # bit-identical on every rank
        mine = dw.view(torch.int16).to(torch.int32).sum(dtype=torch.int64).reshape(1)
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        assert all(int(x) == int(mine) for x in both)
        gathered = [torch.empty_like(dw) for _ in range(world)]
        dist.all_gather(gathered, dw.contiguous())
        assert all(torch.equal(gx, dw) for gx in gathered)
    losses = torch.stack([res.loss.reshape(())])
    dist.all_reduce(losses, op=dist.ReduceOp.AVG)
    assert abs(losses.item() - full.loss.item()) <= 1e-3 * abs(full.loss.item())
    (Path(out_dir) / f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_two_rank_nccl_dp_equivalence(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def _ddp_worker(rank, world, port, out_dir):
    """The nlp trainer's compute_loss under DistributedDataParallel, twice: stock DDP (NCCL buckets reduce every gradient)
    and with lm_head.weight handed to the fused head all-reduce.  Same gradients on every parameter."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import copy
    import types
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from multimodal_unlearning_b200 import distributed as D
    from multimodal_unlearning_b200 import head
    from multimodal_unlearning_b200 import trainer as T
    from oracle import ref_harness as RH
    D.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    base = RH.tiny_llama(0).to(dev).to(torch.bfloat16)
    teacher = RH.perturbed_copy(RH.tiny_llama(0), 5).to(dev).to(torch.bfloat16)
    inputs = [RH.nlp_triple(2, 64, 100 + rank), RH.nlp_triple(2, 64, 200 + rank), RH.nlp_triple(2, 64, 300 + rank)]
    fake = types.SimpleNamespace(loss_type="KL", oracle_model=teacher, loss_beta=1.0, l1_lambda=0.0, l0_lambda=0.0)

    def run(fused, find_unused=False):
        model = copy.deepcopy(base)
        model.train()
        reducer = T.enable_fused_head_allreduce(model) if fused else None
        assert (reducer is not None) == fused
        ddp = DDP(model, device_ids=[rank], find_unused_parameters=find_unused)
        loss = T.CustomTrainerForgetting.compute_loss(fake, ddp, inputs)   # "KL": two wrapped forwards, one backward
        loss.backward()
        torch.cuda.synchronize()
        head.set_dweight_reducer(None)
        return {n: p.grad.detach().float().clone() for n, p in model.named_parameters()}

    want = run(False)
    got = run(True)
    for n in want:
        rel = ((got[n] - want[n]).norm() / want[n].norm().clamp_min(1e-30)).item()
        assert rel < 3e-2, (n, rel)
    # find_unused_parameters=True is the reference's own DDP setting (mm/forget.py); lm_head.weight receives its gradient
    # from the fused head, outside the wrapped forward, and must still be seen as used (the capture module's placeholder
    # output depends on it) — stock DDP reducer and fused head all-reduce alike
    for fused in (False, True):
        again = run(fused, find_unused=True)
        for n in want:
            rel = ((again[n] - want[n]).norm() / want[n].norm().clamp_min(1e-30)).item()
            assert rel < 3e-2, ("find_unused_parameters", fused, n, rel)
    # and the fused gradient is identical on both ranks
    g = got["lm_head.weight"].contiguous()
    both = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(both, g)
    assert torch.equal(both[0], both[1])
    (Path(out_dir) / f"ddp_ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_trainer_under_ddp_with_fused_head_allreduce(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ddp_ok0").exists() and (tmp_path / "ddp_ok1").exists()


def _salun_worker(rank, world, port, out_dir):
    """Forget batches dealt out to two ranks: all-reduced signed accumulators, tensors dealt out to ranks for the
    per-tensor thresholds, bit-packed masks exchanged — equal to the single-process oracle masks, bit for bit."""
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import numpy as np
    import torch.distributed as dist
    from multimodal_unlearning_b200 import distributed as D
    from multimodal_unlearning_b200 import salun
    from oracle import salun_oracle as SO
    D.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    shapes = [("a", (300, 70)), ("b", (1001,)), ("c", (64, 64)), ("d", (5,)), ("e", (2048, 33))]
    layout = salun.FlatLayout(shapes)
    state = salun.SaliencyState(layout, dev)
    grads = []
    for r in range(world):   # every rank can rebuild every rank's gradients (for the oracle)
        g = torch.Generator().manual_seed(50 + r)
        grads.append({n: ((torch.randn(*s, generator=g) * 4).round() / 4).to(torch.bfloat16) for n, s in shapes})
    flat = torch.zeros(layout.total, dtype=torch.bfloat16)
    for i, (n, _) in enumerate(shapes):
        flat[layout.offsets[i]: layout.offsets[i] + layout.numels[i]] = grads[rank][n].reshape(-1)
    state.accumulate_flat(flat.to(dev))
    ths = (0.1, 0.5, 1.0)
    packed = salun.sharded_masks(state, ths)
    total = {n: SO.accumulate(grads[0][n], grads[1][n]) for n, _ in shapes}
    want = SO.masks(total, ths)
    for t, th in enumerate(ths):
        for i, (n, s) in enumerate(shapes):
            nb = (layout.numels[i] + 7) // 8
            got = salun.unpack_mask(packed[t][layout.offsets[i] // 8: layout.offsets[i] // 8 + nb].clone(), s)
            assert np.array_equal(got.cpu().numpy(), want[th][n]), (th, n)
    # global-threshold scope: flat 1/R slices (tensor "e" is cut between the two ranks), NCCL-reduced 32768-bin histogram,
    # tie budget split by an exclusive prefix over ranks — equal to ONE stable ranking over the concatenation
    ths_g = (0.0, 0.1, 0.37, 0.5, 0.9, 1.0)
    packed = salun.sharded_masks(state, ths_g, reduce_acc=False, scope="global")
    want = SO.masks_global(total, ths_g)
    for t, th in enumerate(ths_g):
        for i, (n, s) in enumerate(shapes):
            nb = (layout.numels[i] + 7) // 8
            got = salun.unpack_mask(packed[t][layout.offsets[i] // 8: layout.offsets[i] // 8 + nb].clone(), s)
            assert np.array_equal(got.cpu().numpy(), want[th][n]), ("global", th, n, int((got.cpu().numpy() != want[th][n]).sum()))
    both = [torch.empty_like(packed[2]) for _ in range(world)]
    dist.all_gather(both, packed[2])
    assert torch.equal(both[0], both[1])
    (Path(out_dir) / f"salun_ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_two_rank_sharded_salun_masks(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_salun_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "salun_ok0").exists() and (tmp_path / "salun_ok1").exists()
