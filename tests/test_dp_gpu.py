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
