"""Data-parallel SalUn mask generation on R GPUs (run under torchrun): `salun.sharded_masks` over the Llama-2-7B parameter list
(291 tensors, 6.74 B elements, synthetic bf16 gradient sums), both scopes, 10 thresholds — time per stage and a full-size check
that every rank ends with the masks one GPU computes alone, bit for bit.
    python -m torch.distributed.run --nproc-per-node R --master-addr 127.0.0.1 tools/bench_sharded_masks.py [--layers 32]"""
import argparse, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import torch
import torch.distributed as dist
from multimodal_unlearning_b200 import salun, distributed as D
from bench_salun import llama7b_shapes

ap = argparse.ArgumentParser(); ap.add_argument("--layers", type=int, default=32); args = ap.parse_args()
rank, world, local = D.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
lay = salun.FlatLayout(llama7b_shapes(args.layers))
P = lay.total
st = salun.SaliencyState(lay, dev)
g = torch.Generator(device=dev).manual_seed(7)          # the SAME sums on every rank: stands for the all-reduced accumulator
for o in range(0, P, 1 << 28):
    n = min(1 << 28, P - o)
    st.acc[o:o + n] = (1e-3 * torch.randn(n, generator=g, device=dev)).to(torch.bfloat16)
ths = salun.REFERENCE_THRESHOLDS
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn):
    dist.barrier(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), r


out = {"gpus": world, "params": P, "tensors": len(lay.names), "thresholds": len(ths)}
buf = st.acc.clone()
timed(lambda: dist.all_reduce(buf, op=dist.ReduceOp.SUM))
ms, _ = timed(lambda: dist.all_reduce(buf, op=dist.ReduceOp.SUM))
out["allreduce_acc_bf16"] = {"ms": ms, "bytes": 2 * P, "busbw_gbs": 2 * P * 2 * (world - 1) / world / ms / 1e6}
del buf
for scope in ("per_tensor", "global"):
    def alone():
        sel = st.select(ths, scope)
        return [st.mask_flat(sel, j, packed=True) for j in range(len(ths))]
    alone(); ms1, ref = timed(alone)
    salun.sharded_masks(st, ths, reduce_acc=False, scope=scope)
    msr, got = timed(lambda: salun.sharded_masks(st, ths, reduce_acc=False, scope=scope))
    same = all(torch.equal(a, b) for a, b in zip(ref, got))
    ones = [int(salun.unpack_mask(m, P).sum(dtype=torch.int64).item()) if False else None for m in got[:0]]
    flag = torch.tensor([1.0 if same else 0.0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out[scope] = {"one_gpu_alone_ms": ms1, "sharded_ms": msr, "speedup": ms1 / msr, "bit_identical_to_one_gpu_on_every_rank": bool(flag.item() == 1.0)}
    del ref, got
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
