"""Interleaved A/B of head_loss_fwd_bwd options in the steady (power-capped) state (run under gpurun).
Configs are visited round-robin, a few steps each, many rounds, after a warm-up; reports mean and spread."""
import sys, time, json, statistics
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import head
from multimodal_unlearning_b200._lib import KIND_NPO, KIND_CE, KIND_KL_FWD
dev = torch.device("cuda:0")
N, H, V = 8192, 4096, 32064
g = torch.Generator(device=dev).manual_seed(0)
hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
rh = (hid.float() + 0.05 * torch.randn(N, H, generator=g, device=dev)).to(torch.bfloat16)
w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
y = torch.randint(0, V, (N,), generator=g, device=dev)
def timeit(fn, steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
which = sys.argv[2] if len(sys.argv) > 2 else "npo"
kind = KIND_NPO if which == "npo" else KIND_CE
configs = {"block=4096": dict(block_rows=4096), "block=8192": dict(block_rows=8192), "block=4096 overlap": dict(block_rows=4096, overlap=True),
           "block=2048": dict(block_rows=2048)}
fns = {k: (lambda kw=kw: head.head_loss_fwd_bwd(hid, w, kind, labels=y, ref_hidden=rh, ref_weight=w, **kw)) for k, kw in configs.items()}
for _ in range(40):  # reach the steady state
    fns["block=4096"]()
torch.cuda.synchronize()
res = {k: [] for k in fns}
for r in range(rounds):
    for k, fn in fns.items():
        res[k].append(timeit(fn, 6))
for k, v in res.items():
    print(f"{which} {k:22s} median {statistics.median(v):.3f} ms  min {min(v):.3f}  max {max(v):.3f}  mean-trim {statistics.mean(sorted(v)[1:-2]):.3f}", flush=True)
