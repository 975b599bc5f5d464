"""A/B of head_loss_fwd_bwd options in alternating order (run under gpurun)."""
import sys, time, json
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
def timeit(fn, steps, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
configs = [("npo", KIND_NPO, 4096, True), ("npo", KIND_NPO, 4096, False), ("npo", KIND_NPO, 2048, True), ("npo", KIND_NPO, 8192, False),
           ("ce", KIND_CE, 4096, True), ("ce", KIND_CE, 4096, False), ("ce", KIND_CE, 2048, True), ("ce", KIND_CE, 8192, False)]
res = {}
for rnd in range(3):
    for name, kind, br, ov in configs:
        time.sleep(1.0)  # let the power state recover between runs
        ms = timeit(lambda: head.head_loss_fwd_bwd(hid, w, kind, labels=y, ref_hidden=rh, ref_weight=w, block_rows=br, overlap=ov), steps)
        res.setdefault(f"{name} block={br} overlap={ov}", []).append(round(ms, 3))
for k, v in res.items(): print(k, v, flush=True)
