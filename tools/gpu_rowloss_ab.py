"""A/B timing of the row-pass kernels in isolation (run under gpurun). argv[1] = tag."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import head
from multimodal_unlearning_b200._lib import KIND_NPO, KIND_CE, KIND_KL_FWD, KIND_KL_REV, KIND_ENTROPY
dev = torch.device("cuda:0")
N, H, V = 4096, 4096, 32064
g = torch.Generator(device=dev).manual_seed(0)
hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
rh = (hid.float() + 0.05 * torch.randn(N, H, generator=g, device=dev)).to(torch.bfloat16)
w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
y = torch.randint(0, V, (N,), generator=g, device=dev)
ld = (V + 63) // 64 * 64
nt = head.vtiles(V)
z0 = torch.empty(N, ld, dtype=torch.bfloat16, device=dev); zr = torch.empty_like(z0)
part = torch.empty(N, nt, 2, device=dev); rpart = torch.empty_like(part)
head.project(hid, w, z0, part); head.project(rh, w, zr, rpart)
rows = torch.empty(N, device=dev)
z = z0.clone()
import os
if os.environ.get("NPO_VARIANTS"):
    res = {}
    for rep in range(6):
        for var in os.environ["NPO_VARIANTS"].split(","):
            os.environ["CLEAR_NPO_VARIANT"] = var
            z.copy_(z0); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            head.rowloss(KIND_NPO, z, part, rows, V, ref_logits=zr, ref_lse_part=rpart, grad_scale=1e-4)
            b.record(); torch.cuda.synchronize()
            res.setdefault(var, []).append(a.elapsed_time(b) * 1e3)
    for var, ts in res.items():
        print("variant", var, " ".join(f"{t:.1f}" for t in ts), "us  min", f"{min(ts):.1f}", flush=True)
    sys.exit(0)
out = {}
for kind, name in ((KIND_CE, "ce"), (KIND_KL_FWD, "kl_fwd"), (KIND_KL_REV, "kl_rev"), (KIND_NPO, "npo"), (KIND_ENTROPY, "entropy")):
    ref = kind in (KIND_KL_FWD, KIND_KL_REV, KIND_NPO)
    def run():
        head.rowloss(kind, z, part, rows, V, ref_logits=zr if ref else None, ref_lse_part=rpart if ref else None,
                     labels=y if kind == KIND_CE else None, grad_scale=1e-4)
    best = 1e9
    for rep in range(5):
        z.copy_(z0); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    out[name] = best
    print(sys.argv[1] if len(sys.argv) > 1 else "", name, f"{best*1e3:.1f} us", flush=True)
