"""GPU probe (run under gpurun): block_rows sweep of the NPO head step, sustained vs cuBLAS on the same
shapes, and per-stage timing.  Writes gpurun_out/probe.json."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import head
from multimodal_unlearning_b200._lib import KIND_NPO, KIND_CE

dev = torch.device("cuda:0")
N, H, V = 8192, 4096, 32064
g = torch.Generator(device=dev).manual_seed(0)
hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
rh = (hid.float() + 0.05 * torch.randn(N, H, generator=g, device=dev)).to(torch.bfloat16)
w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
y = torch.randint(0, V, (N,), generator=g, device=dev)
out = {}

def timeit(fn, steps, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for kind, name in ((KIND_NPO, "npo"), (KIND_CE, "ce")):
    for br in (2048, 4096, 8192):
        ms = timeit(lambda: head.head_loss_fwd_bwd(hid, w, kind, labels=y, ref_hidden=rh, ref_weight=w, block_rows=br), steps)
        out[f"{name}_block{br}_ms"] = ms
        print(name, br, ms, flush=True)
# cuBLAS on the same four contractions, same duration regime
z = torch.empty(N, V, dtype=torch.bfloat16, device=dev)
dh = torch.empty(N, H, dtype=torch.bfloat16, device=dev)
dw = torch.empty(V, H, dtype=torch.bfloat16, device=dev)
def cublas4():
    torch.matmul(hid, w.t(), out=z); torch.matmul(rh, w.t(), out=z)
    torch.matmul(z, w, out=dh); torch.matmul(z.t(), hid, out=dw)
out["cublas_4gemm_ms"] = timeit(cublas4, steps)
print("cublas4", out["cublas_4gemm_ms"], flush=True)
with head.StageTimer() as st:
    for _ in range(10):
        head.head_loss_fwd_bwd(hid, w, KIND_NPO, ref_hidden=rh, ref_weight=w, block_rows=8192)
out["stages_block8192_x10"] = st.summary()
with head.StageTimer() as st:
    for _ in range(10):
        head.head_loss_fwd_bwd(hid, w, KIND_NPO, ref_hidden=rh, ref_weight=w, block_rows=4096)
out["stages_block4096_x10"] = st.summary()
print(json.dumps(out, indent=1))
Path(ROOT / "gpurun_out" / "probe.json").write_text(json.dumps(out, indent=1))
