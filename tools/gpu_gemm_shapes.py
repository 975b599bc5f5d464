"""The three contraction launches of the headline step in isolation (projection pair, dH, dW at N = 8192, H = 4096, V = 32064),
for `ncu` A/B runs of library builds.  argv[1] = repetitions (default 3)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import head
dev = torch.device("cuda:0")
N, H, V = 8192, 4096, 32064
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator(device=dev).manual_seed(0)
hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
rh = (hid.float() + 0.05 * torch.randn(N, H, generator=g, device=dev)).to(torch.bfloat16)
w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
ld = (V + 63) // 64 * 64
nt = head.vtiles(V)
z = torch.empty(N, ld, dtype=torch.bfloat16, device=dev); zr = torch.empty_like(z)
part = torch.empty(N, nt, 2, device=dev); rpart = torch.empty_like(part)
dh = torch.empty(N, H, dtype=torch.bfloat16, device=dev)
dw = torch.empty(V, H, dtype=torch.bfloat16, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tot = [0.0, 0.0, 0.0]
for r in range(reps):
    ev[0].record()
    head.project_pair(hid, w, z, part, rh, w, zr, rpart)
    ev[1].record()
    z.mul_(1e-3)   # stands for the row pass: small dlogits
    ev[1].record()
    head.backward(z, V, hid, w, dh, None)
    ev[2].record()
    head.backward(z, V, hid, w, None, dw)
    ev[3].record()
    torch.cuda.synchronize()
    if r:
        for i in range(3):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
n = max(reps - 1, 1)
print("ms: projection pair %.3f  dH %.3f  dW %.3f" % (tot[0] / n, tot[1] / n, tot[2] / n))
