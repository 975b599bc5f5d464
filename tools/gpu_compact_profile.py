"""A few steps of the CE head with dense labels, row compaction off then on (run under `ncu --metrics gpu__time_duration.sum`
to see what the row map itself costs when every row is valid)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import head
from multimodal_unlearning_b200._lib import KIND_CE
dev = torch.device("cuda:0")
N, H, V = 8192, 4096, 32000
g = torch.Generator(device=dev).manual_seed(0)
hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
y = torch.randint(0, V, (N,), generator=g, device=dev)
for compact in (False, True):
    for _ in range(3):
        head.head_loss_fwd_bwd(hid, w, KIND_CE, labels=y, coef=-1.0, compact_ignored=compact)
    torch.cuda.synchronize()
