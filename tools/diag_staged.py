"""Why is the staged (per-launch events) CE path slower than the one-call path?  CPU and GPU time per step of both."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import head
from multimodal_unlearning_b200._lib import KIND_CE, KIND_NPO
dev = torch.device("cuda:0")
N, H = 8192, 4096
for V in (32000, 32064):
    g = torch.Generator(device=dev).manual_seed(0)
    hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
    y = torch.randint(0, V, (N,), generator=g, device=dev)
    for mode in ("one_call", "staged"):
        for _ in range(3):
            head.head_loss_fwd_bwd(hid, w, KIND_CE, labels=y, coef=-1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cpu = []
        a.record()
        ctx = head.StageTimer() if mode == "staged" else None
        if ctx: ctx.__enter__()
        for _ in range(10):
            t0 = time.perf_counter()
            head.head_loss_fwd_bwd(hid, w, KIND_CE, labels=y, coef=-1.0)
            cpu.append((time.perf_counter() - t0) * 1e3)
        b.record()
        torch.cuda.synchronize()
        summ = ctx.summary() if ctx else {}
        if ctx: ctx.__exit__(None, None, None)
        print(V, mode, "gpu ms/step", round(a.elapsed_time(b) / 10, 3), "cpu ms/step", [round(c, 2) for c in cpu], {k: round(v[1] / 10, 3) for k, v in summ.items()}, flush=True)
