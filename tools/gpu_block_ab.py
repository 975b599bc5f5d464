"""Cost of the staging-block size (NS-1: "never materialises [B, T, V]"), measured through the one-call entry the trainers
run: BASELINE configs[1] (NPO, N = 8192 rows per step), block_rows in {8192, 4096, 2048, 1024} interleaved round-robin so that
clocks / temperature drift hits every variant alike; median over rounds.  Writes gpurun_out/block_ab.json."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import head
from multimodal_unlearning_b200._lib import KIND_CE, KIND_NPO

dev = torch.device("cuda:0")
N, H, V = 8192, 4096, 32064
g = torch.Generator(device=dev).manual_seed(0)
hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
ref = (hid.float() + 0.05 * torch.randn(N, H, generator=g, device=dev)).to(torch.bfloat16)
w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
y = torch.randint(0, V, (N,), generator=g, device=dev)
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "8192,4096,2048,1024".split(","))]
rounds, steps = 15, 6


OVERLAP = len(sys.argv) > 2 and sys.argv[2] == "overlap"


def run(kind, br):
    ov = OVERLAP and br < N
    if kind == "npo":
        head.head_loss_fwd_bwd(hid, w, KIND_NPO, ref_hidden=ref, ref_weight=w, beta=1.0, block_rows=br, overlap=ov)
    else:
        head.head_loss_fwd_bwd(hid, w, KIND_CE, labels=y, block_rows=br, overlap=ov)


out = {"overlap": OVERLAP}
for kind in ("npo", "ce"):
    times = {br: [] for br in sizes}
    for br in sizes:
        for _ in range(2):
            run(kind, br)
    torch.cuda.synchronize()
    for r in range(rounds):
        for br in sizes:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                run(kind, br)
            b.record()
            torch.cuda.synchronize()
            times[br].append(a.elapsed_time(b) / steps)
    med = {br: sorted(t)[len(t) // 2] for br, t in times.items()}
    out[kind] = {str(br): {"ms_median": med[br], "ms_min": min(times[br]), "vs_largest": med[br] / med[sizes[0]]} for br in sizes}
    print(kind, json.dumps(out[kind]), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / ("block_ab_overlap.json" if OVERLAP else "block_ab.json")).write_text(json.dumps(out, indent=1))
