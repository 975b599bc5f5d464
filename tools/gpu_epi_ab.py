"""Projection contraction with and without the LSE epilogue (run under gpurun): what the online log-sum-exp costs."""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
N, H, V = 8192, 4096, 32064
LD = (V + 63) // 64 * 64
g = torch.Generator(device=dev).manual_seed(0)
hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
z = torch.empty(N, LD, dtype=torch.bfloat16, device=dev)
lse = torch.empty(N, (V + 255) // 256, 2, dtype=torch.float32, device=dev)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def gemm(with_lse):
    a = _lib.GemmArgs(a=hid.data_ptr(), b=w.data_ptr(), d=z.data_ptr(), lse_part=lse.data_ptr() if with_lse else None, lda=H, ldb=H, ldd=LD,
                      m=N, n=V, k=H, a_major=0, b_major=0, d_f32=0, accumulate=0, cta_group=0, raster_m_fast=-1, k_splits=1)
    _lib.check(lib.clear_gemm_bf16(C.byref(a), st()), "gemm")
res = {True: [], False: []}
for rep in range(6):
    for flag in (True, False):
        gemm(flag); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gemm(flag); gemm(flag); gemm(flag); b.record(); torch.cuda.synchronize()
        res[flag].append(a.elapsed_time(b) / 3)
for flag, ts in res.items():
    print("LSE epilogue" if flag else "plain bf16 epilogue", " ".join(f"{t:.3f}" for t in ts), "min", f"{min(ts):.3f} ms", flush=True)
