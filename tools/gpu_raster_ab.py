"""Tile-order A/B for the three contractions of the head at N = 8192 (run under gpurun; under ncu for DRAM bytes).
argv: 'time' (CUDA-event timing, interleaved) or 'once' (each configuration launched once, for an ncu launch list)."""
import ctypes as C, os, sys
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
dz = (0.01 * torch.randn(N, LD, generator=g, device=dev)).to(torch.bfloat16)
z = torch.empty(N, LD, dtype=torch.bfloat16, device=dev)
lse = torch.empty(N, (V + 255) // 256, 2, dtype=torch.float32, device=dev)
dh = torch.empty(N, H, dtype=torch.bfloat16, device=dev)
dw = torch.empty(V, H, dtype=torch.bfloat16, device=dev)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(a, b, d, lse_part, lda, ldb, ldd, m, n, k, am, bm):
    args = _lib.GemmArgs(a=a.data_ptr(), b=b.data_ptr(), d=d.data_ptr(), lse_part=lse_part.data_ptr() if lse_part is not None else None,
                         lda=lda, ldb=ldb, ldd=ldd, m=m, n=n, k=k, a_major=am, b_major=bm, d_f32=0, accumulate=0, cta_group=0,
                         raster_m_fast=-1, k_splits=1)
    _lib.check(lib.clear_gemm_bf16(C.byref(args), st()), "gemm")


shapes = {
    "project": (lambda: gemm(hid, w, z, lse, H, H, LD, N, V, H, 0, 0), [None, "1,32", "1,16", "1,8", "0,16", "0,126"]),
    "dhidden": (lambda: gemm(dz, w, dh, None, LD, H, H, N, H, V, 0, 1), [None, "0,16", "1,8", "1,16", "1,4", "0,8"]),
    "dweight": (lambda: gemm(dz, hid, dw, None, LD, H, H, V, H, N, 1, 1), [None, "0,16", "0,8", "0,4", "1,8", "1,16"]),
}
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
for name, (fn, cfgs) in shapes.items():
    res = {}
    for rep in range(1 if mode == "once" else 5):
        for cfg in cfgs:
            if cfg is None:
                os.environ.pop("CLEAR_GEMM_RASTER", None)
            else:
                os.environ["CLEAR_GEMM_RASTER"] = cfg
            if mode == "once":
                fn(); torch.cuda.synchronize()
                continue
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); fn(); b.record(); torch.cuda.synchronize()
            res.setdefault(cfg, []).append(a.elapsed_time(b) / 2)
    if mode != "once":
        for cfg, ts in res.items():
            print(f"{name:8s} raster {str(cfg):6s} min {min(ts):.3f} ms  median {sorted(ts)[len(ts)//2]:.3f} ms", flush=True)
    else:
        print(name, "order of launches:", cfgs, flush=True)
