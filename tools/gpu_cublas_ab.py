"""Final-build A/B of the library's tcgen05 contraction against cuBLAS (torch.matmul) on the three contraction shapes of
the loss head at BASELINE configs[1] (N = 8192 rows, H = 4096, V = 32064), from the SAME row-major storage (cuBLAS gets
the transposes as views, i.e. as its own NN / TN / NT modes — no transposed copies on either side):

    projection  Z[N, V]  = Hd[N, H] . W[V, H]^T       ours: K-major x K-major   + LSE epilogue (cuBLAS has none)
    dH          dH[N, H] = dZ[N, V] . W[V, H]          ours: K-major x MN-major
    dW          dW[V, H] = dZ[N, V]^T . Hd[N, H]       ours: MN-major x MN-major

Interleaved (ours, cuBLAS, ours, ...) in groups of `iters` launches, CUDA events on the launching stream; then one
back-to-back second of each ("sustained": the clocks settle under the power cap).  Run under gpurun; writes
gpurun_out/cublas_ab.json and prints a markdown table."""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from multimodal_unlearning_b200 import _lib  # noqa: E402

N, H, V = 8192, 4096, 32064


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator(device=dev).manual_seed(0)
    ld = (V + 63) // 64 * 64
    hid = torch.randn(N, H, generator=g, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(V, H, generator=g, device=dev)).to(torch.bfloat16)
    dz = (1e-3 * torch.randn(N, ld, generator=g, device=dev)).to(torch.bfloat16)
    z = torch.empty(N, ld, dtype=torch.bfloat16, device=dev)
    part = torch.empty(N, (V + 255) // 256, 2, dtype=torch.float32, device=dev)
    dh = torch.empty(N, H, dtype=torch.bfloat16, device=dev)
    dw = torch.empty(V, H, dtype=torch.bfloat16, device=dev)
    zc = torch.empty(N, V, dtype=torch.bfloat16, device=dev)
    stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def ours(a, b, d, m, n, k, am, bm, lse=None):
        args = _lib.GemmArgs(a=a.data_ptr(), b=b.data_ptr(), d=d.data_ptr(), lse_part=lse.data_ptr() if lse is not None else None,
                             lda=a.stride(0), ldb=b.stride(0), ldd=d.stride(0), m=m, n=n, k=k, a_major=am, b_major=bm, d_f32=0,
                             accumulate=0, cta_group=0, raster_m_fast=-1, k_splits=1)
        return lambda: _lib.check(lib.clear_gemm_bf16(C.byref(args), stream()), "clear_gemm_bf16")

    dzv = dz[:, :V]
    cases = [
        ("projection (+LSE epilogue on our side)", ours(hid, w, z, N, V, H, 0, 0, part), lambda: torch.matmul(hid, w.t(), out=zc)),
        ("projection, plain bf16 epilogue", ours(hid, w, z, N, V, H, 0, 0, None), lambda: torch.matmul(hid, w.t(), out=zc)),
        ("dH = dZ . W", ours(dz, w, dh, N, H, V, 0, 1), lambda: torch.matmul(dzv, w, out=dh)),
        ("dW = dZ^T . H", ours(dz, hid, dw, V, H, N, 1, 1), lambda: torch.matmul(dzv.t(), hid, out=dw)),
    ]
    flops = 2.0 * N * H * V
    out = []

    def timed(fn, iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    for name, f_ours, f_cublas in cases:
        for _ in range(3):
            f_ours(); f_cublas()
        torch.cuda.synchronize()
        t_o, t_c = [], []
        for _ in range(6):                      # interleaved groups of 5 launches
            t_o.append(timed(f_ours, 5))
            t_c.append(timed(f_cublas, 5))
        sus_o = timed(f_ours, 600)              # ~0.9 s back to back
        sus_c = timed(f_cublas, 600)
        rec = {"case": name, "ours_ms_best": min(t_o), "cublas_ms_best": min(t_c), "ours_ms_median": sorted(t_o)[3],
               "cublas_ms_median": sorted(t_c)[3], "ours_ms_sustained": sus_o, "cublas_ms_sustained": sus_c,
               "ours_tflops_sustained": flops / sus_o / 1e9, "cublas_tflops_sustained": flops / sus_c / 1e9}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "cublas_ab.json").write_text(json.dumps(out, indent=1))
    print("| contraction | ours best / median / sustained ms | cuBLAS best / median / sustained ms | ours / cuBLAS sustained TFLOP/s |")
    print("|---|---|---|---|")
    for r in out:
        print(f"| {r['case']} | {r['ours_ms_best']:.3f} / {r['ours_ms_median']:.3f} / {r['ours_ms_sustained']:.3f} | "
              f"{r['cublas_ms_best']:.3f} / {r['cublas_ms_median']:.3f} / {r['cublas_ms_sustained']:.3f} | "
              f"{r['ours_tflops_sustained']:.0f} / {r['cublas_tflops_sustained']:.0f} |")


if __name__ == "__main__":
    main()
