"""Bring-up check of the tcgen05 GEMM variants against torch on a B200 (run under gpurun).

Each kernel variant runs in its own subprocess under a timeout, so a trapped or wedged variant
cannot take the others (or the box) down.  Writes a summary to gpurun_out/gemm_check.json.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

VARIANTS = [
    # (name, a_major, b_major, d_f32, lse)
    ("kk_bf16", 0, 0, 0, 0),
    ("kk_bf16_lse", 0, 0, 0, 1),
    ("kmn_bf16", 0, 1, 0, 0),
    ("mnmn_bf16", 1, 1, 0, 0),
    ("mnmn_f32", 1, 1, 1, 0),
]
SIZES = [(128, 256, 64), (256, 512, 256), (384, 768, 320), (200, 328, 136), (1024, 2048, 1024)]
BIG = {"kk": (8192, 32064, 4096), "kmn": (8192, 4096, 32064), "mnmn": (32064, 4096, 8192)}


def run_variant(name: str, cta_group: int, big: bool) -> dict:
    import torch
    from multimodal_unlearning_b200._lib import GemmArgs

    lib = C.CDLL(str(ROOT / "multimodal_unlearning_b200" / "libclear_b200.so"))
    lib.clear_last_error.restype = C.c_char_p
    lib.clear_gemm_bf16.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    _, a_major, b_major, d_f32, lse = next(v for v in VARIANTS if v[0] == name)
    dev = torch.device("cuda:0")
    out = {"variant": name, "cta_group": cta_group, "cases": []}
    g = torch.Generator(device="cpu").manual_seed(1234)

    def call(A, B, D, part, m, n, k, accumulate=0):
        args = GemmArgs(a=A.data_ptr(), b=B.data_ptr(), d=D.data_ptr(), lse_part=part.data_ptr() if part is not None else None,
                        lda=A.stride(0), ldb=B.stride(0), ldd=D.stride(0), m=m, n=n, k=k, a_major=a_major,
                        b_major=b_major, d_f32=d_f32, accumulate=accumulate, cta_group=cta_group, raster_m_fast=-1)
        rc = lib.clear_gemm_bf16(C.byref(args), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(f"clear_gemm_bf16 rc={rc}: {lib.clear_last_error().decode()}")

    def make(m, n, k):
        pad = lambda x: (x + 63) // 64 * 64
        Al = torch.randn(m, k, generator=g).to(torch.bfloat16)
        Bl = torch.randn(n, k, generator=g).to(torch.bfloat16)
        # storage: K-major = logical [MN, K]; MN-major = [K, MN] padded to 64 columns
        if a_major == 0:
            A = torch.zeros(m, pad(k), dtype=torch.bfloat16); A[:, :k] = Al; A = A.to(dev)
        else:
            A = torch.zeros(k, pad(m), dtype=torch.bfloat16); A[:, :m] = Al.t(); A = A.to(dev)
        if b_major == 0:
            B = torch.zeros(n, pad(k), dtype=torch.bfloat16); B[:, :k] = Bl; B = B.to(dev)
        else:
            B = torch.zeros(k, pad(n), dtype=torch.bfloat16); B[:, :n] = Bl.t(); B = B.to(dev)
        return Al.to(dev), Bl.to(dev), A, B

    sizes = list(SIZES)
    for (m, n, k) in sizes:
        Al, Bl, A, B = make(m, n, k)
        ldd = (n + 63) // 64 * 64
        D = torch.full((m, ldd), 7.0, dtype=torch.float32 if d_f32 else torch.bfloat16, device=dev)
        nt = (n + 255) // 256
        part = torch.zeros(m, nt, 2, dtype=torch.float32, device=dev) if lse else None
        call(A, B, D, part, m, n, k)
        torch.cuda.synchronize()
        ref = Al.float() @ Bl.float().t()
        got = D[:, :n].float()
        err = (got - ref).abs().max().item()
        scale = ref.abs().max().item()
        tol = (2e-5 if d_f32 else 1e-2) * max(scale, 1.0)
        rec = {"m": m, "n": n, "k": k, "max_abs_err": err, "ref_max": scale, "ok": bool(err <= tol),
               "pad_untouched": bool((D[:, n:] == 7.0).all().item()) if ldd > n else True}
        if lse:
            refb = ref.to(torch.bfloat16).float()
            lse_ref = torch.logsumexp(refb, dim=1)
            # merge partials
            pm, ps = part[..., 0], part[..., 1]
            mx = pm.max(dim=1, keepdim=True).values
            lse_got = (ps * torch.exp(pm - mx)).sum(dim=1).log() + mx[:, 0]
            rec["lse_err"] = (lse_got - lse_ref).abs().max().item()
            rec["ok"] = rec["ok"] and rec["lse_err"] < 2e-2
        if d_f32:
            call(A, B, D, None, m, n, k, accumulate=1)
            torch.cuda.synchronize()
            err2 = (D[:, :n] - 2 * ref).abs().max().item()
            rec["accum_err"] = err2
            rec["ok"] = rec["ok"] and err2 <= 2 * tol
        out["cases"].append(rec)
        print(json.dumps(rec), flush=True)
        if not rec["ok"]:
            # dump a coarse error map to localise layout bugs
            bad = ((got - ref).abs() > tol)
            rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
            cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
            print("bad rows (first 16):", rows, "bad cols (first 16):", cols, "frac bad:", bad.float().mean().item(), flush=True)
            print("got[0,:8]", got[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist(), flush=True)

    if big:
        key = {"kk_bf16": "kk", "kk_bf16_lse": "kk", "kmn_bf16": "kmn", "mnmn_bf16": "mnmn", "mnmn_f32": "mnmn"}[name]
        m, n, k = BIG[key]
        pad = lambda x: (x + 63) // 64 * 64
        A = (torch.randn(m if a_major == 0 else k, pad(k) if a_major == 0 else pad(m), device=dev) * 0.5).to(torch.bfloat16)
        B = (torch.randn(n if b_major == 0 else k, pad(k) if b_major == 0 else pad(n), device=dev) * 0.5).to(torch.bfloat16)
        D = torch.empty(m, pad(n), dtype=torch.float32 if d_f32 else torch.bfloat16, device=dev)
        part = torch.zeros(m, (n + 255) // 256, 2, dtype=torch.float32, device=dev) if lse else None
        for _ in range(3):
            call(A, B, D, part, m, n, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            call(A, B, D, part, m, n, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = 2.0 * m * n * k / ms / 1e9
        # spot check 32 rows
        idx = torch.randint(0, m, (32,), generator=g).to(dev)
        Arow = (A[idx, :k] if a_major == 0 else A[:k, idx].t()).float()
        Bmat = (B[:, :k] if b_major == 0 else B[:k, :n].t()).float()
        ref = Arow @ Bmat.t()
        err = (D[idx, :n].float() - ref).abs().max().item()
        # cuBLAS reference timing for the same logical product
        Ak = (A[:, :k] if a_major == 0 else A[:k, :m].t().contiguous())
        Bk = (B[:, :k] if b_major == 0 else B[:k, :n].t().contiguous())
        for _ in range(3):
            torch.matmul(Ak, Bk.t())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(Ak, Bk.t())
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1) / iters
        rec = {"big": [m, n, k], "ms": ms, "tflops": tf, "spot_err": err, "ref_max": ref.abs().max().item(),
               "cublas_ms": ms_ref, "cublas_tflops": 2.0 * m * n * k / ms_ref / 1e9}
        out["big"] = rec
        print(json.dumps(rec), flush=True)
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant")
    ap.add_argument("--cta-group", type=int, default=1)
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--groups", default="1,2")
    ap.add_argument("--timeout", type=int, default=150)
    a = ap.parse_args()
    if a.variant:
        res = run_variant(a.variant, a.cta_group, a.big)
        print("RESULT " + json.dumps(res))
        return
    summary = []
    for cg in [int(x) for x in a.groups.split(",")]:
        for v in VARIANTS:
            cmd = [sys.executable, __file__, "--variant", v[0], "--cta-group", str(cg)] + (["--big"] if a.big else [])
            t0 = time.time()
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=a.timeout)
                tail = (r.stdout[-3000:] + "\n" + r.stderr[-1500:])
                res = None
                for line in r.stdout.splitlines():
                    if line.startswith("RESULT "):
                        res = json.loads(line[7:])
                summary.append({"variant": v[0], "cta_group": cg, "rc": r.returncode, "result": res,
                                "tail": None if res else tail, "secs": time.time() - t0})
            except subprocess.TimeoutExpired as e:
                summary.append({"variant": v[0], "cta_group": cg, "rc": "timeout",
                                "tail": ((e.stdout or b"")[-2000:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-2000:])})
            s = summary[-1]
            ok = s.get("result") and all(c["ok"] for c in s["result"]["cases"])
            print(f"== {v[0]} cta_group={cg}: rc={s['rc']} ok={ok} big={s.get('result', {}).get('big') if s.get('result') else None}", flush=True)
            if not ok:
                print(s.get("tail") or json.dumps(s.get("result"))[:3000], flush=True)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "gemm_check.json").write_text(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
