"""Benchmark of the CLEAR unlearning-step loss head on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config {2,3,4,5}]

Metric (BASELINE.json): forget-step loss-head tokens/s (fwd+bwd).  One "step" = one pass of the hot path over one
synthetic batch.  Default workload = BASELINE configs[1] (SURVEY.md §8d "cfg 2"): LLaVA-1.5-7B NPO forget step, bs 8 x
seq 1024 per GPU (N = 8192 rows, H = 4096, V = 32064), frozen teacher given as (hidden_ref, W_ref): student projection,
teacher projection, NPO row pass (loss + dlogits), dH and dW contractions; at N > 1 GPUs each rank owns its own
micro-batch (weak scaling) and dW is averaged over ranks by the all-reduce fused into the dW contraction.
`--config 3 / 4 / 5` select BASELINE configs[2] / [3] / [4] (grad_diff CE + KL at seq 2048; IDK-DPO; Llama-2-7B
grad_ascent bs 16 x seq 512) with the same JSON contract, so the 1/2/4/8-GPU runs of configs[4] can be reproduced.
Prints ONE JSON line (rank 0).

`value`    device-resident inputs, CUDA-event timed, max over ranks; every step runs the one-call C entry
           (clear_head_loss_fwd_bwd: what the trainers run), nothing else on the stream.
`value_with_stage_events`  the same K steps again, the LIBRARY recording CUDA events between its stages
           (ClearHeadArgs.stage_events: projection(s) / row pass / backward contractions): these feed `roofline`.  They stay
           out of the headline loop because an event between two launches keeps the next kernel from filling the SMs
           the previous persistent grid's last tiles leave idle (0.1-0.5 ms per step on the boxes of this pool).
`e2e`      the same step through the public API with HOST (pinned) inputs: every step copies the hidden states, the
           teacher's hidden states and the labels host->device and reads the loss back (copies double-buffered on a side
           stream; all of them inside the timed region).
`roofline` tensor-pipe roofline of the dominant kernel (the tcgen05 bf16 contraction): algorithmic flops of those
           launches / their CUDA-event time, measured live; `frac` is against the BURST cuBLAS peak when the timed
           region is shorter than a second (the sustained figure is beside it); `traffic` comes from the committed ncu
           capture (profiles/*_head_traffic.json), never from a literal.
`cpu_baseline`       the oracle port of the reference's eager loss path (oracle/head_oracle.py), fp32, host cores, bounded sample.
`gpu_eager_baseline` the same eager chain (cuBLAS lm_head + ATen softmax / log / pow + autograd) on the same B200, full N.
`cublas_ab`          final-build A/B of the three contraction shapes against cuBLAS from the same storage.
`extra_configs`      BASELINE configs[2..4] in short form (tokens/s, roofline fraction, row compaction on / off).
`dw_allreduce_check` (N > 1) one step through the fused reducer against one step through NCCL: relative difference and
           bit-identity across ranks; the run FAILS if they disagree.
`--impl reference` times the CPU path alone (the reference is 100 % Python/PyTorch; /root/reference does not travel to
           the GPU box, so the arm runs the oracle port, which restates it line by line).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "forget-step loss-head tokens/s (fwd+bwd)"
UNIT = "tokens/s"
H = 4096
BETA = 1.0

# SURVEY.md §8d numbering: cfg k = BASELINE.json configs[k - 1]
CONFIGS = {
    2: dict(baseline="configs[1]", loss_type="npo", trainer="mm", b=8, t=1024, v=32064, flops_nhv=8, splits=1,
            workload="LLaVA-1.5-7B NPO forget step, loss head only: bs 8 x seq 1024 per GPU (N=8192 rows, H=4096, V=32064), "
                     "frozen teacher as (hidden_ref, W_ref), bf16"),
    3: dict(baseline="configs[2]", loss_type="grad_diff_forget_ce_retain_kl", trainer="mm", b=4, t=2048, v=32064, flops_nhv=14, splits=2,
            workload="LLaVA-1.5-7B grad_diff_forget_ce_retain_kl, loss head only: forget CE + retain KL against the frozen teacher, "
                     "bs 4 x seq 2048 per split and GPU (2 x 8192 rows, H=4096, V=32064), dW of both splits accumulated in fp32, bf16"),
    4: dict(baseline="configs[3]", loss_type="dpo", trainer="mm", b=8, t=1024, v=32064, flops_nhv=6, splits=1,
            workload="LLaVA-1.5-7B IDK-DPO (shipped semantics: mean over sequences of the per-sequence CE on the IDK split), loss head "
                     "only: bs 8 x seq 1024 per GPU (8192 rows, H=4096, V=32064), IDK-length answers (5-15 valid labels per sequence), bf16"),
    5: dict(baseline="configs[4]", loss_type="grad_ascent", trainer="nlp", b=16, t=512, v=32000, flops_nhv=6, splits=1,
            workload="Llama-2-7B text-only grad_ascent (nlp path), loss head only: bs 16 x seq 512 per GPU (8192 rows, H=4096, V=32000), "
                     "dense labels, bf16"),
}


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_burst": d.get("bf16_tflops"), "tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "MEASURED_PEAKS.json (measured)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "B200_PROFILING.md fallback"}


def _committed_traffic():
    """DRAM bytes per contraction launch from the newest committed ncu capture (tools/ncu_traffic.py writes the file)."""
    files = sorted((ROOT / "profiles").glob("r*_head_traffic.json"))
    if not files:
        return None, "no committed ncu capture (profiles/r*_head_traffic.json)"
    d = json.loads(files[-1].read_text())
    return d.get("contraction_dram_bytes_per_launch_mean"), f"{files[-1].name}: {d.get('note', '')}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("CLEAR_BENCH_SMI_MS", "20"), "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed region: only samples taken after this call are reported."""
        self.t_mark = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, ln in self.lines:
            if ts < t_mark:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def _nvlink_counters(index: int):
    """(tx, rx) NVLink data bytes of one GPU since boot, summed over its links (NVML field values), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal) * 1024)   # KiB
        return tuple(out)
    except Exception:
        return None


def idk_labels(b, t, v, seed):
    """IDK-length answers (5-15 tokens, the token lengths of data/idontknow.jsonl) at the end of each sequence, everything else
    ignored (SURVEY.md §8d cfg 4).  Shifted, flat, on the host."""
    import torch
    g = torch.Generator().manual_seed(seed)
    y = torch.full((b, t), -100, dtype=torch.int64)
    for i in range(b):
        k = int(torch.randint(5, 16, (1,), generator=g))
        y[i, t - k:] = torch.randint(0, v, (k,), generator=g)
    return torch.nn.functional.pad(y, (0, 1), value=-100)[..., 1:].reshape(-1).contiguous()


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's eager path
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg_id: int, sample_rows: int, reps: int):
    """tokens/s of lm_head + the config's loss (mm/trainer.py:168-449 / nlp/dataloader.py:144-459 as restated by the oracle)
    forward+backward in fp32 on the host cores, on `sample_rows` rows per split."""
    import torch
    from oracle import head_oracle as HO

    cfg = CONFIGS[cfg_id]
    v = cfg["v"]
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    w = 0.02 * torch.randn(v, H, generator=g)

    def split(rows, seqs=1):
        hid = torch.randn(seqs, rows // seqs, H, generator=g)
        rh = hid + 0.05 * torch.randn(seqs, rows // seqs, H, generator=g)
        labels = torch.randint(0, v, (seqs, rows // seqs), generator=g)
        return hid, rh, labels

    best = float("inf")
    for _ in range(reps):
        if cfg_id == 2:
            hid, rh, labels = split(sample_rows)
            kw = dict(fn=HO.mm_compute_loss, student={"forget": HO.HeadInputs(hid, labels)}, loss_type="npo", teacher={"forget": rh}, ref_weight=w, beta=BETA)
        elif cfg_id == 3:
            hf, _, lf = split(sample_rows)
            hr, rr, lr = split(sample_rows)
            kw = dict(fn=HO.mm_compute_loss, student={"forget": HO.HeadInputs(hf, lf), "retain": HO.HeadInputs(hr, lr)},
                      loss_type="grad_diff_forget_ce_retain_kl", teacher={"retain": rr}, ref_weight=w, beta=BETA)
        elif cfg_id == 4:
            hid, _, labels = split(sample_rows, seqs=max(1, sample_rows // 1024))
            labels[:, :-10] = -100
            kw = dict(fn=HO.mm_compute_loss, student={"idk": HO.HeadInputs(hid, labels)}, loss_type="dpo")
        else:
            hid, _, labels = split(sample_rows)
            kw = dict(fn=HO.nlp_compute_loss, student={"forget": HO.HeadInputs(hid, labels)}, loss_type="grad_ascent")
        fn = kw.pop("fn")
        student = kw.pop("student")
        t0 = time.perf_counter()
        HO.loss_and_grads(fn, student, w, **kw)
        best = min(best, time.perf_counter() - t0)
    return sample_rows * cfg["splits"] / best, torch.get_num_threads(), best


def cpu_salun_rate(n_elems: int = 1 << 24):
    """parameters/s of the reference's per-tensor, per-threshold mask (generate_mask.py:61-72: fp16 |grad| keys, descending
    sort, top-k -> int8) on the host, restated by the oracle (stable sort), on a slice of n_elems synthetic gradients."""
    import torch
    from oracle import salun_oracle as SO

    g = torch.Generator().manual_seed(0)
    acc = (1e-3 * torch.randn(n_elems, generator=g)).to(torch.bfloat16)
    t0 = time.perf_counter()
    keys = SO.saliency_keys(acc)
    mask = SO.mask_of_keys(keys, 0.5)
    dt = time.perf_counter() - t0
    assert int(mask.sum()) == int(0.5 * n_elems)
    return {"value": n_elems / dt, "unit": "parameters/s (one threshold)", "cores": 1, "kind": "port",
            "sample": f"{n_elems} of 6.74e9 parameters, one tensor, threshold 0.5, numpy stable argsort ({dt:.2f} s)"}


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    rows = 2048
    total = 0.0
    times = []
    rate, cores, _ = cpu_reference_rate(args.config, rows, 1)  # warm-up of allocator / thread pool
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_rate(args.config, rows, 1)
    for _ in range(args.steps):
        r, cores, dt = cpu_reference_rate(args.config, rows, 1)
        times.append(dt)
        total += dt
    value = rows * cfg["splits"] * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "baseline_config": cfg["baseline"],
                       "sample": f"{rows} of the 8192 rows per split and step (cost is linear in rows)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{rows} rows x {args.steps} steps, fp32, torch CPU eager (oracle/head_oracle.py "
                                       f"{cfg['trainer']}_compute_loss '{cfg['loss_type']}' + backward)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU-side baselines (timed outside the headline region)
# --------------------------------------------------------------------------------------------------
def gpu_eager_baseline(hidden, ref_hidden, weight, ref_weight, steps=5, warmup=2):
    """The reference's eager chain for configs[1] on the same B200 ("cuBLAS bf16 GEMM + ATen softmax / CE
    eager chain"): bf16 lm_head (cuBLAS), student logits upcast to fp32 as accelerate's autocast wrapper does
    (convert_outputs_to_fp32), bf16 teacher logits under no_grad, mm NPO (mm/trainer.py:343-353, the oracle's restatement
    of it evaluated on the GPU), autograd backward to hidden and lm_head.weight.  Full N = 8192 rows."""
    import torch
    from oracle import head_oracle as HO

    def step():
        h = hidden.detach().requires_grad_(True)
        w = weight.detach().requires_grad_(True)
        z = torch.nn.functional.linear(h, w).float()
        with torch.no_grad():
            zr = torch.nn.functional.linear(ref_hidden, ref_weight)
        loss = HO.npo_mm(z, zr, BETA)
        loss.backward()
        return loss, h.grad, w.grad

    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    for _ in range(warmup):
        loss, _, _ = step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss, dh, dw = step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    peak = torch.cuda.max_memory_allocated() - base
    n = hidden.shape[0]
    del dh, dw
    torch.cuda.empty_cache()
    return {"value": n / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "rows": n, "loss": loss.item(), "peak_extra_bytes": int(peak),
            "kind": "eager PyTorch on the same GPU: cuBLAS bf16 lm_head (student + teacher), fp32 upcast, ATen softmax / pow / log (mm NPO, "
                    "mm/trainer.py:343-353), autograd backward; device-resident inputs, CUDA events, outside the headline region"}


def cublas_ab(n, v, iters=5, groups=4):
    """The library's contraction against cuBLAS (torch.matmul) on the three contraction shapes of this config, from the
    same row-major storage (cuBLAS takes the transposes as views: its NN / TN modes).  Interleaved groups of launches."""
    import ctypes as C

    import torch
    from multimodal_unlearning_b200 import _lib

    dev = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    g = torch.Generator(device=dev).manual_seed(0)
    ld = (v + 63) // 64 * 64
    hid = torch.randn(n, H, generator=g, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(v, H, generator=g, device=dev)).to(torch.bfloat16)
    dz = (1e-3 * torch.randn(n, ld, generator=g, device=dev)).to(torch.bfloat16)
    z = torch.empty(n, ld, dtype=torch.bfloat16, device=dev)
    part = torch.empty(n, (v + 255) // 256, 2, dtype=torch.float32, device=dev)
    dh = torch.empty(n, H, dtype=torch.bfloat16, device=dev)
    dw = torch.empty(v, H, dtype=torch.bfloat16, device=dev)
    zc = torch.empty(n, v, dtype=torch.bfloat16, device=dev)
    stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def ours(a, b, d, m, nn, k, am, bm, lse=None):
        args = _lib.GemmArgs(a=a.data_ptr(), b=b.data_ptr(), d=d.data_ptr(), lse_part=lse.data_ptr() if lse is not None else None,
                             lda=a.stride(0), ldb=b.stride(0), ldd=d.stride(0), m=m, n=nn, k=k, a_major=am, b_major=bm, d_f32=0,
                             accumulate=0, cta_group=0, raster_m_fast=-1, k_splits=1)
        return lambda: _lib.check(lib.clear_gemm_bf16(C.byref(args), stream()), "clear_gemm_bf16")

    dzv = dz[:, :v]
    cases = [("projection (ours: + LSE epilogue)", ours(hid, w, z, n, v, H, 0, 0, part), lambda: torch.matmul(hid, w.t(), out=zc)),
             ("dH = dZ . W", ours(dz, w, dh, n, H, v, 0, 1), lambda: torch.matmul(dzv, w, out=dh)),
             ("dW = dZ^T . H", ours(dz, hid, dw, v, H, n, 1, 1), lambda: torch.matmul(dzv.t(), hid, out=dw))]

    def timed(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    out = []
    flops = 2.0 * n * H * v
    for name, f_ours, f_cublas in cases:
        for _ in range(2):
            f_ours(); f_cublas()
        torch.cuda.synchronize()
        t_o, t_c = [], []
        for _ in range(groups):
            t_o.append(timed(f_ours))
            t_c.append(timed(f_cublas))
        mo, mc = sorted(t_o)[len(t_o) // 2], sorted(t_c)[len(t_c) // 2]
        out.append({"contraction": name, "ours_ms": mo, "cublas_ms": mc, "ours_tflops": flops / mo / 1e9, "cublas_tflops": flops / mc / 1e9,
                    "ours_over_cublas_speed": mc / mo})
    return {"shape": f"N={n}, H={H}, V={v}", "method": f"median of {groups} interleaved groups of {iters} launches, CUDA events", "cases": out}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="SURVEY.md §8d numbering: 2 = BASELINE configs[1] (NPO, the headline), 3 = configs[2], 4 = configs[3], 5 = configs[4]")
    ap.add_argument("--compact", default="auto", choices=["auto", "on", "off"],
                    help="row compaction of CE-type terms (auto: on where the labels leave rows ignored: configs 3 and 4)")
    ap.add_argument("--block-rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-allreduce", action="store_true", help="diagnostic: skip the dW all-reduce at N > 1 (not a valid bench line)")
    ap.add_argument("--allreduce", default="fused", choices=["fused", "nccl"],
                    help="N > 1: dW mean over ranks fused into the contraction (reduce-scatter in the epilogue over NVLink + "
                         "multicast all-gather), or an NCCL all-reduce on a side stream")
    ap.add_argument("--no-salun", action="store_true", help="skip the SalUn mask / masked-AdamW HBM-roofline section")
    ap.add_argument("--no-extras", action="store_true", help="skip gpu_eager_baseline, cublas_ab and extra_configs")
    ap.add_argument("--salun-layers", type=int, default=32, help="Llama-2-7B layer count of the synthetic parameter list")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from multimodal_unlearning_b200 import distributed as D
    from multimodal_unlearning_b200 import head
    from multimodal_unlearning_b200._lib import KIND_CE, KIND_KL_FWD, KIND_NPO

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the loss head has no CPU path (use --impl reference for the CPU arm)")
    rank, world, local = D.init_from_env("nccl")
    if world != args.gpus:
        raise RuntimeError(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = CONFIGS[args.config]
    B, T, V = cfg["b"], cfg["t"], cfg["v"]
    n = B * T
    block_rows = args.block_rows or None
    compact = {"auto": args.config in (3, 4), "on": True, "off": False}[args.compact]

    # CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, best = cpu_reference_rate(args.config, 2048, 3)
        cpu_base = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"2048 of 8192 rows per split, best of 3, fp32 torch CPU eager: lm_head + {cfg['loss_type']} + backward ({best:.2f} s per pass)"}

    # ---- synthetic inputs ----------------------------------------------------------------------------
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    gw = torch.Generator(device=dev).manual_seed(1)
    weight = (0.02 * torch.randn(V, H, generator=gw, device=dev)).to(torch.bfloat16)
    ref_weight = weight.clone()

    def hidden_pair():
        hid = torch.randn(n, H, generator=g, device=dev).to(torch.bfloat16)
        ref = (hid.float() + 0.05 * torch.randn(n, H, generator=g, device=dev)).to(torch.bfloat16)
        return hid, ref

    dev_in = {}
    dev_in["hidden"], dev_in["ref_hidden"] = hidden_pair()
    dev_in["labels"] = torch.randint(0, V, (n,), generator=g, device=dev)      # cfg 2: unused by NPO; part of the batch a step receives
    if args.config == 3:
        dev_in["hidden_r"], dev_in["ref_hidden_r"] = hidden_pair()
        dev_in["labels"].view(B, T)[:, : int(0.65 * T)] = -100                 # image + question part of the forget samples ignored
    if args.config == 4:
        dev_in["labels"] = idk_labels(B, T, V, 3 + rank).to(dev)
    # rows that carry a label (the algorithmic work of a CE term when the ignored rows are compacted away); host read, before any timing
    ce_valid_rows = int((dev_in["labels"] != -100).sum().item()) if args.config != 2 else n
    multi_split = cfg["splits"] > 1
    dw_accum = torch.zeros(V, H, dtype=torch.float32, device=dev) if multi_split else None

    fused = world > 1 and args.allreduce == "fused" and not args.no_allreduce and not multi_split
    fused_note = ""
    if fused:
        # symmetric memory needs peer access between all ranks' GPUs; every rank must take the same branch
        ok = torch.ones(1, device=dev)
        try:
            reducer = D.FusedGradReducer((V, H), dev, average=True)
        except Exception as e:  # noqa: BLE001 - reported in the JSON line, the bench falls back to the NCCL all-reduce
            ok.zero_()
            fused_note = f" (fused reducer unavailable: {type(e).__name__}: {str(e)[:120]})"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            fused = False
            if not fused_note:
                fused_note = " (fused reducer unavailable on another rank)"
    nccl_reducer = D.GradReducer(average=True)
    dws = [None, None]

    def invoke(bufs, **kw):
        """The head invocation(s) of one step of this config; returns the last HeadResult."""
        if args.config == 2:
            return head.head_loss_fwd_bwd(bufs["hidden"], weight, KIND_NPO, ref_hidden=bufs["ref_hidden"], ref_weight=ref_weight, beta=BETA,
                                          block_rows=block_rows, **kw)
        if args.config == 3:
            dw_accum.zero_()
            head.head_loss_fwd_bwd(bufs["hidden"], weight, KIND_CE, labels=bufs["labels"], coef=-BETA, dweight_accum=dw_accum,
                                   block_rows=block_rows, compact_ignored=compact)
            return head.head_loss_fwd_bwd(bufs["hidden_r"], weight, KIND_KL_FWD, ref_hidden=bufs["ref_hidden_r"], ref_weight=ref_weight,
                                          dweight_accum=dw_accum, block_rows=block_rows)
        if args.config == 4:
            return head.head_loss_fwd_bwd(bufs["hidden"], weight, KIND_CE, labels=bufs["labels"], reduction="seq_mean", n_seqs=B,
                                          block_rows=block_rows, compact_ignored=compact, **kw)
        return head.head_loss_fwd_bwd(bufs["hidden"], weight, KIND_CE, labels=bufs["labels"], coef=-1.0, block_rows=block_rows,
                                      compact_ignored=compact, **kw)

    def step(i, bufs):
        if fused:
            target = reducer.begin()
            res = invoke(bufs, dweight_target=target)
            handle, dw = reducer.finish(target)    # (event, averaged dW) — valid through the next step
            dws[i & 1] = (res, handle, dw)
            return res
        res = invoke(bufs)
        handle = None
        if world > 1 and not args.no_allreduce:
            handle = nccl_reducer.launch(res.dweight)
            if multi_split:
                wait_reduced(handle)   # the fp32 accumulator is zeroed again at the start of the next step
        dws[i & 1] = (res, handle)
        return res

    def wait_reduced(handle):
        if handle is not None:
            torch.cuda.current_stream().wait_event(handle)

    def drain():
        for r_h in dws:
            if r_h is not None:
                wait_reduced(r_h[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- N > 1: the fused all-reduce against NCCL, before anything is timed -------------------------------------------
    dw_check = None
    if world > 1 and not args.no_allreduce:
        if fused:
            target = reducer.begin()
            invoke(dev_in, dweight_target=target)
            done, buf = reducer.finish(target)
            wait_reduced(done)
            got = buf.clone()
        else:
            r = invoke(dev_in)
            wait_reduced(nccl_reducer.launch(r.dweight))
            got = r.dweight.to(torch.bfloat16).clone()
        ref = invoke(dev_in).dweight
        ref = ref.clone() if not multi_split else ref.to(torch.bfloat16)
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        rel = ((got.float() - ref.float()).norm() / ref.float().norm()).item()
        first = got.clone()
        dist.broadcast(first, src=0)
        same = torch.tensor([1.0 if torch.equal(first, got) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        dw_check = {"rel": rel, "identical_across_ranks": bool(same.item() == 1.0), "tolerance": 1e-2,
                    "what": ("fused reduce-scatter + multicast all-gather" if fused else "NCCL all-reduce on the side stream")
                            + " vs dist.all_reduce(AVG) of the per-rank bf16 dW, one step before the timed region"}
        if not (rel <= 1e-2 and dw_check["identical_across_ranks"]):
            raise RuntimeError(f"dW all-reduce check failed: {dw_check}")
        del got, ref, first

    # ---- device-resident timing ---------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a moment to start: launched before the warm-up, marked at the timed region
    # peak device memory one step adds on top of its inputs and the weights (workspace with the staged logits + outputs)
    step(0, dev_in)
    drain()
    torch.cuda.synchronize()
    dws[0] = dws[1] = None
    torch.cuda.reset_peak_memory_stats()
    mem0 = torch.cuda.memory_allocated()
    step(0, dev_in)
    drain()
    torch.cuda.synchronize()
    head_peak_extra = int(torch.cuda.max_memory_allocated() - mem0)
    head.StageTimer.reserve(4 * 8 * cfg["splits"] * (args.steps + 8))   # timing events exist before any timed region starts
    # the warm-up below is the last thing before the timed region and steps through the same buffer rotation, so the caching
    # allocator is in its steady state when the clock starts (a cudaMalloc inside the first timed step costs milliseconds)
    for i in range(max(3, args.warmup)):
        res = step(i, dev_in)
    drain()
    barrier()
    launches_per_step = res.launches * (2 if multi_split else 1) + (1 if fused else 0)   # + the all-gather kernel of the fused all-reduce
    # a result kept alive across the timed loop is a FOURTH set of buffers beside the two in `dws` and the one being computed: the
    # caching allocator, warmed up on three, answers the second timed step with a cudaMalloc of a gigabyte (1-40 ms, seen on cfg 5)
    del res

    def timed_loop(tag):
        """K steps between two CUDA events on the launching stream, barrier + synchronize on both sides."""
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dbg = [] if os.environ.get("CLEAR_BENCH_DEBUG") else None
        gc.collect()
        gc.disable()     # a generation-2 collection inside the loop stalls the launching thread for 1-30 ms (seen in the second step)
        barrier()
        t0.record()
        for i in range(args.steps):
            if dbg is not None:
                ev = torch.cuda.Event(enable_timing=True); ev.record(); dbg.append((ev, time.perf_counter()))
            r = step(i, dev_in)
        drain()
        t1.record()
        barrier()
        gc.enable()
        if dbg:
            print(tag, "per-step gpu ms:", [round(dbg[i][0].elapsed_time(dbg[i + 1][0]), 2) for i in range(len(dbg) - 1)],
                  "cpu ms:", [round((dbg[i + 1][1] - dbg[i][1]) * 1e3, 2) for i in range(len(dbg) - 1)], file=sys.stderr)
        return t0.elapsed_time(t1), r.loss   # the scalar only: a HeadResult kept alive would be one more set of buffers

    # ---- the timed region: K steps through the one-call C entry, exactly what the trainers run ------------------------
    nv0 = _nvlink_counters(local) if world > 1 else None
    if rank == 0:
        sampler.mark()
    ms, _ = timed_loop("headline")
    nv1 = _nvlink_counters(local) if world > 1 else None

    # ---- the same K steps again with the library recording CUDA events between its stages (the live roofline numbers) --
    # Kept out of the headline region: an event between two launches makes the next kernel wait for the previous persistent
    # grid to drain completely instead of filling the SMs its last tiles leave idle (measured: 0.1-0.5 ms per step).
    with head.StageTimer() as warm:      # the event pool and the allocator in their steady state for this mode too
        for i in range(2):
            step(i, dev_in)
    warm.summary()
    drain()
    with head.StageTimer() as st:
        ms_staged, loss_t = timed_loop("stage events")
    summ = st.summary()
    clocks = sampler.stop() if rank == 0 else None
    loss_value = loss_t.item()

    # ---- end to end: host buffers in, loss out, copies inside the timed region -------------------
    names = [k for k in dev_in]
    host = {k: dev_in[k].cpu().pin_memory() for k in names}
    copy_stream = torch.cuda.Stream()
    bufs = [{k: torch.empty_like(dev_in[k]) for k in names} for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def upload(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            for k in names:
                bufs[s][k].copy_(host[k], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(k):
        for s in range(2):
            freed[s].record()
        upload(0)
        for i in range(k):
            s = i & 1
            if i + 1 < k:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[s])
            r = step(i, bufs[s])
            freed[s].record()
            loss_host.copy_(r.loss, non_blocking=True)
        drain()

    e2e_loop(max(2, args.warmup))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.collect()
    gc.disable()
    f0.record()
    e2e_loop(args.steps)
    f1.record()
    barrier()
    gc.enable()
    ms_e2e = f0.elapsed_time(f1)
    h2d = sum(t.numel() * t.element_size() for t in host.values())

    # ---- live roofline of the dominant kernel (tcgen05 contraction): events recorded inside the timed region ----
    gemm_ms = sum(v[1] for k, v in summ.items() if k.startswith("gemm")) / args.steps
    gemm_launches = sum(v[0] for k, v in summ.items() if k.startswith("gemm")) // args.steps
    rowloss_ms = summ.get("rowloss", (0, 0.0))[1] / args.steps
    staged_all = bool(gemm_launches)

    extras = {}
    if world == 1 and not args.no_extras:
        del bufs, host
        torch.cuda.empty_cache()
        try:
            if args.config == 2:
                extras["gpu_eager_baseline"] = gpu_eager_baseline(dev_in["hidden"], dev_in["ref_hidden"], weight, ref_weight)
            extras["cublas_ab"] = cublas_ab(n, V)
            sys.path.insert(0, str(ROOT / "tools"))
            import bench_configs
            extras["extra_configs"] = bench_configs.run(steps=5, warmup=2)
        except Exception as e:  # never lose the head's line to a secondary section
            extras["error"] = f"{type(e).__name__}: {e}"

    # ---- SalUn mask + masked AdamW over all parameters of Llama-2-7B (BASELINE configs[3]); HBM roofline ----
    salun_res = None
    if world == 1 and not args.no_salun:
        dev_in.clear()
        torch.cuda.empty_cache()
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import bench_salun
            salun_res = bench_salun.run(layers=args.salun_layers, reps=3, peak_gbs=_peaks()["hbm_gbs"])
            if not args.no_cpu_baseline:
                salun_res["cpu_baseline"] = cpu_salun_rate()
            salun_res["workload"] = ("SalUn over the Llama-2-7B parameter list (synthetic bf16 grads 1e-3*randn): accumulate, "
                                     "histogram+select for 10 thresholds (per-tensor and global scope), one int8 mask, one masked "
                                     "AdamW step; frac = algorithmic bytes / time / measured HBM copy peak")
        except Exception as e:  # never lose the head's line to the secondary section
            salun_res = {"error": f"{type(e).__name__}: {e}"}

    # max over ranks
    t = torch.tensor([ms, ms_e2e, gemm_ms, ms_staged], dtype=torch.float64, device=dev)
    t_min = t.clone()
    nv = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
        if nv0 is not None and nv1 is not None:
            nvt = torch.tensor([nv1[0] - nv0[0], nv1[1] - nv0[1]], dtype=torch.float64, device=dev)
            dist.all_reduce(nvt, op=dist.ReduceOp.SUM)
            nv = (nvt / world / args.steps).tolist()
    ms, ms_e2e, gemm_ms, ms_staged = t.tolist()
    gemm_ms_fastest_rank = t_min[2].item()   # the ranks run in lockstep (one barrier per step): the spread shows in the kernels

    if rank == 0:
        peaks = _peaks()
        tokens = n * cfg["splits"] * world
        flops_step = float(cfg["flops_nhv"]) * n * H * V  # algorithmic: 2NHV per projection / dH / dW, recompute never counts
        if compact and args.config != 2:
            # with row compaction the CE term's algorithmic work is that of its labelled rows (6 NHV of the total belong to it)
            flops_step -= 6.0 * (n - ce_valid_rows) * H * V
        kernel_ms = gemm_ms if staged_all else ms / args.steps
        achieved = flops_step / (kernel_ms * 1e-3) / 1e12
        short = ms < 1000.0                                # timed region shorter than a second: the burst peak is the denominator
        peak = peaks["tflops_burst"] if short else peaks["tflops_sustained"]
        traffic, traffic_note = _committed_traffic()
        line = {
            "metric": METRIC, "value": tokens * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["workload"], "baseline_config": cfg["baseline"], "loss_type": cfg["loss_type"], "rows_per_gpu": n * cfg["splits"],
                       "hidden": H, "vocab": V, "block_rows": block_rows or head.DEFAULT_BLOCK_ROWS, "row_compaction": bool(compact),
                       "ce_rows_with_a_label": ce_valid_rows if args.config != 2 else None,
                       "loss": loss_value, "peak_extra_bytes_per_step": head_peak_extra,
                       "l2": "inputs larger than L2 (W 263 MB + hidden 2 x 67 MB + staged logits vs 126 MB L2); no explicit flush",
                       "dw_allreduce": ("SKIPPED (--no-allreduce diagnostic run)" if args.no_allreduce else
                                        ("fused: dW contraction epilogue TMA-reduce-adds tiles into the owner rank's buffer over NVLink "
                                         "(reduce-scatter), multimem.st all-gather through the NVSwitch multicast mapping, per step"
                                         if fused else "NCCL all-reduce(avg) of dW per step, overlapped with the next step" + fused_note))
                                       if world > 1 else "none (1 GPU)"},
            "value_with_stage_events": {"value": tokens * args.steps / (ms_staged * 1e-3), "ms_per_step": ms_staged / args.steps,
                                        "what": "the same K steps again with the library recording a CUDA event between its stages: the loop "
                                                "`roofline`'s kernel times come from (every event drains the persistent grid before the next launch)"},
            "e2e": {"value": tokens * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "kernel": "gemm_bf16_kernel (tcgen05, cta_group::2): student + teacher projection in one launch, dH and dW",
                         "contraction_stages_per_step": gemm_launches, "kernel_ms_per_step": kernel_ms,
                         "measured": "CUDA events recorded by the library between its stages on the launching stream, over a second loop of "
                                     "the same K steps right after the headline region (value_with_stage_events)",
                         "peak_source": peaks["source"] + (", bf16_tflops (burst: the timed region is %.2f s)" % (ms * 1e-3) if short
                                                           else ", bf16_tflops_sustained (timed region %.1f s)" % (ms * 1e-3)),
                         "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                         "frac_of_sustained_peak": achieved / peaks["tflops_sustained"],
                         "step_frac_of_burst_peak": flops_step / (ms / args.steps * 1e-3) / 1e12 / peaks["tflops_burst"],
                         "rowloss_ms_per_step": rowloss_ms,
                         "kernel_ms_per_step_fastest_rank": gemm_ms_fastest_rank,
                         "stages_ms_per_step": {k: [v[0] // args.steps, v[1] / args.steps] for k, v in sorted(summ.items())}},
            "cpu_baseline": cpu_base,
            "salun_hbm_roofline": salun_res,
            "clocks": clocks,
        }
        line.update(extras)
        if dw_check is not None:
            line["dw_allreduce_check"] = dw_check
        if nv is not None:
            r_ = world
            line["nvlink"] = {"tx_bytes_per_step_per_gpu": nv[0], "rx_bytes_per_step_per_gpu": nv[1],
                              "expected_tx": 2.0 * V * H * ((r_ - 1) / r_ + 1.0 / r_) if fused else None,
                              "expected_rx": 2.0 * V * H * 2 * (r_ - 1) / r_ if fused else None,
                              "source": "NVML NVLINK_THROUGHPUT_DATA_TX / RX field values around the timed region, mean over ranks; expected: "
                                        "reduce-scatter sends (R-1)/R of the bf16 dW and the multicast all-gather 1/R of it once (the switch "
                                        "replicates), every GPU receives (R-1)/R of it twice"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
