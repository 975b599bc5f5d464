"""Benchmark of the CLEAR unlearning-step loss head on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): forget-step loss-head tokens/s (fwd+bwd).  One "step" = one pass of the hot path
over one synthetic batch of BASELINE configs[1]: LLaVA-1.5-7B NPO forget step, bs 8 x seq 1024 per GPU
(N = 8192 rows, H = 4096, V = 32064), frozen teacher given as (hidden_ref, W_ref): student projection,
teacher projection, NPO row pass (loss + dlogits), dH and dW contractions; at N > 1 GPUs each rank owns
its own micro-batch (weak scaling) and dW is averaged over ranks with an NCCL all-reduce that overlaps
the next step's compute.  Prints ONE JSON line (rank 0).

`value`    device-resident inputs, CUDA-event timed, max over ranks.
`e2e`      the same step through the public API with HOST (pinned) inputs: every step copies hidden,
           teacher hidden and labels host->device and reads the loss back (copies double-buffered on a
           side stream; all of them inside the timed region).
`roofline` tensor-pipe roofline of the dominant kernel (the tcgen05 bf16 contraction, 4 launches per
           block of rows): algorithmic flops of those launches / their CUDA-event time, measured live.
`cpu_baseline` the oracle port of the reference's eager loss path (oracle/head_oracle.py), fp32, on the
           host cores, on a bounded sample of the same workload.
`--impl reference` times that CPU path alone (the reference is 100 % Python/PyTorch; /root/reference
           does not travel to the GPU box, so the arm runs the oracle port, which restates it line by line).
"""
from __future__ import annotations

import argparse
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
# BASELINE.json configs[1]
B, T, H, V = 8, 1024, 4096, 32064
BETA = 1.0
WORKLOAD = "LLaVA-1.5-7B NPO forget step, loss head only: bs 8 x seq 1024 per GPU (N=8192 rows, H=4096, V=32064), frozen teacher as (hidden_ref, W_ref), bf16"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_burst": d.get("bf16_tflops"), "tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "MEASURED_PEAKS.json (measured)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "B200_PROFILING.md fallback"}


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
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
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


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's eager path
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_rows: int, reps: int):
    """tokens/s of lm_head + mm NPO (mm/trainer.py:343-353) forward+backward in fp32 on the host cores."""
    import torch
    from oracle import head_oracle as HO

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    hid = torch.randn(1, sample_rows, H, generator=g)
    rh = hid + 0.05 * torch.randn(1, sample_rows, H, generator=g)
    w = 0.02 * torch.randn(V, H, generator=g)
    labels = torch.randint(0, V, (1, sample_rows), generator=g)
    best = float("inf")
    for _ in range(reps):
        student = {"forget": HO.HeadInputs(hid, labels)}
        t0 = time.perf_counter()
        HO.loss_and_grads(HO.mm_compute_loss, student, w, loss_type="npo", teacher={"forget": rh}, ref_weight=w, beta=BETA)
        best = min(best, time.perf_counter() - t0)
    return sample_rows / best, torch.get_num_threads(), best


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
    rows = 2048
    total = 0.0
    times = []
    rate, cores, _ = cpu_reference_rate(rows, 1)  # warm-up of allocator / thread pool
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_rate(rows, 1)
    for _ in range(args.steps):
        r, cores, dt = cpu_reference_rate(rows, 1)
        times.append(dt)
        total += dt
    value = rows * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{rows} of the 8192 rows per step (cost is linear in rows)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{rows} rows x {args.steps} steps, fp32, torch CPU eager (oracle/head_oracle.py mm_compute_loss 'npo' + backward)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--block-rows", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-allreduce", action="store_true", help="diagnostic: skip the dW all-reduce at N > 1 (not a valid bench line)")
    ap.add_argument("--allreduce", default="fused", choices=["fused", "nccl"],
                    help="N > 1: dW mean over ranks fused into the contraction (reduce-scatter in the epilogue over NVLink + "
                         "multicast all-gather), or an NCCL all-reduce on a side stream")
    ap.add_argument("--no-salun", action="store_true", help="skip the SalUn mask / masked-AdamW HBM-roofline section")
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
    from multimodal_unlearning_b200._lib import KIND_NPO

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the loss head has no CPU path (use --impl reference for the CPU arm)")
    rank, world, local = D.init_from_env("nccl")
    if world != args.gpus:
        raise RuntimeError(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = B * T
    block_rows = args.block_rows or None

    # CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, best = cpu_reference_rate(2048, 3)
        cpu_base = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"2048 of 8192 rows, best of 3, fp32 torch CPU eager: lm_head + mm NPO + backward ({best:.2f} s per pass)"}

    # synthetic inputs
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    hidden = torch.randn(n, H, generator=g, device=dev).to(torch.bfloat16)
    ref_hidden = (hidden.float() + 0.05 * torch.randn(n, H, generator=g, device=dev)).to(torch.bfloat16)
    gw = torch.Generator(device=dev).manual_seed(1)
    weight = (0.02 * torch.randn(V, H, generator=gw, device=dev)).to(torch.bfloat16)
    ref_weight = weight.clone()
    labels = torch.randint(0, V, (n,), generator=g, device=dev)  # unused by NPO; part of the batch a step receives

    fused = world > 1 and args.allreduce == "fused" and not args.no_allreduce
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
    if not fused:
        reducer = D.GradReducer(average=True)
    dws = [None, None]

    def step(i, hid, rhid):
        if fused:
            target = reducer.begin()
            res = head.head_loss_fwd_bwd(hid, weight, KIND_NPO, ref_hidden=rhid, ref_weight=ref_weight, beta=BETA,
                                         block_rows=block_rows, dweight_target=target)
            handle, dw = reducer.finish(target)    # (event, averaged dW) — valid for two steps
            dws[i & 1] = (res, handle, dw)
            return res
        res = head.head_loss_fwd_bwd(hid, weight, KIND_NPO, ref_hidden=rhid, ref_weight=ref_weight, beta=BETA,
                                     block_rows=block_rows)
        handle = None if args.no_allreduce else reducer.launch(res.dweight)
        dws[i & 1] = (res, handle)
        return res

    def wait_reduced(handle):
        if handle is None:
            return
        if fused:
            torch.cuda.current_stream().wait_event(handle)
        else:
            reducer.wait(handle)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a moment to start: launched before the warm-up, marked at the timed region
    for i in range(args.warmup):
        res = step(i, hidden, ref_hidden)
    barrier()
    launches_per_step = res.launches + (1 if fused else 0)   # + the all-gather kernel of the fused all-reduce
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.mark()
    e0.record()
    with head.StageTimer() as st:   # CUDA events around every launch of the timed region (the live roofline numbers)
        for i in range(args.steps):
            res = step(i, hidden, ref_hidden)
        for r_h in dws:
            if r_h is not None:
                wait_reduced(r_h[1])
        e1.record()
        barrier()
    summ = st.summary()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    loss_value = res.loss.item()

    # ---- end to end: host buffers in, loss out, copies inside the timed region -------------------
    h_hidden = hidden.cpu().pin_memory()
    h_ref = ref_hidden.cpu().pin_memory()
    h_labels = labels.cpu().pin_memory()
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(hidden), torch.empty_like(ref_hidden), torch.empty_like(labels)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def upload(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            bufs[s][0].copy_(h_hidden, non_blocking=True)
            bufs[s][1].copy_(h_ref, non_blocking=True)
            bufs[s][2].copy_(h_labels, non_blocking=True)
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
            r = step(i, bufs[s][0], bufs[s][1])
            freed[s].record()
            loss_host.copy_(r.loss, non_blocking=True)
        for r_h in dws:
            if r_h is not None:
                wait_reduced(r_h[1])

    e2e_loop(max(2, args.warmup))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(args.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- live roofline of the dominant kernel (tcgen05 contraction): events recorded inside the timed region ----
    gemm_ms = sum(v[1] for k, v in summ.items() if k.startswith("gemm")) / args.steps
    gemm_launches = sum(v[0] for k, v in summ.items() if k.startswith("gemm")) // args.steps
    rowloss_ms = summ.get("rowloss", (0, 0.0))[1] / args.steps

    # ---- SalUn mask + masked AdamW over all parameters of Llama-2-7B (BASELINE configs[3]); HBM roofline ----
    salun_res = None
    if world == 1 and not args.no_salun:
        del bufs, h_hidden, h_ref
        torch.cuda.empty_cache()
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import bench_salun
            salun_res = bench_salun.run(layers=args.salun_layers, reps=3, peak_gbs=_peaks()["hbm_gbs"])
            if not args.no_cpu_baseline:
                salun_res["cpu_baseline"] = cpu_salun_rate()
            salun_res["workload"] = ("SalUn over the Llama-2-7B parameter list (synthetic bf16 grads 1e-3*randn): accumulate, "
                                     "histogram+select for 10 thresholds, one int8 mask, one masked AdamW step; "
                                     "frac = algorithmic bytes / time / measured HBM copy peak")
        except Exception as e:  # never lose the head's line to the secondary section
            salun_res = {"error": f"{type(e).__name__}: {e}"}

    # max over ranks
    t = torch.tensor([ms, ms_e2e, gemm_ms], dtype=torch.float64, device=dev)
    t_min = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
    ms, ms_e2e, gemm_ms = t.tolist()
    gemm_ms_fastest_rank = t_min[2].item()   # the ranks run in lockstep (one barrier per step): the spread shows in the kernels

    if rank == 0:
        peaks = _peaks()
        tokens = n * world
        flops_step = 8.0 * n * H * V  # algorithmic: student fwd 2NHV + teacher fwd 2NHV + dH 2NHV + dW 2NHV
        achieved = flops_step / (gemm_ms * 1e-3) / 1e12
        peak = peaks["tflops_sustained"]
        h2d = hidden.numel() * 2 + ref_hidden.numel() * 2 + labels.numel() * 8
        line = {
            "metric": METRIC, "value": tokens * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "loss_type": "npo", "rows_per_gpu": n, "hidden": H, "vocab": V,
                       "block_rows": block_rows or head.DEFAULT_BLOCK_ROWS, "loss": loss_value,
                       "l2": "inputs larger than L2 (W 263 MB + hidden 2 x 67 MB + staged logits 2 x 525 MB vs 126 MB L2); no explicit flush",
                       "dw_allreduce": ("SKIPPED (--no-allreduce diagnostic run)" if args.no_allreduce else
                                        ("fused: dW contraction epilogue TMA-reduce-adds tiles into the owner rank's buffer over NVLink "
                                         "(reduce-scatter), multimem.st all-gather through the NVSwitch multicast mapping, per step"
                                         if fused else "NCCL all-reduce(avg) of bf16 dW per step, overlapped with the next step" + fused_note))
                                       if world > 1 else "none (1 GPU)"},
            "e2e": {"value": tokens * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": 1.65e9, "traffic_note": "dram bytes read+written per launch, mean of the 4 contraction launches of a step (projection 1.16, teacher projection 1.18, dH 2.51, dW 1.77 GB), ncu --set full at 8192-row blocks (profiles/r01_head_ncu_summary.md); algorithmic operand+result bytes per launch 0.855 GB",
                         "kernel": "gemm_bf16_kernel (tcgen05, cta_group::2)",
                         "launches_per_step": gemm_launches, "kernel_ms_per_step": gemm_ms,
                         "peak_source": peaks["source"] + ", bf16_tflops_sustained (kernel timed inside a long step)",
                         "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                         "step_frac": flops_step / (ms / args.steps * 1e-3) / 1e12 / peak,
                         "rowloss_ms_per_step": rowloss_ms,
                         "kernel_ms_per_step_fastest_rank": gemm_ms_fastest_rank,
                         "stages_ms_per_step": {k: [v[0] // args.steps, v[1] / args.steps] for k, v in sorted(summ.items())}},
            "cpu_baseline": cpu_base,
            "salun_hbm_roofline": salun_res,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
