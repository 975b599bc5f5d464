"""Loss-head throughput on every BASELINE.json config, device-resident inputs, one B200.
bench.py's line is configs[1]; this table is the companion for configs[2..4] (run under gpurun):
    cfg2  npo                              N = 8192,  V = 32064, teacher as (hidden, W)        8 N H V
    cfg3  grad_diff_forget_ce_retain_kl    2 splits x N = 8192 (B 4 x T 2048), V = 32064       6 N H V + 8 N H V
    cfg4  dpo (shipped semantics: mean_b sum_t CE on the IDK split), N = 8192, V = 32064       6 N H V
          dpo_true (flagged sigmoid log-ratio): student IDK + forget (fwd+bwd), teacher IDK + forget (fwd only)
    cfg5  grad_ascent (nlp), N = 8192 (B 16 x T 512), V = 32000                                6 N H V
Flops are the algorithmic ones (recompute never counts); peak = MEASURED_PEAKS.json bf16 (sustained and burst)."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import head
from multimodal_unlearning_b200._lib import KIND_CE, KIND_KL_FWD, KIND_NPO

dev = torch.device("cuda:0")
H = 4096
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
P_SUS, P_BURST = peaks.get("bf16_tflops_sustained", 1400.0), peaks.get("bf16_tflops", 1590.0)


def inputs(n, v, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    hid = torch.randn(n, H, generator=g, device=dev).to(torch.bfloat16)
    ref = (hid.float() + 0.05 * torch.randn(n, H, generator=g, device=dev)).to(torch.bfloat16)
    y = torch.randint(0, v, (n,), generator=g, device=dev)
    return hid, ref, y


def idk_labels(b, t, v, seed):
    """IDK-length answers (5-15 tokens) at the end of each sequence, everything else ignored (SURVEY.md §8d cfg 4)."""
    g = torch.Generator().manual_seed(seed)
    y = torch.full((b, t), -100, dtype=torch.int64)
    for i in range(b):
        k = int(torch.randint(5, 16, (1,), generator=g))
        y[i, t - k:] = torch.randint(0, v, (k,), generator=g)
    return head.shift_labels(y).to(dev)


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def run(steps=10, warmup=3, write=False):
    out = []
    w64 = (0.02 * torch.randn(32064, H, device=dev, generator=torch.Generator(device=dev).manual_seed(1))).to(torch.bfloat16)
    w00 = w64[:32000].contiguous()
    n = 8192
    hid, ref, y = inputs(n, 32064, 1234)
    hid2, ref2, y2 = inputs(n, 32064, 4321)

    def cfg2():
        head.head_loss_fwd_bwd(hid, w64, KIND_NPO, ref_hidden=ref, ref_weight=w64, beta=1.0)

    dw = torch.zeros(32064, H, dtype=torch.float32, device=dev)

    def cfg3():
        dw.zero_()
        head.head_loss_fwd_bwd(hid, w64, KIND_CE, labels=y, coef=-1.0, dweight_accum=dw)
        head.head_loss_fwd_bwd(hid2, w64, KIND_KL_FWD, ref_hidden=ref2, ref_weight=w64, dweight_accum=dw)

    yi = idk_labels(8, 1024, 32064, 3)
    yf = idk_labels(8, 1024, 32064, 4)

    def cfg4():
        head.head_loss_fwd_bwd(hid, w64, KIND_CE, labels=yi, reduction="seq_mean", n_seqs=8)

    def cfg4_true():
        dw.zero_()
        for h_, y_ in ((ref, yi), (ref2, yf)):     # teacher: forward only
            head.head_loss_fwd_bwd(h_, w64, KIND_CE, labels=y_, reduction="seq_mean", n_seqs=8, need_dhidden=False,
                                   need_dweight=False)
        for h_, y_, c in ((hid, yi, 1.0), (hid2, yf, -1.0)):
            head.head_loss_fwd_bwd(h_, w64, KIND_CE, labels=y_, reduction="seq_mean", n_seqs=8, coef=c, dweight_accum=dw)

    y5 = y.clamp(max=31999)

    def cfg5():
        head.head_loss_fwd_bwd(hid, w00, KIND_CE, labels=y5, coef=-1.0)

    # "realistic" label variants of SURVEY.md §8d: the first 65 % of every sequence (image + question) ignored
    y5r = y5.clone().view(16, 512)
    y5r[:, : int(0.65 * 512)] = -100
    y5r = y5r.reshape(-1)

    def cfg5_real():
        head.head_loss_fwd_bwd(hid, w00, KIND_CE, labels=y5r, coef=-1.0)

    def cfg5_real_compact():
        head.head_loss_fwd_bwd(hid, w00, KIND_CE, labels=y5r, coef=-1.0, compact_ignored=True)

    def cfg4_compact():
        head.head_loss_fwd_bwd(hid, w64, KIND_CE, labels=yi, reduction="seq_mean", n_seqs=8, compact_ignored=True)

    def cfg5_compact_dense():
        head.head_loss_fwd_bwd(hid, w00, KIND_CE, labels=y5, coef=-1.0, compact_ignored=True)

    nhv = lambda v: n * H * v
    cases = [("cfg2 npo (N=8192, V=32064, teacher projected)", cfg2, n, 8 * nhv(32064)),
             ("cfg3 grad_diff_forget_ce_retain_kl (2 x N=8192, V=32064, fp32 dW accumulation across splits)", cfg3, 2 * n, 14 * nhv(32064)),
             ("cfg4 dpo, shipped semantics (IDK split N=8192, IDK-length labels)", cfg4, n, 6 * nhv(32064)),
             ("cfg4 dpo_true (student IDK+forget fwd+bwd, teacher IDK+forget fwd)", cfg4_true, 4 * n, 16 * nhv(32064)),
             ("cfg5 grad_ascent nlp (N=8192, V=32000)", cfg5, n, 6 * nhv(32000)),
             ("cfg5, 65 % of the rows ignored (image + question), all rows evaluated like the reference", cfg5_real, n, 6 * nhv(32000)),
             ("cfg5, 65 % ignored, compact_ignored=True (flops counted for all rows: the reference's work)", cfg5_real_compact, n, 6 * nhv(32000)),
             ("cfg4 dpo shipped semantics, compact_ignored=True (5-15 valid rows per sequence of 1024)", cfg4_compact, n, 6 * nhv(32064)),
             ("cfg5 dense labels, compact_ignored=True (every row valid: the cost of the row map itself)", cfg5_compact_dense, n, 6 * nhv(32000))]
    for name, fn, tokens, flops in cases:
        ms = timed(fn, steps, warmup)
        tf = flops / ms / 1e9
        out.append({"config": name, "ms_per_step": ms, "tokens_per_s": tokens / ms * 1e3, "algorithmic_tflops": flops / 1e12,
                    "tflops_per_s": tf, "frac_sustained_peak": tf / P_SUS, "frac_burst_peak": tf / P_BURST})
        if write:
            print(json.dumps(out[-1]), flush=True)
    if not write:
        return out
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "configs_bench.json").write_text(json.dumps(out, indent=1))
    md = ["# Loss-head throughput on every BASELINE.json config (one B200, device-resident inputs)\n",
          "`python tools/bench_configs.py`: 10 timed steps after 3 warm-ups, CUDA events; algorithmic flops only (for the compacted",
          f"rows: the flops of evaluating every row, i.e. the reference's work); peaks from MEASURED_PEAKS.json: bf16 {P_SUS} TFLOP/s",
          f"sustained / {P_BURST} burst.\n",
          "| config | ms/step | tokens/s | algorithmic TFLOP | TFLOP/s | of sustained peak | of burst peak |", "|---|---|---|---|---|---|---|"]
    for c in out:
        md.append(f"| {c['config']} | {c['ms_per_step']:.3f} | {c['tokens_per_s']:.0f} | {c['algorithmic_tflops']:.2f} | "
                  f"{c['tflops_per_s']:.0f} | {c['frac_sustained_peak']:.3f} | {c['frac_burst_peak']:.3f} |")
    md.append("\ncfg4's IDK-length labels leave most rows ignored: without compaction their dlogits rows are written as zeros and the")
    md.append("contractions run over mostly-zero operands (less power, higher clocks than cfg5's dense labels at equal flops); with")
    md.append("compaction only the valid rows are projected at all, so the 'TFLOP/s' of those lines is an equivalent rate, not an achieved one.")
    (ROOT / "gpurun_out" / "configs_bench.md").write_text("\n".join(md) + "\n")


    return out


if __name__ == "__main__":
    run(write=True)
