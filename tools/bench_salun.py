"""SalUn mask / masked-AdamW throughput on B200 against the HBM roofline (run under gpurun).
Workload: BASELINE configs[3] — saliency over all parameters of Llama-2-7B (291 tensors, 6.74 B elements)
laid out as the real tensor list, synthetic bf16 gradients 1e-3*randn, thresholds 0.1..1.0, one AdamW step.
Prints a JSON object; --scale divides the layer count for quick runs."""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import salun


def llama7b_shapes(layers=32, vocab=32000, h=4096, ff=11008):
    s = [("model.embed_tokens.weight", (vocab, h))]
    for i in range(layers):
        p = f"model.layers.{i}."
        s += [(p + "self_attn.q_proj.weight", (h, h)), (p + "self_attn.k_proj.weight", (h, h)),
              (p + "self_attn.v_proj.weight", (h, h)), (p + "self_attn.o_proj.weight", (h, h)),
              (p + "mlp.gate_proj.weight", (ff, h)), (p + "mlp.up_proj.weight", (ff, h)), (p + "mlp.down_proj.weight", (h, ff)),
              (p + "input_layernorm.weight", (h,)), (p + "post_attention_layernorm.weight", (h,))]
    s += [("model.norm.weight", (h,)), ("lm_head.weight", (vocab, h))]
    return s


def run(layers=32, reps=3, peak_gbs=6525.2, adam=True):
    dev = torch.device("cuda:0")
    lay = salun.FlatLayout(llama7b_shapes(layers))
    P = lay.total
    st = salun.SaliencyState(lay, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    grads = torch.empty(P, dtype=torch.bfloat16, device=dev)
    step = 1 << 28
    for o in range(0, P, step):
        n = min(step, P - o)
        grads[o:o + n] = (1e-3 * torch.randn(n, generator=g, device=dev)).to(torch.bfloat16)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=reps):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = ev(), ev(); a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    out = {"params": P, "tensors": len(lay.names)}
    ms = timed(lambda: st.accumulate_flat(grads))
    out["accumulate"] = {"ms": ms, "bytes_per_param": 6, "gbs": 6 * P / ms / 1e6, "frac": 6 * P / ms / 1e6 / peak_gbs}
    ths = salun.REFERENCE_THRESHOLDS
    sel = st.select(ths)
    ms = timed(lambda: st.select(ths))
    out["histogram+select(10 thresholds)"] = {"ms": ms, "bytes_per_param": 2, "gbs": 2 * P / ms / 1e6, "frac": 2 * P / ms / 1e6 / peak_gbs}
    ms = timed(lambda: st.select(ths, scope="global"))
    out["histogram+select(10 thresholds, global scope)"] = {"ms": ms, "bytes_per_param": 2, "gbs": 2 * P / ms / 1e6,
                                                            "frac": 2 * P / ms / 1e6 / peak_gbs}
    mask = torch.empty(P, dtype=torch.int8, device=dev)
    sel_g = st.select(ths, scope="global")
    ms = timed(lambda: st.mask_flat(sel_g, 4, out=mask))
    out["mask(one threshold, global scope)"] = {"ms": ms, "bytes_per_param": 3, "gbs": 3 * P / ms / 1e6, "frac": 3 * P / ms / 1e6 / peak_gbs,
                                               "ones": int(mask.sum(dtype=torch.int64).item()), "expected_ones": int(ths[4] * sum(lay.numels))}
    ms = timed(lambda: st.mask_flat(sel, 4, out=mask))
    out["mask(one threshold)"] = {"ms": ms, "bytes_per_param": 3, "gbs": 3 * P / ms / 1e6, "frac": 3 * P / ms / 1e6 / peak_gbs,
                                  "ones": int(mask.sum(dtype=torch.int64).item())}
    bits = torch.empty(P // 8, dtype=torch.uint8, device=dev)
    ms = timed(lambda: st.mask_flat(sel, 4, out=bits, packed=True))
    b = 2.125
    out["mask_bits(one threshold)"] = {"ms": ms, "bytes_per_param": b, "gbs": b * P / ms / 1e6, "frac": b * P / ms / 1e6 / peak_gbs}
    from multimodal_unlearning_b200 import norms
    ms = timed(lambda: norms.clip_coef(norms.grad_sqnorm([grads], [bits]), 1.0))
    out["grad_sqnorm(masked, packed)+clip_coef"] = {"ms": ms, "bytes_per_param": b, "gbs": b * P / ms / 1e6,
                                                     "frac": b * P / ms / 1e6 / peak_gbs}
    if adam:
        param = torch.zeros(P, dtype=torch.bfloat16, device=dev)
        m = torch.zeros(P, dtype=torch.float32, device=dev)
        v = torch.zeros(P, dtype=torch.float32, device=dev)
        ms = timed(lambda: salun.masked_adamw_flat(param, grads, m, v, mask, 1))
        out["masked_adamw"] = {"ms": ms, "bytes_per_param": 23, "gbs": 23 * P / ms / 1e6, "frac": 23 * P / ms / 1e6 / peak_gbs}
        coef = norms.clip_coef(norms.grad_sqnorm([grads], [bits]), 1.0)
        ms = timed(lambda: salun.masked_adamw_flat(param, grads, m, v, bits, 1, grad_scale=coef))
        b = 22.125
        out["masked_adamw(packed mask, fused clip)"] = {"ms": ms, "bytes_per_param": b, "gbs": b * P / ms / 1e6,
                                                        "frac": b * P / ms / 1e6 / peak_gbs}
        del m, v
        torch.cuda.empty_cache()
        s1, s2, a1, a2 = salun.adamw8_init(P, dev)
        salun.masked_adamw8_flat(param, grads, s1, s2, a1, a2, bits, 1, grad_scale=coef)     # states away from zero
        ms = timed(lambda: salun.masked_adamw8_flat(param, grads, s1, s2, a1, a2, bits, 2, grad_scale=coef))
        b = 2 + 2 + 2 + 1 + 1 + 1 + 1 + 0.125 + 4 * 4 / 256     # p r/w, g, two 8-bit states r/w, packed mask, two absmax r/w per 256
        out["masked_adamw_8bit_states(packed mask, fused clip)"] = {"ms": ms, "bytes_per_param": b, "gbs": b * P / ms / 1e6,
                                                                    "frac": b * P / ms / 1e6 / peak_gbs,
                                                                    "optimizer_state_bytes_per_param": 2 + 8 / 256}
        del s1, s2, a1, a2
        ms = timed(lambda: norms.param_l1l0([param], [grads]))
        out["param_l1l0(vs reference params)"] = {"ms": ms, "bytes_per_param": 4, "gbs": 4 * P / ms / 1e6, "frac": 4 * P / ms / 1e6 / peak_gbs}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    res = run(a.layers, a.reps)
    print(json.dumps(res, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "salun_bench.json").write_text(json.dumps(res, indent=1))
