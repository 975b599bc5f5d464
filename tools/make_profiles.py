"""profiles/r02_{configs,cublas_ab,block_sizes,bench_final}.* from the JSON / logs a `tools/final_round.sh` run left under gpurun_out/.
Usage: python tools/make_profiles.py [prefix]   (prefix of the run's files, default r02f_)"""
import json, re, shutil, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
O, P = ROOT / "gpurun_out", ROOT / "profiles"
pre = sys.argv[1] if len(sys.argv) > 1 else "r02f_"


def plain_ms(d):
    """ms per step of the loop without stage events (the headline loop since the final build; `value_one_call` before)."""
    return d["ms_per_step"] if "value_with_stage_events" in d else d["value_one_call"]["ms_per_step"]


def staged_ms(d):
    return d["value_with_stage_events"]["ms_per_step"] if "value_with_stage_events" in d else d["ms_per_step"]


def last_json(path):
    return json.loads(Path(path).read_text().strip().splitlines()[-1])


b = last_json(O / f"{pre}bench.json")
(P / "r02_bench_final.json").write_text(json.dumps(b, indent=1) + "\n")
ref = last_json(O / f"{pre}bench_ref.json")
(P / "r02_bench_reference_arm.json").write_text(json.dumps(ref, indent=1) + "\n")

# ---- configs ----
md = ["# Round 2 — every BASELINE.json config on one B200 (loss head only, device-resident inputs)\n",
      "Two sources, same build, same box (`tools/final_round.sh`): (1) `bench.py --config K` — the driver's JSON contract per config",
      "(SURVEY.md §8d numbering: cfg K = BASELINE configs[K - 1]); (2) the `extra_configs` block of the default `bench.py` line",
      "(`tools/bench_configs.py`: 5 timed steps after 2 warm-ups each).  Algorithmic flops only (SURVEY.md §8d: 2·N·H·V per projection /",
      "dH / dW; for compacted CE terms in (1): the flops of the rows that carry a label); peaks from MEASURED_PEAKS.json",
      f"(bf16 {b['roofline']['peak']:.1f} TFLOP/s burst).\n",
      "## (1) `bench.py --config K`\n",
      "| config | loss_type | row compaction | ms/step | tokens/s | same loop with stage events ms/step | e2e tokens/s | contraction TFLOP/s | frac of burst peak | stages ms/step |", "|---|---|---|---|---|---|---|---|---|---|"]
rows = [(2, b)] + [(c, last_json(O / f"{pre}cfg{c}.json")) for c in (3, 4, 5) if (O / f"{pre}cfg{c}.json").exists()]
for c, d in rows:
    r = d["roofline"]
    st = ", ".join(f"{k} {v[1]:.3f}" for k, v in r["stages_ms_per_step"].items())
    md.append(f"| cfg {c} = {d['config']['baseline_config']} | {d['config']['loss_type']} | {d['config']['row_compaction']} | {plain_ms(d):.3f} | {d['config']['rows_per_gpu'] / plain_ms(d) * 1e3:.0f} | "
              f"{staged_ms(d):.3f} | {d['e2e']['value']:.0f} | {r['achieved']:.0f} | {r['frac_of_burst_peak']:.3f} | {st} |")
md += ["", "cfg 4's labels leave 5-15 rows per sequence: after compaction the step is a handful of small launches (0.3-0.4 ms; the 501-K-block",
       "tiles of dH over one row tile bound it), its 'fraction of peak' says nothing about the kernels.  cfg 3 compacts its CE split (65 % of",
       "the forget rows ignored) and evaluates every row of its KL split, as the reference does.\n",
       "## (2) `extra_configs` of the default line\n",
       "| config | ms/step | tokens/s | algorithmic TFLOP | TFLOP/s | of burst peak |", "|---|---|---|---|---|---|"]
for e in b.get("extra_configs", []):
    md.append(f"| {e['config']} | {e['ms_per_step']:.3f} | {e['tokens_per_s']:.0f} | {e['algorithmic_tflops']:.2f} | {e['tflops_per_s']:.0f} | {e['frac_burst_peak']:.3f} |")
md += ["", "The compacted lines count the flops of evaluating EVERY row (the reference's work): their TFLOP/s is an equivalent rate — 2.3x",
       "(65 % ignored) and 11-12x (IDK-length answers) over evaluating all rows.  With dense labels the row map itself costs +10-12 %."]
g = b.get("gpu_eager_baseline")
if g:
    md += ["", "## GPU-side eager baseline (the reference's chain on the same B200)\n",
           f"`gpu_eager_baseline`: {g['kind']}.  **{g['ms_per_step']:.2f} ms per step ({g['value']:.0f} tokens/s), {g['peak_extra_bytes'] / 1e9:.2f} GB of extra",
           f"device memory** against {plain_ms(b):.2f} ms and {b['config'].get('peak_extra_bytes_per_step', 0) / 1e9:.2f} GB here: "
           f"**{g['ms_per_step'] / plain_ms(b):.2f}x** faster at {g['peak_extra_bytes'] / max(b['config'].get('peak_extra_bytes_per_step', 1), 1):.1f}x less memory "
           f"(same loss: {g['loss']:.6f} vs {b['config']['loss']:.6f}).  CPU arm (oracle port of the reference, {ref['cpu_baseline']['cores']} host threads, fp32, "
           f"2048-row sample): {ref['value']:.0f} tokens/s."]
(P / "r02_configs.md").write_text("\n".join(md) + "\n")

# ---- cuBLAS A/B ----
ab = b.get("cublas_ab")
if ab:
    md = ["# Round 2 — final-build A/B of the tcgen05 contraction against cuBLAS\n",
          f"`bench.py: cublas_ab` ({ab['method']}; {ab['shape']}): the library's `clear_gemm_bf16` against `torch.matmul` on the three",
          "contraction shapes of the head, both from the SAME row-major storage (cuBLAS gets the transposes as views: its NN / TN modes; no",
          "transposed copies on either side).\n", "| contraction | ours ms | cuBLAS ms | ours TFLOP/s | cuBLAS TFLOP/s | cuBLAS time / ours |", "|---|---|---|---|---|---|"]
    for c in ab["cases"]:
        md.append(f"| {c['contraction']} | {c['ours_ms']:.3f} | {c['cublas_ms']:.3f} | {c['ours_tflops']:.0f} | {c['cublas_tflops']:.0f} | {c['ours_over_cublas_speed']:.3f} |")
    f = O / "cublas_ab.json"
    if f.exists():
        md += ["", "Longer form (`tools/gpu_cublas_ab.py`, first call of the round, round-1 kernels: best / median of 6 interleaved groups of 5 launches, then",
               "600 launches back to back — 'sustained': both settle at the 1 kW power cap):\n",
               "| contraction | ours best / median / sustained ms | cuBLAS best / median / sustained ms |", "|---|---|---|"]
        for r in json.loads(f.read_text()):
            md.append(f"| {r['case']} | {r['ours_ms_best']:.3f} / {r['ours_ms_median']:.3f} / {r['ours_ms_sustained']:.3f} | {r['cublas_ms_best']:.3f} / {r['cublas_ms_median']:.3f} / {r['cublas_ms_sustained']:.3f} |")
    md += ["", "Reading.  The first table is the FINAL build (fence-free scheduler ring with decoded work items, second TMA thread: tensor pipe 98-99 %",
           "active on all three shapes): ahead of cuBLAS on every contraction of the head, the projection although it also emits the log-sum-exp",
           "partials in its epilogue.  The longer-form table below it dates from the first call of the round (round-1 kernels: within +-3 % on the projection",
           "and dW in bursts, 11-13 % faster on dH, 3-7 % BEHIND cuBLAS back to back under the power cap) and is kept for the comparison."]
    (P / "r02_cublas_ab.md").write_text("\n".join(md) + "\n")

# ---- block sizes ----
f = O / f"{pre}block_ab.log"
if f.exists():
    md = ["# Round 2 — cost of the staging-block size (NS-1)\n",
          "`tools/gpu_block_ab.py`: BASELINE configs[1] (NPO) and a CE step, N = 8192 rows, through the one-call entry; block sizes interleaved",
          "round-robin in one process (15 rounds of 6 steps), so that clock / temperature drift hits every variant alike.  Final build (student +",
          "teacher projection in one launch; dH + dW in one launch; fence-free scheduler ring).\n",
          "| kind | block_rows | staged per operand | median ms | min ms | vs 8192 (median) | vs 8192 (min) |", "|---|---|---|---|---|---|---|"]
    for line in f.read_text().splitlines():
        m = re.match(r"(npo|ce) (\{.*\})", line)
        if not m:
            continue
        d = json.loads(m.group(2))
        base = d["8192"]
        for br, v in d.items():
            md.append(f"| {m.group(1)} | {br} | {int(br) * 32064 * 2 / 1e6:.0f} MB | {v['ms_median']:.3f} | {v['ms_min']:.3f} | {v['ms_median'] / base['ms_median']:.3f} | {v['ms_min'] / base['ms_min']:.3f} |")
    md += ["", "Earlier runs of the round, same tool (7 rounds of 8 steps; other boxes — the pool's boxes differ by several percent in how hard the power cap",
           "bites, ratios move with it):\n",
           "| schedule | NPO 4096 / 2048 / 1024 | CE 4096 / 2048 / 1024 |", "|---|---|---|",
           "| round-1 schedule (one launch per contraction, split-K dH over an fp32 scratch) | +11.7 % / +18.9 % / +35.7 % | +7.2 % / +22.1 % / +45.4 % |",
           "| every pair grouped (projection pair, dH + dW at every size) | +4.1 % / +6.7 % / — | +4.2 % / +8.9 % / — |",
           "| alternating processes on one box, 4096 rows: grouped vs separate | +3.3-5.3 % vs +4.4-6.9 % | +3.9-4.2 % vs +6.2-7.5 % |",
           "| final build before the fence-free ring (same tool as the first table, another box) | +10.3 % / +14.1 % / +31.8 % | +10.1 % / +18.5 % / +45.9 % |",
           "", "Where the cost comes from and why the default stays at 8192 rows (one micro-batch): DESIGN.md §9."]
    (P / "r02_block_sizes.md").write_text("\n".join(md) + "\n")
print("written:", sorted(p.name for p in P.glob("r02_*")))
