"""Turn gpurun_out/launches.csv (ncu launch list) and a .ncu-rep (ncu --set full) into the markdown tables kept under
profiles/.  Usage: python tools/ncu_summarize.py launches.csv prof.ncu-rep > profiles/rNN_head_ncu_summary.md"""
import csv, io, subprocess, sys
from collections import OrderedDict


def launch_table(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        rows.append(dict(zip(hdr, r)))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = r["Kernel Name"]
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    out = ["| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:110]}` | {n} | {t/1e3:.2f} | {100*t/total:.1f}% | {t/n:.1f} |")
    return "\n".join(out)


WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("smsp__issue_active.avg.pct", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__cycles_active.avg", "SM active cycles")]


def full_table(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {}
    for key, _ in WANT:
        for i, h in enumerate(hdr):
            if h == key or h.endswith("." + key) or h.endswith(key):
                idx[key] = i
                break
    name_i = hdr.index("Kernel Name")
    lines = ["| kernel | " + " | ".join(lbl for k, lbl in WANT if k in idx) + " |", "|---|" + "---|" * len(idx)]
    for r in rows[2:]:
        cells = []
        for k, _ in WANT:
            if k in idx:
                v = r[idx[k]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[idx[k]]}".strip())
        lines.append(f"| `{r[name_i][:70]}` | " + " | ".join(cells) + " |")
    return "\n".join(lines)


if __name__ == "__main__":
    print("## Launch list (cold-cache, serialised; compare shares)\n")
    print(launch_table(sys.argv[1]))
    print("\n## `ncu --set full` (per launch)\n")
    print(full_table(sys.argv[2]))
