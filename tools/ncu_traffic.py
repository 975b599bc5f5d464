"""DRAM traffic per launch of the loss head's kernels from an `ncu --set full` report -> profiles/rNN_head_traffic.json, the
file bench.py reads its `roofline.traffic` from (dram__bytes_read.sum + dram__bytes_write.sum per launch; mean over the
contraction launches of a step).  Usage: python tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/r02_head_traffic.json "<note>" """
import csv
import io
import json
import subprocess
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ni, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    per = {}
    for r in rows[2:]:
        name = r[ni]
        per.setdefault(name, []).append({"dram_read": to_bytes(r[ri], units[ri]), "dram_write": to_bytes(r[wi], units[wi]),
                                         "time": r[ti] + " " + units[ti]})
    gem = [x["dram_read"] + x["dram_write"] for k, v in per.items() if "gemm_bf16_kernel" in k for x in v]
    res = {"source": rep.split("/")[-1], "note": note,
           "contraction_dram_bytes_per_launch_mean": sum(gem) / len(gem) if gem else None,
           "kernels": {k: {"launches": len(v), "dram_read_mean": sum(x["dram_read"] for x in v) / len(v),
                           "dram_write_mean": sum(x["dram_write"] for x in v) / len(v), "time_first": v[0]["time"]} for k, v in per.items()}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:2000])


if __name__ == "__main__":
    main()
