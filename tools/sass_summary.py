"""profiles/rNN_sass_summary.txt: mnemonic counts that prove what the library's kernels are made of (cuobjdump -sass of the
in-tree libclear_b200.so).  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "multimodal_unlearning_b200" / "libclear_b200.so"
WANT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|LDTM(?:\.x\d+)?|UTMALDG\.2D(?:\.2CTA)?|UTMASTG\.2D|UTMAREDG\.2D\.ADD|UTCBAR(?:\.2CTA\.MULTICAST)?|UTCATOMSWS|STAS(?:\.128)?|"
                  r"LDG\.E\.ENL2\.256|STG\.E\.ENL2\.256|MUFU\.EX2|MUFU\.LG2|MUFU\.RCP|MEMBAR\.[A-Z]+\.[A-Z]+|ERRBAR|CGAERRBAR|SYNCS\.ARRIVE[.A-Z0-9]*)\b")
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
dem = subprocess.run(["cu++filt"], input=sass, capture_output=True, text=True).stdout or sass
total = collections.Counter()
per = collections.OrderedDict()
cur = None
for line in dem.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        cur = m.group(1).strip()
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for t in WANT.findall(line):
        total[t] += 1
        per[cur][t] += 1
print("# SASS evidence (cuobjdump -sass multimodal_unlearning_b200/libclear_b200.so, sm_100a) — round 2, final build\n")
print("Mnemonic counts over the whole library (tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, TMA load / store / reduce-add = UTMALDG / UTMASTG / UTMAREDG,")
print("tcgen05.commit -> mbarrier = UTCBAR, st.async (the scheduler ring's hand-over: 16 bytes + complete_tx) = STAS.128, 256-bit global accesses =")
print("LDG/STG.E.ENL2.256; MEMBAR / ERRBAR / CGAERRBAR = fences: in the contraction kernels the start-up / teardown cluster barriers, the counters' re-arm")
print("and the generic -> async proxy fence in front of each TMA store of the epilogue (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC) are left — none on the path of")
print("the MMA-issuing thread or of the two TMA threads, none around any mbarrier hand-over):\n")
for k in sorted(total):
    print(f"    {k:28s} {total[k]}")
print("\nPer kernel (kernels that use tensor-core / TMA / st.async instructions):\n")
for fn, c in per.items():
    if any(k.startswith(("UTC", "UTMA", "LDTM", "STAS")) for k in c):
        print("    " + fn[:150])
        print("        " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if not k.startswith("MUFU")))
