"""Histogram / select timing split on the Llama-2-7B parameter list (run under gpurun)."""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import torch
from multimodal_unlearning_b200 import salun, _lib
import bench_salun
dev = torch.device("cuda:0")
lay = salun.FlatLayout(bench_salun.llama7b_shapes(32))
st = salun.SaliencyState(lay, dev)
g = torch.Generator(device=dev).manual_seed(0)
P = lay.total
for o in range(0, P, 1 << 28):
    n = min(1 << 28, P - o)
    st.acc[o:o + n] = (1e-3 * torch.randn(n, generator=g, device=dev)).to(torch.bfloat16)
lib = _lib.load()
n_segs = len(lay.names)
hist = torch.zeros((n_segs, 32768), dtype=torch.int32, device=dev)
s = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
h_ms = t(lambda: lib.clear_saliency_histogram(st.acc.data_ptr(), st.seg_off.data_ptr(), n_segs, lay.total, hist.data_ptr(), s()))
ths = salun.REFERENCE_THRESHOLDS
th = (C.c_double * len(ths))(*ths)
sel = torch.full((n_segs, len(ths), 5), -1, dtype=torch.int64, device=dev)
s_ms = t(lambda: lib.clear_saliency_select(hist.data_ptr(), st.seg_off.data_ptr(), st.seg_numel.data_ptr(), n_segs, th, len(ths), sel.data_ptr(), s()))
z_ms = t(lambda: hist.zero_())
print(f"histogram {h_ms:.3f} ms ({2*P/h_ms/1e6:.0f} GB/s)  select {s_ms:.3f} ms  hist.zero_ {z_ms:.3f} ms")
