"""NVLink evidence for the fused reduce-scatter (run under `gpurun --gpus 2`, ONE process): the dW contraction on GPU 0
reduce-adds the tiles owned by "rank 1" straight into a buffer that lives on GPU 1 (peer access), exactly the path of
distributed.FusedGradReducer but without a second process, a barrier or a collective — so the launch can be profiled with
ncu's NVLink counters (an ncu replay of a multi-rank program with spinning barriers would hang).  Checks the result too."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from multimodal_unlearning_b200 import _lib, head

assert torch.cuda.device_count() >= 2, "needs 2 GPUs"
torch.cuda.set_device(0)
d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
cudart = C.CDLL("libcudart.so")
rc = cudart.cudaDeviceEnablePeerAccess(1, 0)
print("cudaDeviceEnablePeerAccess(1) ->", rc, "(0 = ok, 704 = already enabled); can access:", torch.cuda.can_device_access_peer(0, 1))
cudart.cudaGetLastError()
N, H, V = 8192, 4096, 32064
ld = (V + 63) // 64 * 64
g = torch.Generator(device=d0).manual_seed(0)
dz = (1e-3 * torch.randn(N, ld, generator=g, device=d0)).to(torch.bfloat16)
dz[:, V:] = 0
hid = torch.randn(N, H, generator=g, device=d0).to(torch.bfloat16)
w = torch.empty(V, H, dtype=torch.bfloat16, device=d0)
buf0 = torch.zeros(V, H, dtype=torch.bfloat16, device=d0)
buf1 = torch.zeros(V, H, dtype=torch.bfloat16, device=d1)
buf1[:8].copy_(buf0[:8])          # a torch peer copy (makes sure the runtime has the peer mapping)
torch.cuda.synchronize()
lib = _lib.load()
r0, r1 = C.c_int32(), C.c_int32()
_lib.check(lib.clear_head_dweight_slab(V, 2, 1, 0, C.byref(r0), C.byref(r1)), "slab")
print("rows owned by peer 1:", r0.value, r1.value)
ref = torch.empty(V, H, dtype=torch.bfloat16, device=d0)
head.backward(dz, V, hid, w, None, ref)                      # plain local dW
iters = 3
for _ in range(iters):
    head.backward(dz, V, hid, w, None, None, dweight_peers=[buf0.data_ptr(), buf1.data_ptr()])
torch.cuda.synchronize()
got1 = buf1[r0.value:r1.value].to(d0).float() / iters
got0 = buf0[:r0.value].float() / iters
e1 = ((got1 - ref[r0.value:r1.value].float()).norm() / ref[r0.value:r1.value].float().norm()).item()
e0 = ((got0 - ref[:r0.value].float()).norm() / ref[:r0.value].float().norm()).item()
print(f"peer slab rel err {e1:.3e}, local slab rel err {e0:.3e}; bytes of dW owned by the peer: {(r1.value - r0.value) * H * 2}")
assert e1 < 2e-2 and e0 < 2e-2
assert buf0[r0.value:].abs().sum().item() == 0 and buf1[:r0.value].abs().sum().item() == 0
print("ok")
