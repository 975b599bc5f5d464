"""Probe (2+ GPUs, torchrun): symmetric memory over NVLink — peer copies by the copy engines, TMA store / reduce-add
of the tcgen05 GEMM epilogue straight into a PEER's buffer, multicast availability."""
import ctypes as C, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm
from multimodal_unlearning_b200 import _lib
from multimodal_unlearning_b200 import distributed as D

rank, world, local = D.init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
lib = _lib.load()
V, H, N = 32064, 4096, 8192
log = lambda *a: print(f"[r{rank}]", *a, flush=True)

t = sm.empty(V * H, dtype=torch.bfloat16, device=dev)
hdl = sm.rendezvous(t, dist.group.WORLD)
log("rendezvous ok: world", hdl.world_size, "multicast", hdl.has_multicast_support, hex(hdl.multicast_ptr) if hdl.has_multicast_support else None,
    "ptrs", [hex(p) for p in hdl.buffer_ptrs])
peer = (rank + 1) % world

# (b) copy-engine peer copy bandwidth
src = torch.randn(V * H, device=dev).to(torch.bfloat16)
pbuf = hdl.get_buffer(peer, (V * H,), torch.bfloat16)
hdl.barrier()
for _ in range(2):
    pbuf.copy_(src)
torch.cuda.synchronize(); hdl.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    pbuf.copy_(src)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
log(f"peer copy 263 MB: {ms:.3f} ms = {V*H*2/ms/1e6:.0f} GB/s")
hdl.barrier()
# verify what landed in MY buffer came from rank-1
prev = (rank - 1) % world
g = torch.Generator(device=dev)
dist.barrier()

# (c) GEMM epilogue into the peer's buffer: dW = dZ^T H with D = peer pointer
torch.manual_seed(100 + rank)
dz = (0.01 * torch.randn(N, V, device=dev)).to(torch.bfloat16)
hid = torch.randn(N, H, device=dev).to(torch.bfloat16)
def gemm(dptr, accumulate):
    a_ = _lib.GemmArgs(a=dz.data_ptr(), b=hid.data_ptr(), d=dptr, lse_part=None, lda=dz.stride(0), ldb=hid.stride(0), ldd=H,
                       m=V, n=H, k=N, a_major=1, b_major=1, d_f32=0, accumulate=accumulate, cta_group=0, raster_m_fast=0, k_splits=1)
    _lib.check(lib.clear_gemm_bf16(C.byref(a_), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "gemm")
local_out = torch.empty(V, H, dtype=torch.bfloat16, device=dev)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms_local = timeit(lambda: gemm(local_out.data_ptr(), 0))
hdl.barrier()
ms_peer_store = timeit(lambda: gemm(hdl.buffer_ptrs[peer], 0))
hdl.barrier(); torch.cuda.synchronize(); dist.barrier()
# my buffer now holds prev rank's product; compare with a local recomputation of prev's inputs
torch.manual_seed(100 + prev)
dz_p = (0.01 * torch.randn(N, V, device=dev)).to(torch.bfloat16)
hid_p = torch.randn(N, H, device=dev).to(torch.bfloat16)
want = (dz_p.float().t() @ hid_p.float())
got = t.view(V, H).float()
log(f"TMA store into peer: local {ms_local:.3f} ms, peer {ms_peer_store:.3f} ms, rel err {((got-want).norm()/want.norm()).item():.2e}")
dist.barrier()
# reduce-add into the peer: zero, then every rank adds its product into BOTH its own buffer... here: into peer's, twice
t.zero_(); torch.cuda.synchronize(); dist.barrier(); hdl.barrier()
ms_peer_red = timeit(lambda: gemm(hdl.buffer_ptrs[peer], 1), n=2)   # 3 adds in total (1 warm + 2)
hdl.barrier(); torch.cuda.synchronize(); dist.barrier()
got = t.view(V, H).float()
log(f"TMA reduce-add into peer: {ms_peer_red:.3f} ms, rel err vs 3x product {((got-3*want).norm()/(3*want).norm()).item():.2e}")
# (d) multicast store
if hdl.has_multicast_support:
    try:
        dist.barrier(); hdl.barrier()
        if rank == 0:
            gemm(hdl.multicast_ptr, 0)
        torch.cuda.synchronize(); dist.barrier(); hdl.barrier()
        torch.manual_seed(100)
        dz0 = (0.01 * torch.randn(N, V, device=dev)).to(torch.bfloat16); hid0 = torch.randn(N, H, device=dev).to(torch.bfloat16)
        want0 = dz0.float().t() @ hid0.float()
        log(f"TMA store to MULTICAST ptr from rank 0: rel err on this rank {((t.view(V,H).float()-want0).norm()/want0.norm()).item():.2e}")
    except Exception as e:
        log("multicast store failed:", e)
dist.barrier()
dist.destroy_process_group()
