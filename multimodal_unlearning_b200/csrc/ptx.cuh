// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld / fences) and cluster helpers.  Nothing here is portable to other architectures and
// nothing here tries to be: the library is built with -gencode arch=compute_100a,code=sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace clear {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Arrive on the barrier at the same smem offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 raddr;\n"
      "mapa.shared::cluster.u32 raddr, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [raddr];\n"
      "}" ::"r"(bar),
      "r"(cta)
      : "memory");
}

// The same for "I am done READING" hand-backs (a consumer returning a scheduler slot or a TMEM accumulator stage to the
// leader CTA): nothing this thread wrote has to become visible, so the arrive carries the default semantics instead of
// .release.cluster — the latter compiles to MEMBAR.ALL.CTA (.GPU after a global store) + ERRBAR in front of the
// SYNCS.ARRIVE, which held the single MMA-issuing thread ~1.7 us per output tile (ncu source page, round 2: 6.5 % of its
// samples) while the tensor pipe drained.  The tcgen05 side is ordered by tcgen05.fence::before_thread_sync / after_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster_noflush(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 raddr;\n"
      "mapa.shared::cluster.u32 raddr, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [raddr];\n"
      "}" ::"r"(bar),
      "r"(cta)
      : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// try_wait that acquires at cluster scope: for data another CTA of the cluster stored into this CTA's shared memory
// (st.shared::cluster) before its mbarrier.arrive.release.cluster.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// 32-bit store into the shared memory of CTA `cta` of the cluster, at the same offset as local address `addr`.
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t addr, uint32_t cta, uint32_t value) {
  asm volatile(
      "{\n"
      ".reg .b32 raddr;\n"
      "mapa.shared::cluster.u32 raddr, %0, %1;\n"
      "st.shared::cluster.u32 [raddr], %2;\n"
      "}" ::"r"(addr),
      "r"(cta), "r"(value)
      : "memory");
}

// Hand 16 bytes to CTA `cta` of the cluster and signal its mbarrier in ONE asynchronous operation: the store carries
// complete_tx on the destination CTA's barrier, so data and signal are ordered by the hardware and the sender needs no fence
// (st.shared::cluster + mbarrier.arrive.release.cluster costs MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR per destination).  The
// barrier's phase is armed by mbar_arrive_expect_tx_cluster (16 bytes); the receiver waits with a plain try_wait, as for TMA.
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t bar, uint32_t cta, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile(
      "{\n"
      ".reg .b32 raddr, rbar;\n"
      "mapa.shared::cluster.u32 raddr, %0, %2;\n"
      "mapa.shared::cluster.u32 rbar, %1, %2;\n"
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [raddr], {%3, %4, %5, %6}, [rbar];\n"
      "}" ::"r"(addr),
      "r"(bar), "r"(cta), "r"(a), "r"(b), "r"(c), "r"(d)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar, uint32_t cta, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .b32 rbar;\n"
      "mapa.shared::cluster.u32 rbar, %0, %1;\n"
      "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [rbar], %2;\n"
      "}" ::"r"(bar),
      "r"(cta), "r"(bytes)
      : "memory");
}

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t value) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(value) : "memory");
}

__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// Bounded wait: a protocol bug must surface as a trapped launch (cudaErrorLaunchFailure), never as a
// hung GPU.  The bound is wall-clock (%globaltimer), checked every 4096 failed probes, so the fast
// path is the bare try_wait loop.
#ifndef CLEAR_MBAR_TIMEOUT_NS
#define CLEAR_MBAR_TIMEOUT_NS 4000000000ull
#endif
template <bool kClusterAcquire = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  auto probe = [&]() { return kClusterAcquire ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity); };
  if (probe()) return;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!probe()) {
    if ((++spins & 4095u) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > CLEAR_MBAR_TIMEOUT_NS) {
        printf("[clear] mbarrier timeout: block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
               (int)threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

template <int kCtaGroup>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1) {
  if constexpr (kCtaGroup == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
  }
}

template <int kCtaGroup>
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
  if constexpr (kCtaGroup == 1) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  }
}

__device__ __forceinline__ void tma_store_2d(const void* desc, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_reduce_add_2d(const void* desc, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// cluster
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05
// ------------------------------------------------------------------------------------------------
template <int kCtaGroup>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}

template <int kCtaGroup>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  } else {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  }
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16/fp16 inputs, fp32 accumulate.  One thread issues.
template <int kCtaGroup>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  if constexpr (kCtaGroup == 1) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// Arrive (count 1) on `bar` once every tcgen05.mma previously issued by this thread has completed.
// In the 2-CTA form the arrive is multicast to the barrier at the same offset in both CTAs of the pair.
template <int kCtaGroup>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(static_cast<uint16_t>(3))
        : "memory");
  }
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (quadrant*32 + t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace ptx
}  // namespace clear
