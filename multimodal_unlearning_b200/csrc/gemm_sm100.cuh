// Persistent warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (fp32 accumulate).
//
//   * operands staged by TMA into 128B-swizzled shared memory, K-major or MN-major (the backward
//     projections dH = dZ*W and dW = dZ^T*H read W / dZ / H "transposed" straight from their row-major
//     storage — no transposed copies are ever made);
//   * tcgen05.mma (cta_group::1: 128x256x16 per CTA; cta_group::2: 256x256x16 per CTA pair) issued by
//     one thread, fp32 accumulators in TMEM, double-buffered (2 x 256 columns) so the epilogue of tile
//     i overlaps the main loop of tile i+1;
//   * work items (output tile x K split) are CLAIMED, not assigned: the TMA thread of each CTA pair's leader takes the
//     next item from a global counter one item ahead of its own pipeline, DECODES it (problem, tile coordinates, K-block
//     range: the integer divisions run once per item, in the thread with the most slack) and publishes the decoded item
//     through an 8-slot shared-memory ring to the MMA thread, the second TMA thread, the epilogue warps and the peer CTA —
//     one st.async per CTA (16 bytes + complete_tx on that CTA's mbarrier: no fence anywhere), a few K blocks into the
//     current item, so every reader knows its next item a whole tile early and the MMA-issuing thread's tile boundary
//     is a 16-byte shared load.  Consumers hand slots and TMEM stages back with fence-free remote arrives.
//     A CTA pair that the hardware launches late — its SMs were held by an NCCL kernel of the overlapped gradient
//     all-reduce — simply finds less work left, instead of holding its statically assigned tiles hostage
//     (measured at 2 GPUs with a static schedule: +0.5 ms per 6.5 ms step, +1.9 ms with NCCL limited to 4 CTAs);
//   * operand A and operand B are loaded by TWO TMA threads (warps 0 and 3): with MN-major operands a K block takes four
//     bulk-tensor instructions per CTA, and one thread issuing all of them was busy 80 % of the K-block time;
//     ncu, N = 8192 / H = 4096 / V = 32064, before -> after (fence-free hand-backs, decoded ring, second TMA thread):
//     tensor pipe active 91.7 -> 99.3 % (projection pair), 98.1 -> 99.1 % (dH), 94.3 -> 98.4 % (dW);
//   * epilogue warps read TMEM with tcgen05.ld, optionally fold an online log-sum-exp over the
//     (bf16-rounded) logits of their rows, convert, stage through swizzled smem and TMA-store
//     (or TMA reduce-add for fp32 accumulation into an existing dW).
//
// This is the contraction behind reference call sites A1/K1/K9: lm_head forward
// inside `model(**inputs)` (mm/trainer.py:180-189 etc.) and its autograd backward.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace clear {

enum OperandMajor : int { kMajorK = 0, kMajorMN = 1, kMajorRuntime = 2 /* A only: GemmProblem::a_major */ };
enum EpilogueKind : int {
  kEpiStoreBF16 = 0,     // D bf16
  kEpiStoreBF16Lse = 1,  // D bf16 + per-(row, n-tile) (max, sum-exp) partials
  kEpiStoreF32 = 2,      // D fp32 (store or reduce-add)
};

// One contraction of a launch.  A launch runs one problem, or TWO independent problems that share the operand majors of B
// and the epilogue kind (student + teacher projection; dH + dW of one block of rows): their work items form one list, the
// first problem's (longer) tiles first, so that the CTA pairs that finish the first problem early start on the second
// instead of idling until a kernel boundary — one pipeline ramp and one tail per pair of contractions instead of two.
struct GemmProblem {
  int M, N, K;
  int m_tiles, n_tiles;  // in units of (128*cta_group) x 256
  int raster_m_fast;     // 1: consecutive tiles walk M first (A small / re-read from L2), 0: N first
  int raster_group;      // tiles of the fast dimension per chunk: a chunk is swept across the whole slow dimension before
                         // the next chunk starts, so that the chunk's operand stays in L2 (see pick_raster in gemm_host.cu)
  int reduce_add;        // accumulate into D instead of overwriting (TMA reduce-add; bf16 D rounds after each add)
  int a_major;           // operand A's major when the kernel is instantiated with kMajorRuntime
  float2* lse_part;      // [M][n_tiles] when kEpiStoreBF16Lse
  int n_peers;           // > 0: the output is sharded by M tiles over the peers' buffers (fused reduce-scatter)
  int mtiles_per_peer;   // owner of M tile mt = min(n_peers - 1, mt / mtiles_per_peer)
  int dyn_dim;           // 0: extents as given; 1: M, 2: K comes from *dyn_count (see GemmParams)
};

struct GemmParams {
  GemmProblem q[2];
  int n_problems;
  int k_splits;          // > 1 (single problem only): each output tile is cut into k_splits work items along K (fp32 reduce-add epilogue)
  int kb_per_split;      // K blocks (of 64) per work item
  unsigned int* work_counter;  // [0] next unclaimed work item (dynamic tile scheduler), [1] CTA groups that have finished;
                               // both zero at launch: the last group to finish resets them, so no memset node sits between
                               // two launches on a stream
  // Device-side extent (row compaction without a host round trip): *dyn_count - dyn_offset, clamped to [0, M or K], replaces
  // M (dyn_dim = 1: projection / dH over the valid rows only) or K rounded up to a K block (dyn_dim = 2: dW; the operand
  // rows of the last block beyond the count must be zero).  The tensor maps, the grid and the tile order are those of the
  // host-side maximum.  dyn_splits (single problem, dyn_dim = 1): the number of K splits is chosen in the kernel from the actual M.
  const int* dyn_count;
  int dyn_offset;
  int dyn_splits;              // 1: choose k_splits in-kernel (fp32 reduce-add epilogue, D pre-zeroed); 0: k_splits as given
  int sm_groups;               // CTA groups of the launch (in-kernel split choice)
};

// Output tensor maps of the peers' buffers (symmetric memory mapped over NVLink / NVSwitch): with n_peers > 0 the epilogue
// reduce-adds every output tile into the buffer of the rank that OWNS its rows, so that the gradient reduce-scatter of
// data-parallel training happens inside the contraction, tile by tile, instead of in a collective that follows it.
constexpr int kMaxPeers = 8;
struct PeerMaps {
  CUtensorMap tm[kMaxPeers];
};

template <int kCtaGroup>
struct GemmShape {
  static constexpr int BM = 128;                    // D rows per CTA
  static constexpr int BN = 256;                    // D cols per tile (UMMA N)
  static constexpr int BK = 64;                     // 128 bytes of bf16: one swizzle span
  static constexpr int UMMA_K = 16;
  static constexpr int B_ROWS = BN / kCtaGroup;     // B rows (N extent) each CTA stages
  static constexpr int A_BYTES = BM * BK * 2;       // 16 KB
  static constexpr int B_BYTES = B_ROWS * BK * 2;   // 32 KB / 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (kCtaGroup == 2) ? 6 : 4;
  static constexpr int EPI_WARP_BYTES = 2 * 4096;   // two 32-row x 128 B staging buffers per warp
  static constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;
  static constexpr int SCHED_SLOTS = 8;             // ring of claimed, decoded work items shared by every role of the CTA pair
  static constexpr int BAR_BYTES = (2 * STAGES + 4 + 2 * SCHED_SLOTS) * 8 + 16 + SCHED_SLOTS * 16;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static constexpr int THREADS = 256;               // warps 0..3: TMA (operand A, scheduler) / MMA / TMEM-alloc / TMA (operand B); 4..7 epilogue
  static constexpr int TMEM_COLS = 512;             // 2 accumulator stages x 256 fp32 columns
};

// 64-bit shared-memory matrix descriptor (sm_100 "version 1"), 128B swizzle.
//   K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: atoms of 64 MN-elements x 8 k-rows (1024 B); SBO = next 8 k-rows (1024 B),
//             LBO = next 64 MN-elements (BK * 128 B = 8192 B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

// 32-bit instruction descriptor for kind::f16: bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_major, int b_major) {
  return (1u << 4)                       // D format fp32
         | (1u << 7)                     // A format bf16
         | (1u << 10)                    // B format bf16
         | (uint32_t(a_major) << 15)     // 0 K-major, 1 MN-major
         | (uint32_t(b_major) << 16)
         | (uint32_t(n >> 3) << 17)
         | (uint32_t(m >> 4) << 24);
}

constexpr int kHandOnAfter = 8;   // K blocks of an item after which the leader's TMA thread decodes and publishes the next one

template <int kCtaGroup, int kMajorA, int kMajorB, int kEpi>
__global__ void __launch_bounds__(GemmShape<kCtaGroup>::THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmD1,
                 const __grid_constant__ PeerMaps peers, const GemmParams p) {
  using S = GemmShape<kCtaGroup>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (smem_raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_raw_addr);

  const uint32_t epi_addr = smem_base + S::STAGES * S::STAGE_BYTES;
  const uint32_t bar_addr = epi_addr + S::EPI_BYTES;
  const uint32_t full_bar = bar_addr;                        // [STAGES]
  const uint32_t empty_bar = bar_addr + S::STAGES * 8;       // [STAGES]
  const uint32_t tfull_bar = bar_addr + 2 * S::STAGES * 8;   // [2]
  const uint32_t tempty_bar = tfull_bar + 16;                // [2]
  const uint32_t tmem_slot = tempty_bar + 16;
  const uint32_t sfull_bar = tmem_slot + 16;                       // [SCHED_SLOTS] item published (one per CTA)
  const uint32_t sempty_bar = sfull_bar + S::SCHED_SLOTS * 8;      // [SCHED_SLOTS] every reader is done (leader CTA's copy is used)
  const uint32_t sched_item = sempty_bar + S::SCHED_SLOTS * 8;     // [SCHED_SLOTS] uint4: work item, M tile, N tile, K-block range | problem
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCtaGroup == 2) ? ptx::cluster_ctarank() : 0u;
  const bool two = p.n_problems > 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    ptx::prefetch_tensormap(&tmD);
    if (two) {
      ptx::prefetch_tensormap(&tmA1);
      ptx::prefetch_tensormap(&tmB1);
      ptx::prefetch_tensormap(&tmD1);
    }
    const int np = max(p.q[0].n_peers, p.q[1].n_peers);
    for (int r = 0; r < np; ++r) ptx::prefetch_tensormap(&peers.tm[r]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      ptx::mbar_init(full_bar + s * 8, 1);
      ptx::mbar_init(empty_bar + s * 8, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar + a * 8, 1);
      ptx::mbar_init(tempty_bar + a * 8, 4 * kCtaGroup);  // one arrive per epilogue warp of each CTA
    }
    for (int i = 0; i < S::SCHED_SLOTS; ++i) {
      ptx::mbar_init(sfull_bar + i * 8, 1);
      // readers of a slot: operand B's TMA thread + 4 epilogue warps + (leader CTA: the MMA thread; peer CTA: operand A's TMA thread)
      ptx::mbar_init(sempty_bar + i * 8, 6 * kCtaGroup);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<kCtaGroup>(tmem_slot, S::TMEM_COLS);
  }
  ptx::tcgen05_fence_before();
  if constexpr (kCtaGroup == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // Extents of the (up to two) problems; device-side ones (see GemmParams::dyn_count) are resolved here: every thread
  // derives the same values.
  int M0 = p.q[0].M, K0 = p.q[0].K, mt0 = p.q[0].m_tiles;
  int M1 = p.q[1].M, K1 = p.q[1].K, mt1 = p.q[1].m_tiles;
  int k_splits = p.k_splits, kb_per_split = p.kb_per_split;
  if (p.dyn_count != nullptr) {
    const int c = max(*p.dyn_count - p.dyn_offset, 0);
    auto resolve = [&](const GemmProblem& q, int& M, int& K, int& mt) {
      if (q.dyn_dim == 1) {
        M = min(c, q.M);
        mt = (M + S::BM * kCtaGroup - 1) / (S::BM * kCtaGroup);
      } else if (q.dyn_dim == 2) {
        K = min((c + S::BK - 1) / S::BK * S::BK, (q.K + S::BK - 1) / S::BK * S::BK);
      }
    };
    resolve(p.q[0], M0, K0, mt0);
    if (two) resolve(p.q[1], M1, K1, mt1);
    if (p.dyn_splits && p.q[0].dyn_dim == 1) {  // same rule as gemm_pick_k_splits (gemm_host.cu), on the actual number of output tiles
      const int tiles = mt0 * p.q[0].n_tiles, num_kb = (p.q[0].K + S::BK - 1) / S::BK;
      int best = 1;
      float best_eff = 0.f;
      for (int sp = 1; sp <= 4 && tiles > 0; ++sp) {
        if (sp > 1 && num_kb / sp < 64) break;
        const float waves = (float)tiles * sp / (float)p.sm_groups;
        const float eff = waves / ceilf(waves - 1e-6f);
        if (eff > best_eff + 0.03f) { best_eff = eff; best = sp; }
      }
      kb_per_split = (num_kb + best - 1) / best;
      k_splits = (num_kb + kb_per_split - 1) / kb_per_split;
    } else if (p.q[0].dyn_dim == 2) {
      kb_per_split = (K0 + S::BK - 1) / S::BK;  // a device-side K excludes split-K
      k_splits = 1;
    }
  }
  // Work item w: problem 0's (output tile, K split) pairs — consecutive items are the splits of one tile — then problem 1's tiles.
  const int base1 = mt0 * p.q[0].n_tiles * k_splits;
  const int num_tiles = base1 + (two ? mt1 * p.q[1].n_tiles : 0);
  const int nkb0 = (K0 + S::BK - 1) / S::BK;   // 0 (no valid row at all): the epilogue emits zeros
  const int nkb1 = (K1 + S::BK - 1) / S::BK;

  // Work item -> problem, output tile, K-block range.  Chunks of `raster_group` tiles along the fast dimension, each swept
  // across the slow one.  Only the leader's TMA thread decodes; everybody else reads the result from the ring.
  struct Item {
    int w, second, mt, nt, kb0, kb1;
  };
  auto decode = [&](int w) {
    Item it;
    it.w = w;
    it.second = (two && w >= base1) ? 1 : 0;
    const GemmProblem& q = p.q[it.second];
    const int m_tiles = it.second ? mt1 : mt0;
    int t;
    if (it.second) {
      t = w - base1;
      it.kb0 = 0;
      it.kb1 = nkb1;
    } else {
      t = w / k_splits;
      it.kb0 = (w % k_splits) * kb_per_split;
      it.kb1 = min(nkb0, it.kb0 + kb_per_split);
    }
    const int fast_tiles = q.raster_m_fast ? m_tiles : q.n_tiles;
    const int slow_tiles = q.raster_m_fast ? q.n_tiles : m_tiles;
    const int per_chunk = max(q.raster_group * slow_tiles, 1);
    const int c = t / per_chunk;
    const int r = t - c * per_chunk;
    const int g = max(min(q.raster_group, fast_tiles - c * q.raster_group), 1);
    const int slow = r / g;
    const int fast = c * q.raster_group + (r - slow * g);
    it.mt = q.raster_m_fast ? fast : slow;
    it.nt = q.raster_m_fast ? slow : fast;
    return it;
  };
  // ring slot <-> Item: {w, mt, nt, kb0 | kb1 << 15 | second << 30}  (K blocks < 32768: K < 2 M)
  auto unpack_item = [&](const uint4 v) {
    Item it;
    it.w = (int)v.x; it.mt = (int)v.y; it.nt = (int)v.z;
    it.kb0 = (int)(v.w & 0x7fffu); it.kb1 = (int)((v.w >> 15) & 0x7fffu); it.second = (int)(v.w >> 30);
    return it;
  };

  // Reader side of the scheduler ring.  Every reader consumes every slot in order, the end-of-work item (w >= num_tiles)
  // included.  Called by one thread; the epilogue warps broadcast the result with shuffles.
  auto sched_read = [&](uint32_t& slot, uint32_t& sphase) -> Item {
    ptx::mbar_wait(sfull_bar + slot * 8, sphase);
    const uint4 v = ptx::ld_shared_v4(sched_item + slot * 16);
    if constexpr (kCtaGroup == 2) ptx::mbar_arrive_cluster_noflush(sempty_bar + slot * 8, 0);   // a read hand-back: no fence
    else                          ptx::mbar_arrive(sempty_bar + slot * 8);
    if (++slot == S::SCHED_SLOTS) { slot = 0; sphase ^= 1u; }
    return unpack_item(v);
  };

  if (warp == 0) {
    // ===================================== TMA producer ==========================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, slot = 0, sphase = 0;
      // Leader: claim an item, decode it, hand it to both CTAs of the pair.
      auto publish = [&](const Item& it) {
        ptx::mbar_wait(sempty_bar + slot * 8, sphase ^ 1u);
        const uint32_t packed = (uint32_t)it.kb0 | ((uint32_t)it.kb1 << 15) | ((uint32_t)it.second << 30);
#pragma unroll
        for (uint32_t c = 0; c < (uint32_t)kCtaGroup; ++c) {
          if constexpr (kCtaGroup == 2) {
            ptx::mbar_arrive_expect_tx_cluster(sfull_bar + slot * 8, c, 16);
            ptx::st_async_v4(sched_item + slot * 16, sfull_bar + slot * 8, c, (uint32_t)it.w, (uint32_t)it.mt, (uint32_t)it.nt, packed);
          } else {
            ptx::st_shared_v4(sched_item + slot * 16, (uint32_t)it.w, (uint32_t)it.mt, (uint32_t)it.nt, packed);
            ptx::mbar_arrive(sfull_bar + slot * 8);
          }
        }
        if (++slot == S::SCHED_SLOTS) { slot = 0; sphase ^= 1u; }
      };
      const bool leader = (cta_rank == 0);
      Item it;
      if (leader) {
        it = decode((int)atomicAdd(p.work_counter, 1u));
        publish(it);
      } else {
        it = sched_read(slot, sphase);
      }
      while (it.w < num_tiles) {
        // The claim for the NEXT item is issued before this item's loads and handed on kHandOnAfter K blocks in, when the
        // counter's reply has long arrived and the pipeline is full again: the round trip to L2 and the decoding hide behind
        // the K loop, and every reader has its next item long before it reaches its own tile boundary.  (Measured, ncu, N =
        // 8192: handing on after 2 / 8 K blocks: projection pair 98.9 / 99.3 % tensor-pipe active.)
        unsigned int next_claim = 0;
        if (leader) next_claim = atomicAdd(p.work_counter, 1u);
        const CUtensorMap* mA = it.second ? &tmA1 : &tmA;
        const bool a_mn = (kMajorA == kMajorRuntime) ? (p.q[it.second].a_major == kMajorMN) : (kMajorA == kMajorMN);
        const int m0 = (it.mt * kCtaGroup + (int)cta_rank) * S::BM;
        // Operand A of one K block (operand B comes from the second TMA thread, warp 3: with MN-major operands a K block
        // takes four bulk-tensor instructions per CTA, and one thread issuing all of them ran at 80 % of the K-block time —
        // any hiccup starved the tensor pipe).  K-major: one box [rows][64 k].  MN-major: one box [64 k][64 mn] per
        // 64-element MN chunk, chunks 8 KB apart (= the descriptor's LBO).
        auto load_a = [&](int kb) {
          ptx::mbar_wait(empty_bar + stage * 8, phase ^ 1u);
          const uint32_t fb = full_bar + stage * 8;
          if (cta_rank == 0) ptx::mbar_arrive_expect_tx(fb, S::STAGE_BYTES * kCtaGroup);
          const uint32_t a_dst = smem_base + stage * S::STAGE_BYTES;
          const int k0 = kb * S::BK;
          if (!a_mn) {
            ptx::tma_load_2d<kCtaGroup>(a_dst, mA, fb, k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < S::BM / 64; ++c)
              ptx::tma_load_2d<kCtaGroup>(a_dst + c * (S::BK * 128), mA, fb, m0 + c * 64, k0);
          }
          if (++stage == S::STAGES) { stage = 0; phase ^= 1u; }
        };
        const int kb_hand_on = min(it.kb0 + kHandOnAfter, it.kb1);
        for (int kb = it.kb0; kb < kb_hand_on; ++kb) load_a(kb);
        if (leader) {
          const Item nxt = decode((int)next_claim);
          publish(nxt);
          for (int kb = kb_hand_on; kb < it.kb1; ++kb) load_a(kb);
          it = nxt;
        } else {
          for (int kb = kb_hand_on; kb < it.kb1; ++kb) load_a(kb);
          it = sched_read(slot, sphase);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    if (lane == 0 && cta_rank == 0) {
      // K-major: +32 B per UMMA_K inside the swizzle span; MN-major: +16 k-rows * 128 B.
      constexpr uint32_t b_adv = (kMajorB == kMajorK) ? (S::UMMA_K * 2) : (S::UMMA_K * 128);
      constexpr uint32_t b_lbo = (kMajorB == kMajorK) ? 16u : (S::BK * 128);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, slot = 0, sphase = 0;
      for (Item it = sched_read(slot, sphase); it.w < num_tiles; it = sched_read(slot, sphase)) {
        const bool a_mn = (kMajorA == kMajorRuntime) ? (p.q[it.second].a_major == kMajorMN) : (kMajorA == kMajorMN);
        const uint32_t idesc = make_idesc(S::BM * kCtaGroup, S::BN, a_mn ? 1 : 0, kMajorB);
        const uint32_t a_adv = a_mn ? (S::UMMA_K * 128) : (S::UMMA_K * 2);
        const uint32_t a_lbo = a_mn ? (uint32_t)(S::BK * 128) : 16u;
        ptx::mbar_wait(tempty_bar + acc * 8, acc_phase ^ 1u);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * S::BN;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          ptx::mbar_wait(full_bar + stage * 8, phase);
          ptx::tcgen05_fence_after();
          const uint32_t a_src = smem_base + stage * S::STAGE_BYTES;
          const uint32_t b_src = a_src + S::A_BYTES;
          const uint64_t adesc = make_smem_desc(a_src, a_lbo, 1024);
          const uint64_t bdesc = make_smem_desc(b_src, b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < S::BK / S::UMMA_K; ++k) {
            ptx::umma_f16<kCtaGroup>(d_tmem, adesc + ((k * a_adv) >> 4), bdesc + ((k * b_adv) >> 4), idesc,
                                     ((kb - it.kb0) | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit<kCtaGroup>(empty_bar + stage * 8);  // frees the smem slot once the MMAs retire
          if (++stage == S::STAGES) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit<kCtaGroup>(tfull_bar + acc * 8);      // accumulator complete -> epilogue
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp == 3) {
    // ===================================== TMA producer, operand B =================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, slot = 0, sphase = 0;
      for (Item it = sched_read(slot, sphase); it.w < num_tiles; it = sched_read(slot, sphase)) {
        const CUtensorMap* mB = it.second ? &tmB1 : &tmB;
        const int n0 = it.nt * S::BN + (int)cta_rank * S::B_ROWS;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          // the stage's barrier is armed by operand A's thread of the leader CTA (expect_tx covers both operands of both
          // CTAs); bytes that land before that leave the transaction count negative for a moment, which mbarriers allow
          ptx::mbar_wait(empty_bar + stage * 8, phase ^ 1u);
          const uint32_t fb = full_bar + stage * 8;
          const uint32_t b_dst = smem_base + stage * S::STAGE_BYTES + S::A_BYTES;
          const int k0 = kb * S::BK;
          if constexpr (kMajorB == kMajorK) {
            ptx::tma_load_2d<kCtaGroup>(b_dst, mB, fb, k0, n0);
          } else {
#pragma unroll
            for (int c = 0; c < S::B_ROWS / 64; ++c)
              ptx::tma_load_2d<kCtaGroup>(b_dst + c * (S::BK * 128), mB, fb, n0 + c * 64, k0);
          }
          if (++stage == S::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ==============================================
    const uint32_t q4 = warp & 3u;  // TMEM lane quadrant this warp may read
    const uint32_t stg = epi_addr + q4 * S::EPI_WARP_BYTES;
    const uint32_t sw = (lane & 7u);
    uint32_t nstores = 0, acc = 0, acc_phase = 0, slot = 0, sphase = 0;
    while (true) {
      Item it = {};
      if (lane == 0) it = sched_read(slot, sphase);
      it.w = __shfl_sync(0xffffffffu, it.w, 0);
      if (it.w >= num_tiles) break;
      it.mt = __shfl_sync(0xffffffffu, it.mt, 0);
      it.nt = __shfl_sync(0xffffffffu, it.nt, 0);
      it.kb0 = __shfl_sync(0xffffffffu, it.kb0, 0);
      it.kb1 = __shfl_sync(0xffffffffu, it.kb1, 0);
      it.second = __shfl_sync(0xffffffffu, it.second, 0);
      const GemmProblem& q = p.q[it.second];
      const int dynM = it.second ? M1 : M0;
      const bool no_k = it.kb1 <= it.kb0;   // no MMA ran: the accumulator is not defined, the tile is all zeros
      const int mt = it.mt, nt = it.nt;
      const int row0 = (mt * kCtaGroup + (int)cta_rank) * S::BM + (int)q4 * 32;
      const int col0 = nt * S::BN;
      // destination: this rank's D, or the buffer of the rank that owns these rows (always reduce-add there)
      const CUtensorMap* tmOut = it.second ? &tmD1 : &tmD;
      if (q.n_peers > 0) tmOut = &peers.tm[min(q.n_peers - 1, mt / q.mtiles_per_peer)];
      const bool reduce_out = q.reduce_add || q.n_peers > 0;
      ptx::mbar_wait(tfull_bar + acc * 8, acc_phase);
      ptx::tcgen05_fence_after();
      const uint32_t t_base = tmem_base + ((q4 * 32u) << 16) + acc * S::BN;

      if constexpr (kEpi == kEpiStoreF32) {
#pragma unroll 1
        for (int c = 0; c < S::BN / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_base + c * 32, v);
          const uint32_t buf = stg + (nstores & 1u) * 4096u;
          if (lane == 0) ptx::tma_store_wait_read<1>();
          __syncwarp();
          ptx::tmem_ld_wait();
          if (no_k) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          const uint32_t rowp = buf + lane * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ptx::st_shared_v4(rowp + (((uint32_t)j ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < dynM && col0 + c * 32 < q.N) {
            if (reduce_out) ptx::tma_reduce_add_2d(tmOut, buf, col0 + c * 32, row0);
            else            ptx::tma_store_2d(tmOut, buf, col0 + c * 32, row0);
          }
          if (lane == 0) ptx::tma_store_commit();
          ++nstores;
        }
      } else {
        constexpr float kLog2e = 1.4426950408889634f;
        float run_max = -INFINITY, run_sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < S::BN / 64; ++c) {
          uint32_t v0[32], v1[32];
          ptx::tmem_ld_32x32b_x32(t_base + c * 64, v0);
          ptx::tmem_ld_32x32b_x32(t_base + c * 64 + 32, v1);
          const uint32_t buf = stg + (nstores & 1u) * 4096u;
          if (lane == 0) ptx::tma_store_wait_read<1>();
          __syncwarp();
          ptx::tmem_ld_wait();
          if (no_k) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { v0[j] = 0u; v1[j] = 0u; }
          }
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
            __nv_bfloat162 g = __floats2bfloat162_rn(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
            pk[16 + j] = *reinterpret_cast<uint32_t*>(&g);
          }
          if constexpr (kEpi == kEpiStoreBF16Lse) {
            // Online log-sum-exp over the values exactly as they are stored (bf16-rounded), so that
            // the row pass that later reads the stored logits sees probabilities that sum to one.
            const int cbase = col0 + c * 64;
            float x[64];
            float cmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&pk[j]));
              x[2 * j] = (cbase + 2 * j < q.N) ? f.x : -INFINITY;
              x[2 * j + 1] = (cbase + 2 * j + 1 < q.N) ? f.y : -INFINITY;
              cmax = fmaxf(cmax, fmaxf(x[2 * j], x[2 * j + 1]));
            }
            const float new_max = fmaxf(run_max, cmax);
            const float nm2 = new_max * kLog2e;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) s += ptx::ex2_approx(fmaf(x[j], kLog2e, -nm2));
            run_sum = run_sum * ptx::ex2_approx(fmaf(run_max, kLog2e, -nm2)) + s;
            run_max = new_max;
          }
          const uint32_t rowp = buf + lane * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ptx::st_shared_v4(rowp + (((uint32_t)j ^ sw) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < dynM && col0 + c * 64 < q.N) {
            if (reduce_out) ptx::tma_reduce_add_2d(tmOut, buf, col0 + c * 64, row0);  // bf16 += (rounded in L2)
            else            ptx::tma_store_2d(tmOut, buf, col0 + c * 64, row0);
          }
          if (lane == 0) ptx::tma_store_commit();
          ++nstores;
        }
        if constexpr (kEpi == kEpiStoreBF16Lse) {
          const int row = row0 + (int)lane;
          if (row < dynM) q.lse_part[(size_t)row * q.n_tiles + nt] = make_float2(run_max, run_sum);
        }
      }
      // every tcgen05.ld of this accumulator stage has completed (wait::ld above): hand it back
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtaGroup == 2) ptx::mbar_arrive_cluster_noflush(tempty_bar + acc * 8, 0);
        else                          ptx::mbar_arrive(tempty_bar + acc * 8);
      }
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) ptx::tma_store_wait<0>();
  }

  // The producer of the leader CTA has claimed its last item long ago; once every group has reported here, nobody touches
  // the counters any more and the last one re-arms them for the launch that will reuse this slot of the ring.
  if (warp == 0 && lane == 0 && cta_rank == 0) {
    const unsigned int groups = gridDim.x / kCtaGroup;
    if (atomicAdd(p.work_counter + 1, 1u) == groups - 1) {
      p.work_counter[0] = 0;
      p.work_counter[1] = 0;
      __threadfence();
    }
  }

  // ======================================= teardown ==============================================
  __syncwarp();
  ptx::tcgen05_fence_before();
  if constexpr (kCtaGroup == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<kCtaGroup>(tmem_base, S::TMEM_COLS);
  }
}

}  // namespace clear
