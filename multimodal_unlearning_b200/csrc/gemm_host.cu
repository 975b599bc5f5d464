// Host side of the tcgen05 GEMM: TMA tensor-map encoding (driver entry point fetched through the
// runtime, so libcuda is never linked), kernel-variant dispatch, persistent-grid launch.
#include <cudaTypedefs.h>

#include <atomic>
#include <mutex>

#include "common.h"
#include "gemm_sm100.cuh"

namespace clear {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return static_cast<int>(e);
}

const char* last_error() { return g_err; }

constexpr int kMaxDevices = 64;

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

// cached per device ordinal: one process may drive several GPUs (device_map="auto" teachers, multi-GPU tests)
int device_sm_count() {
  static std::atomic<int> cached[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0) return -1;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached[dev].store(n, std::memory_order_relaxed);
  return n;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// Operand tile map.  rows_box = MN extent staged per CTA (128 or 256).
static int make_operand_map(CUtensorMap* tm, const GemmOperand& op, int mn_extent, int k_extent, int rows_box) {
  auto enc = encode_fn();
  if (!enc) return fail(CLEAR_ERR_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  CUresult r;
  if (op.major == kMajorK) {
    cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)mn_extent};
    cuuint64_t strides[1] = {(cuuint64_t)op.ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)rows_box};
    cuuint32_t estr[2] = {1, 1};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op.ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // storage [K, MN] row-major; one box = 64 k-rows x 64 mn-elements (128 B rows) = one column of
    // canonical MN-major SWIZZLE_128B atoms.
    cuuint64_t dims[2] = {(cuuint64_t)mn_extent, (cuuint64_t)k_extent};
    cuuint64_t strides[1] = {(cuuint64_t)op.ld * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op.ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return fail(CLEAR_ERR_SHAPE, "cuTensorMapEncodeTiled(operand, major=%d) failed: CUresult %d", op.major, (int)r);
  return 0;
}

static int make_output_map(CUtensorMap* tm, void* d, int d_is_f32, int64_t ldd, int m, int n) {
  auto enc = encode_fn();
  if (!enc) return fail(CLEAR_ERR_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  const int esz = d_is_f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m};
  cuuint64_t strides[1] = {(cuuint64_t)ldd * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, d_is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CLEAR_ERR_SHAPE, "cuTensorMapEncodeTiled(output) failed: CUresult %d", (int)r);
  return 0;
}

// Work counters of the dynamic tile scheduler: a ring of (claim, finished) word pairs, one pair per launch in flight
// (launches on different streams may overlap), allocated and zeroed once per device on first use — the library's only
// device allocation.  A kernel's last CTA group re-arms its pair, so consecutive launches need no memset between them.
static int next_work_counter(unsigned int** out) {
  constexpr int kRing = 4096;
  static unsigned int* ring[kMaxDevices] = {};
  static unsigned int cursor[kMaxDevices] = {};
  static std::mutex mu;
  int dev = 0;
  CLEAR_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(CLEAR_ERR_DEVICE, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!ring[dev]) {
    unsigned int* p = nullptr;
    CLEAR_CUDA_OK(cudaMalloc(&p, kRing * 2 * sizeof(unsigned int)));
    CLEAR_CUDA_OK(cudaMemset(p, 0, kRing * 2 * sizeof(unsigned int)));
    ring[dev] = p;
  }
  *out = ring[dev] + 2 * (cursor[dev]++ % kRing);
  return 0;
}

template <int kCtaGroup, int kMajorA, int kMajorB, int kEpi>
static int launch_variant(const CUtensorMap (&tm)[6], const PeerMaps& peers, const GemmParams& p, cudaStream_t stream) {
  using S = GemmShape<kCtaGroup>;
  auto kern = gemm_bf16_kernel<kCtaGroup, kMajorA, kMajorB, kEpi>;
  static std::atomic<bool> configured[kMaxDevices] = {};  // per template instantiation and device (the attribute is per device)
  const int dev = current_device();
  if (dev < 0) return fail(CLEAR_ERR_DEVICE, "cannot query the current device");
  if (!configured[dev].load(std::memory_order_acquire)) {
    CLEAR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    configured[dev].store(true, std::memory_order_release);
  }
  const int sms = device_sm_count();
  if (sms <= 0) return fail(CLEAR_ERR_DEVICE, "cannot query SM count");
  long long tiles = (long long)p.q[0].m_tiles * p.q[0].n_tiles * (p.dyn_splits ? 4 : p.k_splits);
  if (p.n_problems > 1) tiles += (long long)p.q[1].m_tiles * p.q[1].n_tiles;
  int groups = sms / kCtaGroup;
  if (groups > tiles) groups = (int)tiles;
  GemmParams pl = p;
  pl.sm_groups = groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * kCtaGroup));
  cfg.blockDim = dim3(S::THREADS);
  cfg.dynamicSmemBytes = S::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtaGroup;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CLEAR_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], peers, pl));
  return 0;
}

int gemm_pick_k_splits(int cta_group, int m, int n, int k) {
  if (cta_group == 0) cta_group = CLEAR_DEFAULT_CTA_GROUP;
  const int sms = device_sm_count();
  if (sms <= 0) return 1;
  const int groups = sms / cta_group;
  const int tiles = ((m + 128 * cta_group - 1) / (128 * cta_group)) * ((n + 255) / 256);
  const int num_kb = (k + 63) / 64;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 4; ++s) {
    if (s > 1 && num_kb / s < 64) break;  // keep >= 4096 of K per item: the fp32 epilogue must stay hidden
    const double waves = (double)tiles * s / groups;
    const double eff = waves / (double)(long long)(waves + 0.999999);
    if (eff > best_eff + 0.03) { best_eff = eff; best = s; }  // a split must buy at least 3 % to be worth the fp32 pass
  }
  return best;
}

// Tile order.  The CTA pairs work on ~`groups` consecutive tiles of the linear order at any time.  Estimated DRAM reads of
// an order (fast dimension, chunk of G fast tiles swept across the slow dimension):
//   * the chunk's operand (G tiles x bytes per tile) fits in L2 next to the streams: it is read once, the other operand
//     once per chunk;
//   * it does not fit (long K): every wave of `groups` tiles reads its G fast-side and groups / G slow-side operand tiles.
// Measured on B200 with whole-dimension chunks (ncu, N = 8192 rows): projection 2.0 GB, dH 3.2 GB, dW 2.4 GB read per
// launch against 0.33 / 0.79 / 0.59 GB algorithmic — the 67 MB `hidden` operand does not survive in the two 63 MB L2
// halves next to a 525 MB output stream.
static void pick_raster(int cta_group, int m, int n, int k, int* m_fast, int* group) {
  const int sms = device_sm_count();
  const double groups = (sms > 0 ? sms : 148) / (double)cta_group;
  const int m_tiles = (m + 128 * cta_group - 1) / (128 * cta_group), n_tiles = (n + 255) / 256;
  const double a_tile = 128.0 * cta_group * k * 2, b_tile = 256.0 * k * 2;
  const double l2_budget = 40e6;
  double best = 1e300;
  for (int mf = 0; mf < 2; ++mf) {
    const int fast = mf ? m_tiles : n_tiles, slow = mf ? n_tiles : m_tiles;
    const double f = mf ? a_tile : b_tile, s = mf ? b_tile : a_tile;
    for (int g = 1; g <= fast; g = (g * 2 < fast || g == fast) ? g * 2 : fast) {
      double cost;
      if (g * f <= l2_budget) {
        cost = fast * f + (double)((fast + g - 1) / g) * slow * s;
      } else {
        const double per_wave = g * f + (groups / g > 1.0 ? groups / g : 1.0) * s;
        cost = per_wave * ((double)fast * slow / groups);
      }
      if (cost < best * 0.97) { best = cost; *m_fast = mf; *group = g; }
      if (g == fast) break;
    }
  }
}

static int peer_mtiles(int cta_group, int m, int n_peers) {
  const int m_tiles = (m + 128 * cta_group - 1) / (128 * cta_group);
  return (m_tiles + n_peers - 1) / n_peers;
}

void gemm_peer_rows(int cta_group, int m, int n_peers, int rank, int* row0, int* row1) {
  if (cta_group == 0) cta_group = CLEAR_DEFAULT_CTA_GROUP;
  const int per = peer_mtiles(cta_group, m, n_peers) * 128 * cta_group;
  long long r0 = (long long)rank * per, r1 = (rank == n_peers - 1) ? m : r0 + per;
  if (r0 > m) r0 = m;
  if (r1 > m) r1 = m;
  *row0 = (int)r0;
  *row1 = (int)r1;
}

// diagnostic tile-order override CLEAR_GEMM_RASTER="<m_fast>,<group>", read once
static bool raster_override(int* m_fast, int* group) {
  static const struct Ov { bool on; int mf, grp; } ov = [] {
    Ov o{false, 0, 0};
    if (const char* e = getenv("CLEAR_GEMM_RASTER")) o.on = (sscanf(e, "%d,%d", &o.mf, &o.grp) == 2 && o.grp > 0);
    return o;
  }();
  if (ov.on) { *m_fast = ov.mf ? 1 : 0; *group = ov.grp; }
  return ov.on;
}

// One problem of a launch: argument checks, tensor maps, tile order.
static int fill_problem(int cta_group, const GemmDesc& g, CUtensorMap* tmA, CUtensorMap* tmB, CUtensorMap* tmD, GemmProblem* q,
                        PeerMaps* peer_maps, bool* uses_peers) {
  const GemmOperand &a = g.a, &b = g.b;
  const int m = g.m, n = g.n, k = g.k;
  if (!a.ptr || !b.ptr || !g.d) return fail(CLEAR_ERR_NULL, "gemm: null operand");
  if (m <= 0 || n <= 0 || k <= 0) return fail(CLEAR_ERR_SHAPE, "gemm: non-positive dims m=%d n=%d k=%d", m, n, k);
  if ((a.ld % 8) || (b.ld % 8) || (g.ldd % (g.d_is_f32 ? 4 : 8)))
    return fail(CLEAR_ERR_ALIGN, "gemm: leading dimensions must be multiples of 16 bytes (lda=%lld ldb=%lld ldd=%lld)",
                (long long)a.ld, (long long)b.ld, (long long)g.ldd);
  if ((reinterpret_cast<uintptr_t>(a.ptr) | reinterpret_cast<uintptr_t>(b.ptr) | reinterpret_cast<uintptr_t>(g.d)) & 15)
    return fail(CLEAR_ERR_ALIGN, "gemm: pointers must be 16-byte aligned");
  if (a.major == kMajorMN && a.ld < m) return fail(CLEAR_ERR_SHAPE, "gemm: MN-major A needs lda >= m");
  if (b.major == kMajorMN && b.ld < n) return fail(CLEAR_ERR_SHAPE, "gemm: MN-major B needs ldb >= n");
  if (a.major == kMajorK && a.ld < k) return fail(CLEAR_ERR_SHAPE, "gemm: lda < k");
  if (b.major == kMajorK && b.ld < k) return fail(CLEAR_ERR_SHAPE, "gemm: ldb < k");
  if (g.ldd < n) return fail(CLEAR_ERR_SHAPE, "gemm: ldd < n");
  if (g.lse_part && g.d_is_f32) return fail(CLEAR_ERR_KIND, "gemm: lse_part requires a bf16 output");
  if (g.reduce_add && g.lse_part) return fail(CLEAR_ERR_KIND, "gemm: accumulate cannot be combined with lse_part");
  if (g.dyn_dim != 0 && g.dyn_dim != 1 && g.dyn_dim != 2) return fail(CLEAR_ERR_KIND, "gemm: dyn_dim must be 0, 1 (M) or 2 (K)");
  int rc;
  if ((rc = make_operand_map(tmA, a, m, k, 128))) return rc;
  if ((rc = make_operand_map(tmB, b, n, k, 256 / cta_group))) return rc;
  if ((rc = make_output_map(tmD, g.d, g.d_is_f32, g.ldd, m, n))) return rc;
  if (g.n_peers > 0) {
    if (g.n_peers > kMaxPeers || !g.peer_d) return fail(CLEAR_ERR_SHAPE, "gemm: 1 <= n_peers <= %d with a pointer table required", kMaxPeers);
    if (g.lse_part) return fail(CLEAR_ERR_KIND, "gemm: the peer-sharded epilogue excludes lse_part");
    if (*uses_peers) return fail(CLEAR_ERR_KIND, "gemm: only one problem of a launch can be peer-sharded");
    for (int r = 0; r < g.n_peers; ++r) {
      if (!g.peer_d[r] || (reinterpret_cast<uintptr_t>(g.peer_d[r]) & 15)) return fail(CLEAR_ERR_ALIGN, "gemm: peer buffer %d is NULL or misaligned", r);
      if ((rc = make_output_map(&peer_maps->tm[r], g.peer_d[r], g.d_is_f32, g.ldd, m, n))) return rc;
    }
    for (int r = g.n_peers; r < kMaxPeers; ++r) peer_maps->tm[r] = peer_maps->tm[0];
    *uses_peers = true;
  }
  q->M = m; q->N = n; q->K = k;
  q->m_tiles = (m + 128 * cta_group - 1) / (128 * cta_group);
  q->n_tiles = (n + 255) / 256;
  if (g.raster_m_fast < 0) {
    pick_raster(cta_group, m, n, k, &q->raster_m_fast, &q->raster_group);
  } else {
    q->raster_m_fast = g.raster_m_fast ? 1 : 0;
    q->raster_group = g.raster_m_fast ? q->m_tiles : q->n_tiles;  // whole-dimension chunks
  }
  raster_override(&q->raster_m_fast, &q->raster_group);
  {
    const int fast = q->raster_m_fast ? q->m_tiles : q->n_tiles;
    if (q->raster_group > fast) q->raster_group = fast;
    if (q->raster_group < 1) q->raster_group = 1;
  }
  q->reduce_add = g.reduce_add;
  q->a_major = a.major;
  q->lse_part = reinterpret_cast<float2*>(g.lse_part);
  q->n_peers = g.n_peers;
  q->mtiles_per_peer = g.n_peers > 0 ? peer_mtiles(cta_group, m, g.n_peers) : 1;
  q->dyn_dim = g.dyn_dim;
  return 0;
}

static int epilogue_kind(const GemmDesc& g) { return g.d_is_f32 ? kEpiStoreF32 : (g.lse_part ? kEpiStoreBF16Lse : kEpiStoreBF16); }

// One launch over one or two problems (second == nullptr: one).  Two problems must agree in B's major and in the epilogue
// kind; A's major may differ (the kernel is then instantiated with a run-time A major); no split-K.
int launch_gemm_group(int cta_group, const GemmDesc& first, const GemmDesc* second, cudaStream_t stream, int k_splits,
                      const int* dyn_count, int dyn_offset, int dyn_splits) {
  if (cta_group == 0) cta_group = CLEAR_DEFAULT_CTA_GROUP;
  if (cta_group != 1 && cta_group != 2) return fail(CLEAR_ERR_KIND, "gemm: cta_group must be 0, 1 or 2");
  if (k_splits < 1) k_splits = 1;
  if (k_splits > 1 && !first.reduce_add) return fail(CLEAR_ERR_KIND, "gemm: split-K requires the reduce-add epilogue (accumulate = 1)");
  if (second && (k_splits > 1 || dyn_splits)) return fail(CLEAR_ERR_KIND, "gemm: a two-problem launch excludes split-K");
  if (second && (second->b.major != first.b.major || epilogue_kind(*second) != epilogue_kind(first)))
    return fail(CLEAR_ERR_KIND, "gemm: the two problems of a launch must share B's major and the epilogue kind");
  if (first.n_peers > 0 && k_splits > 1) return fail(CLEAR_ERR_KIND, "gemm: the peer-sharded epilogue excludes split-K");

  CUtensorMap tm[6];
  static PeerMaps no_peers = {};
  PeerMaps peer_maps_storage;
  bool uses_peers = false;
  GemmParams p = {};
  int rc;
  if ((rc = fill_problem(cta_group, first, &tm[0], &tm[1], &tm[2], &p.q[0], &peer_maps_storage, &uses_peers))) return rc;
  if (second) {
    if ((rc = fill_problem(cta_group, *second, &tm[3], &tm[4], &tm[5], &p.q[1], &peer_maps_storage, &uses_peers))) return rc;
  } else {
    tm[3] = tm[0]; tm[4] = tm[1]; tm[5] = tm[2];
    p.q[1] = p.q[0];
    p.q[1].n_peers = 0;
  }
  p.n_problems = second ? 2 : 1;
  {
    const int num_kb = (first.k + 63) / 64;
    if (k_splits > num_kb) k_splits = num_kb;
    p.k_splits = k_splits;
    p.kb_per_split = (num_kb + k_splits - 1) / k_splits;
    p.k_splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty item
  }
  if ((rc = next_work_counter(&p.work_counter))) return rc;
  const bool any_dyn = first.dyn_dim != 0 || (second && second->dyn_dim != 0);
  p.dyn_count = any_dyn ? dyn_count : nullptr;
  if (any_dyn && !dyn_count) return fail(CLEAR_ERR_NULL, "gemm: dyn_dim without a device-side count");
  if (!p.dyn_count) { p.q[0].dyn_dim = 0; p.q[1].dyn_dim = 0; }
  p.dyn_offset = dyn_offset;
  p.dyn_splits = (p.dyn_count && first.dyn_dim == 1 && !second) ? dyn_splits : 0;
  p.sm_groups = 1;
  if (first.dyn_dim == 2 && p.k_splits > 1) return fail(CLEAR_ERR_KIND, "gemm: a device-side K extent excludes split-K");
  if (p.dyn_splits && !(first.reduce_add && first.d_is_f32)) return fail(CLEAR_ERR_KIND, "gemm: in-kernel split-K needs the fp32 reduce-add epilogue");

  const PeerMaps& peers = uses_peers ? peer_maps_storage : no_peers;
  const int epi = epilogue_kind(first);
  const int a_major = (second && second->a.major != first.a.major) ? (int)kMajorRuntime : first.a.major;
  const int b_major = first.b.major;
#define CLEAR_DISPATCH(CG, MA, MB, EPI)                                                     \
  if (cta_group == CG && a_major == MA && b_major == MB && epi == EPI)                      \
    return launch_variant<CG, MA, MB, EPI>(tm, peers, p, stream);
  CLEAR_DISPATCH(1, kMajorK, kMajorK, kEpiStoreBF16)
  CLEAR_DISPATCH(1, kMajorK, kMajorK, kEpiStoreBF16Lse)
  CLEAR_DISPATCH(1, kMajorK, kMajorMN, kEpiStoreBF16)
  CLEAR_DISPATCH(1, kMajorK, kMajorMN, kEpiStoreF32)
  CLEAR_DISPATCH(1, kMajorMN, kMajorMN, kEpiStoreBF16)
  CLEAR_DISPATCH(1, kMajorMN, kMajorMN, kEpiStoreF32)
  CLEAR_DISPATCH(1, kMajorRuntime, kMajorMN, kEpiStoreBF16)
  CLEAR_DISPATCH(2, kMajorK, kMajorK, kEpiStoreBF16)
  CLEAR_DISPATCH(2, kMajorK, kMajorK, kEpiStoreBF16Lse)
  CLEAR_DISPATCH(2, kMajorK, kMajorMN, kEpiStoreBF16)
  CLEAR_DISPATCH(2, kMajorK, kMajorMN, kEpiStoreF32)
  CLEAR_DISPATCH(2, kMajorMN, kMajorMN, kEpiStoreBF16)
  CLEAR_DISPATCH(2, kMajorMN, kMajorMN, kEpiStoreF32)
  CLEAR_DISPATCH(2, kMajorRuntime, kMajorMN, kEpiStoreBF16)
#undef CLEAR_DISPATCH
  return fail(CLEAR_ERR_KIND, "gemm: operand layout combination (a_major=%d, b_major=%d, epilogue=%d) is not built",
              a_major, b_major, epi);
}

int launch_gemm_bf16(int cta_group, GemmOperand a, GemmOperand b, void* d, int d_is_f32, int64_t ldd, int m, int n,
                     int k, float* lse_part, int reduce_add, int raster_m_fast, cudaStream_t stream, int k_splits,
                     void* const* peer_d, int n_peers, GemmDyn dyn) {
  GemmDesc g;
  g.a = a; g.b = b; g.d = d; g.d_is_f32 = d_is_f32; g.ldd = ldd; g.m = m; g.n = n; g.k = k;
  g.lse_part = lse_part; g.reduce_add = reduce_add; g.raster_m_fast = raster_m_fast;
  g.peer_d = peer_d; g.n_peers = n_peers;
  g.dyn_dim = dyn.count ? dyn.dim : 0;
  return launch_gemm_group(cta_group, g, nullptr, stream, k_splits, dyn.count, dyn.offset, dyn.splits);
}

}  // namespace clear

extern "C" {

const char* clear_last_error(void) { return clear::last_error(); }
int clear_abi_version(void) { return CLEAR_ABI_VERSION; }

int clear_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0, maj = 0, min = 0, sms = 0;
  CLEAR_CUDA_OK(cudaGetDevice(&dev));
  CLEAR_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  CLEAR_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  CLEAR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) return clear::fail(CLEAR_ERR_DEVICE, "device is sm_%d%d; this library is built for sm_100a only", maj, min);
  return 0;
}

int clear_event_create(void** event) {
  if (!event) return clear::fail(CLEAR_ERR_NULL, "clear_event_create: NULL");
  cudaEvent_t e;
  CLEAR_CUDA_OK(cudaEventCreate(&e));
  *event = e;
  return 0;
}

int clear_event_elapsed_ms(void* start, void* stop, float* ms_host) {
  if (!start || !stop || !ms_host) return clear::fail(CLEAR_ERR_NULL, "clear_event_elapsed_ms: NULL");
  CLEAR_CUDA_OK(cudaEventElapsedTime(ms_host, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  return 0;
}

int clear_event_destroy(void* event) {
  if (!event) return 0;
  CLEAR_CUDA_OK(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return 0;
}

int clear_gemm_bf16(const ClearGemmArgs* g, void* stream) {
  if (!g) return clear::fail(CLEAR_ERR_NULL, "clear_gemm_bf16: args is NULL");
  clear::GemmOperand a{g->a, g->lda, g->a_major};
  clear::GemmOperand b{g->b, g->ldb, g->b_major};
  const int raster = g->raster_m_fast;  // < 0: chosen by the traffic model (pick_raster)
  return clear::launch_gemm_bf16(g->cta_group, a, b, g->d, g->d_f32, g->ldd, g->m, g->n, g->k, g->lse_part,
                                 g->accumulate, raster, static_cast<cudaStream_t>(stream), g->k_splits);
}

}  // extern "C"
