// Host-side shared helpers: error codes / thread-local last-error string, device attributes.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/clear_b200.h"

namespace clear {

// Records a message retrievable through clear_last_error() and returns `code` (so call sites can
// `return fail(CLEAR_ERR_SHAPE, "...")`).
int fail(int code, const char* fmt, ...);
// Maps a cudaError_t to the ABI convention (positive = cudaError_t) and records its string.
int cuda_fail(cudaError_t e, const char* what);

#define CLEAR_CUDA_OK(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::clear::cuda_fail(_e, #expr); \
  } while (0)

int device_sm_count();  // cached; <0 on error
const char* last_error();

// MMA grouping used when the caller passes cta_group = 0.
#ifndef CLEAR_DEFAULT_CTA_GROUP
#define CLEAR_DEFAULT_CTA_GROUP 2
#endif

// bf16 GEMM launcher used by the head entry points and by clear_gemm_bf16 (tests / roofline runs).
struct GemmOperand {
  const void* ptr;
  int64_t ld;  // leading dimension in elements of the row-major storage
  int major;   // 0: K-major  (storage [MN, K]);  1: MN-major (storage [K, MN])
};
// peer_d / n_peers: optional output buffers of the data-parallel peers (same shape and ldd as d; d itself may be one of
// them): M tiles are reduce-added into the buffer of their owner rank (see PeerMaps in gemm_sm100.cuh).
// dyn: optional device-side extent.  *count - offset (clamped) replaces m (dim = 1) or k (dim = 2, rounded up to a multiple of
// 64; the operand rows of the last K block beyond the count must be zero); m / k are then upper bounds.  splits = 1 (dim = 1,
// fp32 reduce-add output, pre-zeroed): the number of K splits is chosen in the kernel from the actual m.
struct GemmDyn {
  const int* count = nullptr;
  int offset = 0;
  int dim = 0;
  int splits = 0;
};
int launch_gemm_bf16(int cta_group, GemmOperand a, GemmOperand b, void* d, int d_is_f32, int64_t ldd, int m, int n,
                     int k, float* lse_part, int reduce_add, int raster_m_fast, cudaStream_t stream, int k_splits = 1,
                     void* const* peer_d = nullptr, int n_peers = 0, GemmDyn dyn = GemmDyn());
int current_device();   // ordinal of the current device, < 0 on error (per-device caches)
// One contraction D[m, n] = A[m, k] B[n, k]^T of a launch (see GemmProblem in gemm_sm100.cuh).
struct GemmDesc {
  GemmOperand a, b;
  void* d = nullptr;
  int d_is_f32 = 0;
  int64_t ldd = 0;
  int m = 0, n = 0, k = 0;
  float* lse_part = nullptr;
  int reduce_add = 0;
  int raster_m_fast = -1;       // < 0: the traffic model picks the tile order
  void* const* peer_d = nullptr;
  int n_peers = 0;
  int dyn_dim = 0;              // 1: m, 2: k is bounded by the launch's device-side count
};
// One launch over one or two independent problems that share B's major and the epilogue kind: a single list of work items
// (the first problem's first), one pipeline ramp and one tail instead of two.
int launch_gemm_group(int cta_group, const GemmDesc& first, const GemmDesc* second, cudaStream_t stream, int k_splits = 1,
                      const int* dyn_count = nullptr, int dyn_offset = 0, int dyn_splits = 0);
// Rows [row0, row1) of an [m, *] output that rank `rank` of `n_peers` owns under that sharding.
void gemm_peer_rows(int cta_group, int m, int n_peers, int rank, int* row0, int* row1);
// Number of K splits (1..4) that best fills the last wave of CTA groups for this problem.
int gemm_pick_k_splits(int cta_group, int m, int n, int k);

}  // namespace clear
