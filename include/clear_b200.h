/*
 * clear_b200.h — C ABI of libclear_b200.so, the B200 (sm_100a) implementation of the CLEAR
 * unlearning-step hot path (reference: somvy/multimodal_unlearning).
 *
 * The reference is pure Python: its "FFI" for this path is the set of eager PyTorch calls made by
 *   MMTrainerForgetting.compute_loss          mm/trainer.py:168-483
 *   CustomTrainerForgetting.compute_loss      nlp/dataloader.py:144-485
 *   logits2probs / get_batch_loss             mm/trainer_utils.py:16-21,73-81
 *   save_gradient_ratio                       methods/salun/generate_mask.py:17-89
 * Each entry point below names the reference call sites it replaces.  A maintainer binds this
 * header with ctypes; multimodal_unlearning_b200/_lib.py is that binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a CUDA device pointer unless marked "host";
 *   - the caller owns every buffer (inputs, outputs, workspace); the library frees nothing of the
 *     caller's and keeps no pointer after a call returns.  Its one device allocation is made on the
 *     first contraction on a device: a 32 KB ring of work counters for the dynamic tile scheduler;
 *   - all work is enqueued on the cudaStream_t passed in (as a void*); no hidden synchronisation;
 *   - return value: 0 success, <0 a ClearStatus argument/shape error detected before any launch,
 *     >0 a cudaError_t.  clear_last_error() returns a thread-local message for the last failure;
 *   - bf16 is the storage type of hidden states, weights, logits and gradients; all arithmetic
 *     accumulates in fp32.
 */
#ifndef CLEAR_B200_H_
#define CLEAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLEAR_ABI_VERSION 3

typedef enum ClearStatus {
  CLEAR_OK = 0,
  CLEAR_ERR_NULL = -1,      /* required pointer is NULL */
  CLEAR_ERR_SHAPE = -2,     /* size / leading-dimension constraint violated */
  CLEAR_ERR_ALIGN = -3,     /* pointer or stride not 16-byte aligned */
  CLEAR_ERR_KIND = -4,      /* unknown enum value */
  CLEAR_ERR_WORKSPACE = -5, /* workspace too small */
  CLEAR_ERR_DEVICE = -6,    /* not an sm_100 device / driver entry point missing */
} ClearStatus;

/* Loss kinds evaluated row-wise on the logits. */
typedef enum ClearLossKind {
  CLEAR_KIND_CE = 0,      /* -log p[y]; HF causal-LM CE (A2) and get_batch_loss (A6c)             */
  CLEAR_KIND_KL_FWD = 1,  /* sum_v q (log q - log p): kl_div(student, teacher)   mm/trainer.py:203 */
  CLEAR_KIND_KL_REV = 2,  /* sum_v p (log p - log q): kl_div(teacher, student)   mm/trainer.py:237 */
  CLEAR_KIND_NPO = 3,     /* sum_v log(1 + (p/(q+1e-8))^beta)                    mm/trainer.py:348 */
  CLEAR_KIND_ENTROPY = 4, /* sum_v p log p                                       mm/trainer.py:195 */
} ClearLossKind;

const char* clear_last_error(void);
int clear_abi_version(void);
/* host out-params; fails with CLEAR_ERR_DEVICE when the current device is not compute capability 10.x */
int clear_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Dense bf16 contraction on tcgen05/TMEM:  D[M,N] = A[M,K] * B[N,K]^T, fp32 accumulate.
 * Building block of the three projections of the head (forward logits, dH, dW); exported so that
 * tests and the roofline benchmark can drive each operand layout directly.
 *   a_major / b_major: 0 = K-major  (operand stored row-major as [M or N, K], ld = row stride)
 *                      1 = MN-major (operand stored row-major as [K, M or N], ld = row stride)
 *   d_f32: 0 -> D is bf16, 1 -> D is fp32;  accumulate: D += instead of D = (TMA reduce-add; a bf16 D is
 *          rounded to bf16 after each add)
 *   lse_part: optional float2[M][ceil(N/256)] = (row max, sum exp(x - max)) over each 256-column
 *             tile of the bf16-rounded D (bf16 D only); NULL to skip.
 *   cta_group: 1, 2 (CTA-pair MMA, 256x256 tiles) or 0 = library default.
 * Constraints: all leading dimensions multiples of 8 elements, pointers 16-byte aligned; an
 * MN-major operand needs ld >= round_up(extent, 64).
 * ------------------------------------------------------------------------------------------- */
typedef struct ClearGemmArgs {
  const void* a;
  const void* b;
  void* d;
  float* lse_part;
  int64_t lda, ldb, ldd;
  int32_t m, n, k;
  int32_t a_major, b_major;
  int32_t d_f32, accumulate;
  int32_t cta_group;
  int32_t raster_m_fast; /* 1: walk M first (default for forward), 0: walk N first, -1: auto */
  int32_t k_splits;      /* 0/1: none; 2..4: split K over work items (needs d_f32 = accumulate = 1, D pre-zeroed) */
} ClearGemmArgs;
int clear_gemm_bf16(const ClearGemmArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loss head, stage 1 — projection  z = hidden * weight^T   (reference: the lm_head nn.Linear inside
 * `model(**inputs)`, mm/trainer.py:180,184,188-189,...; teacher: :200,217,235,...).
 *   logits: out bf16 [n_rows, logits_ld], logits_ld >= round_up(vocab, 64), multiple of 64.
 *   lse_part: out float2 [n_rows][clear_head_vtiles(vocab)] partial log-sum-exp statistics.
 * ------------------------------------------------------------------------------------------- */
int clear_head_vtiles(int vocab);
typedef struct ClearProjectArgs {
  const void* hidden; /* bf16 [n_rows, hidden_dim], row stride hidden_ld */
  const void* weight; /* bf16 [vocab, hidden_dim], row stride weight_ld */
  void* logits;
  float* lse_part;
  int64_t hidden_ld, weight_ld, logits_ld;
  int64_t n_rows;
  int32_t hidden_dim, vocab;
  int32_t cta_group;
  /* optional device-side row count (row compaction without a host round trip): only the first
   * clamp(*n_rows_dev - n_rows_dev_offset, 0, n_rows) rows are projected; n_rows stays the allocation / upper bound */
  const int32_t* n_rows_dev;
  int32_t n_rows_dev_offset;
} ClearProjectArgs;
int clear_head_project(const ClearProjectArgs* args, void* stream);
/* The student's and the frozen teacher's projection of the same rows as ONE launch (two independent contractions whose
 * work items share the persistent grid: one pipeline ramp and one tail instead of two).  Both must agree in lse_part
 * presence, cta_group and n_rows_dev. */
int clear_head_project_pair(const ClearProjectArgs* student, const ClearProjectArgs* teacher, void* stream);
/* Same statistics for teacher logits handed over already materialised (`self.teacher_model(**x).logits`,
 * mm/trainer.py:199-200): lse_part float2 [n_rows][clear_head_vtiles(vocab)] from bf16 logits. */
int clear_head_lse_partials(const void* logits, int64_t logits_ld, int64_t n_rows, int32_t vocab, float* lse_part,
                            void* stream);
/* the same for fp32 teacher logits (a teacher evaluated under autocast, or an fp32 oracle model) */
int clear_head_lse_partials_f32(const void* logits, int64_t logits_ld, int64_t n_rows, int32_t vocab, float* lse_part,
                                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loss head, stage 2 — row-wise loss and d(loss)/d(logits), written in place over the logits.
 * Replaces logits2probs + kl_div / CE / NPO / entropy arithmetic (mm/trainer_utils.py:16-21,
 * mm/trainer.py:195-208,237-255,348-353; HF ForCausalLMLoss) and their autograd backward.
 *   labels (CE only): already shifted (labels[b,t+1] at row b*T+t), -100 = ignored row.
 *   loss_rows[i] = unscaled loss term of row i (0 for ignored CE rows).
 *   dlogits[i,v] = grad_scale * (scale_dev ? *scale_dev : 1) * (row_weight ? row_weight[i] : 1) * dterm(i,v)
 *   (dterm per App. A of SURVEY.md; NPO dterm = s - p*sum(s), i.e. WITHOUT the 2/(N V) factor).
 *   Padding columns [vocab, logits_ld) are zeroed.
 * ------------------------------------------------------------------------------------------- */
typedef struct ClearRowLossArgs {
  void* logits;             /* in: bf16 logits; out: bf16 dlogits */
  const float* lse_part;    /* from clear_head_project on the same rows */
  const void* ref_logits;   /* teacher logits bf16 (KL_FWD, KL_REV, NPO) else NULL */
  const float* ref_lse_part;
  const int64_t* labels;    /* CE only */
  const float* row_weight;  /* optional [n_rows] */
  const float* scale_dev;   /* optional device scalar */
  float* loss_rows;         /* out [n_rows] */
  int64_t logits_ld, ref_ld;
  int64_t n_rows;
  int32_t vocab;
  int32_t kind;             /* ClearLossKind */
  float grad_scale;
  float beta;               /* NPO exponent (loss_beta) */
  int32_t write_grad;       /* 0: forward only (logits left untouched) */
  int32_t ref_logits_f32;   /* ref_logits is fp32 [n_rows, ref_ld] instead of bf16 */
  const int32_t* n_rows_dev; /* CE only: device-side row count (see ClearProjectArgs); the dlogits rows between the count and the */
  int32_t n_rows_dev_offset; /* next multiple of 64 are zeroed (they fall into the last K block of the dW contraction)        */
  int32_t* status_flags;    /* optional device int32: bit 0 is set when a CE label lies outside [0, vocab) (torch asserts there;
                               the row's loss term is NaN, its dlogits zero) */
} ClearRowLossArgs;
int clear_head_rowloss(const ClearRowLossArgs* args, void* stream);

/* inv_count[0] = 1 / #{i : labels[i] != -100}   (inf when none: the reference's mean CE is NaN then) */
int clear_head_count_valid(const int64_t* labels, int64_t n_rows, float* inv_count, void* stream);

/* Row compaction for CE-type losses: rows whose (shifted) label is -100 contribute nothing to the loss, to dW or to
 * their own dH (HF ForCausalLMLoss ignore_index; the reference still projects them and backpropagates zero rows —
 * in CLEAR's batches the padding, the 576 image positions and the question).  row_index[0 .. *count) = ascending
 * indices of the valid rows (device int32; row_index holds n_rows entries).  The caller reads *count and runs the head
 * on the gathered rows; gather / scatter move rows of row_bytes (a multiple of 4) between strided buffers:
 *   gather:  dst[i] = src[row_index[i]]       scatter: dst[row_index[i]] = src[i]       i < n_idx */
int clear_head_compact_rows(const int64_t* labels, int64_t n_rows, int32_t* row_index, int32_t* count, void* stream);
int clear_rows_gather(const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx, int32_t row_bytes,
                      void* dst, int64_t dst_ld_bytes, void* stream);
int clear_rows_scatter(const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx, int32_t row_bytes,
                       void* dst, int64_t dst_ld_bytes, void* stream);
/* The same with the number of indices read from device memory (*n_idx_dev <= n_idx_max; NULL = n_idx_max), so that the
 * count of clear_head_compact_rows never visits the host.  Gather: row_index NULL = identity, and with pad_to > 0 the
 * destination rows [n, round_up(n, pad_to)) below n_idx_max are zero-filled.
 * clear_index_compose: out[i] = outer[inner[i]], i < n — valid rows of a batch whose rows are themselves a row map
 * (remove_image_tokens, mm/trainer_utils.py:48-70) of the trunk's output. */
int clear_rows_gather_dev(const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx_max,
                          const int32_t* n_idx_dev, int32_t pad_to, int32_t row_bytes, void* dst, int64_t dst_ld_bytes, void* stream);
int clear_rows_scatter_dev(const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx_max,
                           const int32_t* n_idx_dev, int32_t row_bytes, void* dst, int64_t dst_ld_bytes, void* stream);
int clear_index_compose(const int32_t* inner, const int32_t* outer, int64_t n_max, const int32_t* n_dev, int32_t* out, void* stream);
/* out[0] (+)= scale * (scale_dev ? *scale_dev : 1) * sum_i loss_rows[i]   — deterministic single-block tree */
int clear_head_reduce_rows(const float* loss_rows, int64_t n_rows, float scale, const float* scale_dev,
                           float* out, int accumulate, void* stream);
/* seq_sums[b] = sum_t loss_rows[b*rows_per_seq + t]   (get_batch_loss, mm/trainer_utils.py:73-81) */
int clear_head_reduce_sequences(const float* loss_rows, int64_t n_seqs, int64_t rows_per_seq, float* seq_sums,
                                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loss head, stage 3 — backward projections (autograd of nn.Linear, K9):
 *   dhidden[n_rows, hidden_dim] = dlogits * weight           (bf16 out; NULL to skip)
 *   dweight[vocab, hidden_dim]  = dlogits^T * hidden         (bf16 or fp32; accumulate: += , bf16 rounds per add)
 * ------------------------------------------------------------------------------------------- */
typedef struct ClearBackwardArgs {
  const void* dlogits; /* bf16 [n_rows, dlogits_ld] */
  const void* hidden;
  const void* weight;
  void* dhidden;
  void* dweight;
  int64_t dlogits_ld, hidden_ld, weight_ld, dhidden_ld, dweight_ld;
  int64_t n_rows;
  int32_t hidden_dim, vocab;
  int32_t dweight_f32, accumulate;
  int32_t cta_group;
  /* optional fp32 [n_rows, hidden_dim] workspace (row stride hidden_dim): lets the library cut the dH contraction
   * along the vocabulary (split-K, fp32 reduce-add, one conversion pass) when its output tiles alone would leave
   * the last wave of SM pairs mostly idle.  NULL: never split. */
  float* dhidden_ws;
  /* Fused gradient reduce-scatter (data parallel): n_peers > 0 and dweight_peers[r] = rank r's bf16
   * [vocab, dweight_ld] accumulation buffer mapped into this process (symmetric memory over NVLink / NVSwitch).  The dW
   * contraction's epilogue then reduce-adds (bf16, like DDP's bf16 gradient all-reduce) every output tile into the buffer
   * of the rank that owns its rows (clear_head_dweight_slab) instead of writing `dweight`, which may be NULL.  The owners'
   * slabs must have been zeroed, and all ranks synchronised, before any rank launches; clear_dweight_allgather finishes
   * the all-reduce. */
  void* dweight_peers[8];
  int32_t n_peers;
  /* optional device-side row count (see ClearProjectArgs): dH is formed for the live rows only (with dhidden_ws the K split
   * that fills the SMs is chosen inside the kernel); dW contracts over round_up(count, 64) rows — dlogits and hidden must be
   * zero between the count and that bound (clear_head_rowloss / clear_rows_gather_dev see to it). */
  const int32_t* n_rows_dev;
  int32_t n_rows_dev_offset;
} ClearBackwardArgs;
int clear_head_backward(const ClearBackwardArgs* args, void* stream);
/* Rows [row0, row1) of dW that `rank` owns among n_peers (contiguous slabs of whole output tiles). */
int clear_head_dweight_slab(int32_t vocab, int32_t n_peers, int32_t rank, int32_t cta_group, int32_t* row0, int32_t* row1);
/* All-gather half of the fused all-reduce: rows [row0, row1) of this rank's reduced slab, times `scale` (1 / world size
 * for DDP's mean), are written to the same rows of EVERY rank's buffer — through the NVSwitch multicast mapping of the
 * buffers when multicast_buf != NULL (one multimem.st per 16 bytes; the switch replicates), else by plain peer stores to
 * each of the n_peers pointers.  bf16, row stride ld elements, hidden_dim a multiple of 8. */
int clear_dweight_allgather(const void* local_buf, void* multicast_buf, void* const* peer_bufs, int32_t n_peers, int64_t ld,
                            int32_t row0, int32_t row1, int32_t hidden_dim, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Loss head in ONE call: everything the reference evaluates between "the trunk produced hidden states"
 * and "autograd filled dH and lm_head.weight.grad" for one `model(**split)` — MMTrainerForgetting.compute_loss
 * mm/trainer.py:168-449, CustomTrainerForgetting.compute_loss nlp/dataloader.py:144-459.  Rows are streamed in blocks of
 * block_rows through the three stages above; the workspace holds the one bf16 staging block per projected operand.
 *   loss      = coef * norm * sum_i loss_rows[i]   with the reference's normalisation (reduction = REFERENCE: CE mean over
 *               non-ignored rows, KL / N, mm NPO (2/beta) mean over N V, entropy raw sum), a plain sum, or the mean
 *               over n_seqs of per-sequence sums (get_batch_loss, mm/trainer_utils.py:73-81);
 *   dhidden   = d loss / d hidden (bf16), dweight = d loss / d weight (bf16 or fp32; dweight_accumulate: +=), or the
 *               fused reduce-scatter targets dweight_peers (see ClearBackwardArgs); all NULL = forward only.
 * ------------------------------------------------------------------------------------------- */
typedef enum ClearReduction {
  CLEAR_REDUCE_REFERENCE = 0,
  CLEAR_REDUCE_SUM = 1,
  CLEAR_REDUCE_SEQ_MEAN = 2,
} ClearReduction;

typedef struct ClearHeadArgs {
  const void* hidden;      /* bf16 [n_rows, hidden_ld] */
  const void* weight;      /* bf16 [vocab, weight_ld] */
  const int64_t* labels;   /* CE: shifted labels [n_rows], -100 = ignored; NULL for the other kinds */
  const void* ref_hidden;  /* frozen teacher as (hidden, weight): projected here block by block ... */
  const void* ref_weight;
  const void* ref_logits;  /* ... or as materialised bf16 logits [n_rows, ref_logits_ld] */
  const float* row_weight; /* optional [n_rows] */
  float* loss;             /* out fp32 [1] */
  float* loss_rows;        /* out fp32 [n_rows]: unscaled row terms */
  float* seq_loss;         /* out fp32 [n_seqs] per-sequence sums of the row terms, or NULL */
  void* dhidden;           /* out bf16 [n_rows, dhidden_ld], or NULL */
  void* dweight;           /* out bf16 / fp32 [vocab, dweight_ld], or NULL */
  void* dweight_peers[8];  /* fused gradient reduce-scatter: see ClearBackwardArgs */
  int32_t n_peers;
  int64_t hidden_ld, weight_ld, ref_hidden_ld, ref_weight_ld, ref_logits_ld, dhidden_ld, dweight_ld;
  int64_t n_rows;
  int32_t hidden_dim, vocab;
  int32_t kind;            /* ClearLossKind */
  double beta;             /* NPO exponent (loss_beta); double: the scale factors are formed in double, rounded once */
  double coef;             /* the recipe's coefficient on this term (e.g. -loss_beta for a forget term) */
  int32_t reduction;       /* ClearReduction */
  int32_t n_seqs;
  int32_t block_rows;      /* rows staged per pass; 0 = 8192 */
  int32_t dweight_f32, dweight_accumulate;
  int32_t cta_group;
  /* ---- ABI 3 ---- */
  /* Row map (remove_image_tokens, mm/trainer_utils.py:48-70, as a kernel input instead of a copy of [B, L, V] logits): row i
   * of this invocation is row row_index[i] of hidden / ref_hidden, which then hold n_src_rows rows; its gradient goes to that
   * row of dhidden [n_src_rows, dhidden_ld], whose other rows are zero-filled.  labels, ref_logits, row_weight, loss_rows and
   * the sequences of seq_loss are indexed by i.  NULL = identity. */
  const int32_t* row_index;
  int64_t n_src_rows;
  /* CE only: evaluate the rows whose label is not -100 only (they alone contribute to the loss, to dW and to their own dH):
   * valid rows are compacted on the device and their count never visits the host — every stage reads it from device memory.
   * Same loss, loss_rows, dhidden, dweight as the full evaluation (dW up to the summation order of the bf16 products). */
  int32_t compact_ignored;
  int32_t ref_logits_f32;  /* ref_logits is fp32 instead of bf16 */
  int32_t* status_flags;   /* optional device int32, OR-ed: bit 0 = a CE label outside [0, vocab) (torch raises a device assert) */
  /* optional cudaEvent_t: the stream waits on it (cudaStreamWaitEvent) right before the first launch that writes into
   * dweight_peers — the owners have zeroed their slabs — instead of the caller holding the whole invocation back */
  void* dweight_ready_event;
  /* optional profiling hooks: an array of cudaEvent_t (4 per block of rows, n_stage_events >= 4 * number of blocks or the
   * surplus blocks go unrecorded) that the library records on the stream before the projection(s) of a block, after them,
   * after the row pass and after the backward contractions — so that per-kernel times can be measured with CUDA events
   * INSIDE a timed region that runs this very entry point (bench.py's roofline numbers) */
  void* const* stage_events;
  int32_t n_stage_events;
} ClearHeadArgs;
/* timing events for ClearHeadArgs::stage_events (cudaEvent_t handles of the library's own CUDA runtime instance) */
int clear_event_create(void** event);
int clear_event_elapsed_ms(void* start, void* stop, float* ms_host); /* both events must have completed */
int clear_event_destroy(void* event);
int clear_head_loss_workspace_bytes(const ClearHeadArgs* args, size_t* out_bytes);
int clear_head_loss_fwd_bwd(const ClearHeadArgs* args, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SalUn saliency mask (methods/salun/generate_mask.py:17-89) and the masked update.
 * All tensors of a model live in flat buffers; `seg_offsets` (host, n_segs+1 int64, element units,
 * every offset a multiple of 8) delimits the per-parameter tensors — thresholds are per tensor, as
 * in the reference (:54-85).
 * ------------------------------------------------------------------------------------------- */
/* acc (bf16) += grad (bf16), rounding like torch's bf16 add (generate_mask.py:41-44) */
int clear_saliency_accumulate(void* acc, const void* grad, int64_t n, void* stream);
/* hist[seg][bin] += #elements of segment seg whose fp16(|acc|) bit pattern == bin; bins = 32768
 * (patterns 0x0000..0x7FFF; |x| clears the sign).  hist: uint32 [n_segs][32768], caller-zeroed.
 * seg_offsets_dev: device copy of the offsets.  (generate_mask.py:47-49 abs + fp16 cast) */
int clear_saliency_histogram(const void* acc, const int64_t* seg_offsets_dev, int32_t n_segs, int64_t n_total,
                             uint32_t* hist, void* stream);
/* For each segment and each of n_th thresholds: k = (int64)(th * numel) (double product truncated,
 * as Python int()); finds tau = bit pattern of the k-th largest value and the number of ties to
 * take.  select: int64 [n_segs][n_th][5] = (tau_bits or -1 when k == 0, ties_to_take, lo, hi, total) with
 * [lo, hi] the bf16 magnitude bit patterns whose key is tau and total = #{key == tau} in the tensor
 * (both used by the mask pass).
 * seg_numel_dev: optional device int64 [n_segs] true element counts for tensors stored zero-padded
 * to the next multiple of 8 (NULL: numel = offset difference).  thresholds: host doubles. */
int clear_saliency_select(const uint32_t* hist, const int64_t* seg_offsets_dev, const int64_t* seg_numel_dev,
                          int32_t n_segs, const double* thresholds_host, int32_t n_th, int64_t* select, void* stream);
/* mask[i] = 1 iff fp16(|acc[i]|) > tau, or == tau and the element is among the first `ties` such
 * elements of its tensor in flat index order (stable descending sort; SURVEY.md §8c tie contract).
 * th_index selects which threshold's (tau, ties) of `select` to use.
 * Work is cut into chunks of clear_saliency_chunk_elems() elements that never straddle tensors:
 * chunk_prefix_dev: device int64 [n_segs+1], chunk_prefix[s] = sum_{t<s} ceil(numel_t / chunk);
 * n_chunks = chunk_prefix[n_segs]; tie_scratch: uint64 [n_chunks] workspace.
 * One streaming pass writes the mask speculating on the tie boundary from the histogram; a scan of the
 * chunk tie counts and an audit pass rewrite the few chunks that guessed wrong (result is exact). */
/* Global-threshold scope (upstream SalUn's global rank, of which the reference's `start_index` at generate_mask.py:59,83 is
 * the vestige): ONE ranking over all tensors, k = (int64)(th * n_total), ties by global flat index.
 *   clear_saliency_hist_total         total[bin] = sum over tensors of hist (zero padding excluded); int64 [32768].  With the
 *                                     flat buffer sharded over ranks the caller all-reduces `total` (256 KB) at this point.
 *   clear_saliency_select_global_tau  tau_budget int64 [n_th][2] = (threshold key or -1 when k == 0, tie budget).
 *   clear_saliency_select_global_fill the per-tensor records clear_saliency_select writes, for the common key: tensor s takes
 *                                     clamp(budget - #ties in earlier tensors, 0, #ties in s) ties; budget_override (device
 *                                     int64 [n_th], optional) replaces the budget — this rank's share of it.
 * The mask entry points consume the records unchanged. */
int clear_saliency_hist_total(const uint32_t* hist, const int64_t* seg_offsets_dev, const int64_t* seg_numel_dev, int32_t n_segs,
                              int64_t* total, void* stream);
int clear_saliency_select_global_tau(const int64_t* total, int64_t n_total, const double* thresholds_host, int32_t n_th,
                                     int64_t* tau_budget, void* stream);
int clear_saliency_select_global_fill(const uint32_t* hist, const int64_t* seg_offsets_dev, const int64_t* seg_numel_dev,
                                      int32_t n_segs, const int64_t* tau_budget, const int64_t* budget_override, int32_t n_th,
                                      int64_t* select, void* stream);
int64_t clear_saliency_chunk_elems(void);
int clear_saliency_mask(const void* acc, const int64_t* seg_offsets_dev, const int64_t* chunk_prefix_dev,
                        int32_t n_segs, int64_t n_chunks, const int64_t* select, int32_t n_th, int32_t th_index,
                        uint64_t* tie_scratch, int8_t* mask, void* stream);
/* Same selection, bit-packed output: bit (i & 7) of byte i >> 3 is the mask of flat element i (layout offsets are
 * multiples of 8, so tensors start on byte boundaries).  1/8 of the int8 format's traffic: 2.125 instead of 3 bytes
 * per parameter and threshold here, 22.125 instead of 23 in the masked update. */
int clear_saliency_mask_bits(const void* acc, const int64_t* seg_offsets_dev, const int64_t* chunk_prefix_dev,
                             int32_t n_segs, int64_t n_chunks, const int64_t* select, int32_t n_th, int32_t th_index,
                             uint64_t* tie_scratch, uint8_t* mask_bits, void* stream);
/* Converters between the reference's on-disk mask format (`with_{th}.pt`: int8, one byte per parameter,
 * generate_mask.py:72-89) and the bit-packed one.  n elements; mask_bits holds (n + 7) / 8 bytes. */
int clear_mask_pack(const int8_t* mask, int64_t n, uint8_t* mask_bits, void* stream);
int clear_mask_unpack(const uint8_t* mask_bits, int64_t n, int8_t* mask, void* stream);
/* Masked AdamW: g' = g * mask; torch _single_tensor_adam op order, fp32 states.
 * param, grad bf16; exp_avg, exp_avg_sq fp32; mask int8 or mask_bits bit-packed (both NULL = all ones). step >= 1. */
typedef struct ClearAdamWArgs {
  void* param;
  const void* grad;
  float* exp_avg;
  float* exp_avg_sq;
  const int8_t* mask;
  int64_t n;
  double lr, beta1, beta2, eps, weight_decay; /* doubles: the scalar prologue runs in double like torch's */
  int32_t step;
  const float* grad_scale_dev; /* optional device scalar (gradient-clipping coefficient): g = bf16(g * scale) first,
                                  like torch's in-place `grad.mul_(clip_coef)` on a bf16 gradient */
  const uint8_t* mask_bits;    /* optional bit-packed mask (clear_saliency_mask_bits / clear_mask_pack); excludes `mask` */
} ClearAdamWArgs;
int clear_masked_adamw(const ClearAdamWArgs* args, void* stream);

/* Masked AdamW with 8-bit block-wise states.  The reference trains with optim="adamw_bnb_8bit"
 * (mm/forget.py:156, nlp/forget.py:166 -> bitsandbytes==0.43.3, absent here: PARITY UNPINNED; the oracle restates what this
 * computes).  Both moments are one byte per parameter: an index into a 256-entry sorted codebook (code1 signed for exp_avg,
 * code2 non-negative for exp_avg_sq; host-generated "dynamic" maps) times one fp32 absmax per block of
 * clear_adamw8_block_elems() = 256 parameters (absmax1 / absmax2: ceil(n / 256) floats each).  One fused pass: dequantise ->
 * the update of clear_masked_adamw (same operation order) -> new block absmax -> requantise.  Quantisation of x: n = x * (1 /
 * absmax) (IEEE reciprocal, one multiply; 0 for an all-zero block), key = (bits(n) + 0x8000) >> 16 (n rounded to sign, exponent
 * and 7 mantissa bits), index = qlut[key].  qlut1 (65536 bytes, every 16-bit pattern) / qlut2 (32768 bytes, the non-negative
 * ones): the index of the code value nearest to the float whose top 16 bits are the key (host-generated from the codebooks:
 * #{fp32 midpoints of neighbouring code values that are < it}).  States start as the index of 0.0 with absmax 0. */
typedef struct ClearAdamW8Args {
  void* param;            /* bf16 [n] */
  const void* grad;       /* bf16 [n] */
  uint8_t* state1;        /* [n] code indices of exp_avg */
  uint8_t* state2;        /* [n] code indices of exp_avg_sq */
  float* absmax1;         /* [ceil(n / 256)] */
  float* absmax2;
  const float* code1;     /* device fp32 [256], ascending */
  const float* code2;
  const uint8_t* qlut1;   /* device uint8 [65536] */
  const uint8_t* qlut2;   /* device uint8 [32768] */
  const int8_t* mask;     /* optional int8 [n] ... */
  const uint8_t* mask_bits; /* ... or bit-packed; both NULL = all ones */
  int64_t n;
  double lr, beta1, beta2, eps, weight_decay;
  int32_t step;
  const float* grad_scale_dev;
} ClearAdamW8Args;
int32_t clear_adamw8_block_elems(void);
int clear_masked_adamw_8bit(const ClearAdamW8Args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-model reductions around the update (one pass over HBM each, bitwise reproducible).
 * partials: double workspace of clear_reduce_partials_elems() elements; out: device doubles;
 * accumulate != 0 adds to *out (tensor-by-tensor use), else overwrites.
 *   clear_grad_sqnorm: out[0] (+)= sum (grad * mask)^2 (mask int8 or bit-packed or none) — clip_grad_norm_ (HF Trainer's
 *                      default max_grad_norm = 1.0; the override is commented out at mm/forget.py:151)
 *   clear_clip_coef:   coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))
 *   clear_param_l1l0:  out[0] (+)= sum |p - ref|, out[1] (+)= #{p != ref}; ref NULL = zeros
 *                      — torch.norm(torch.cat(params), p=1 / p=0), mm/trainer.py:451-481 (A9)
 * ------------------------------------------------------------------------------------------- */
int64_t clear_reduce_partials_elems(void);
int clear_grad_sqnorm(const void* grad, const int8_t* mask, const uint8_t* mask_bits, int64_t n, double* partials,
                      double* out, int accumulate, void* stream);
int clear_clip_coef(const double* sumsq, double max_norm, float* coef, void* stream);
/* grad[i] = bf16(grad[i] * *coef) in place (torch: grads.mul_(clip_coef_clamped)); the fused alternative is
 * ClearAdamWArgs.grad_scale_dev */
int clear_grad_scale(void* grad, int64_t n, const float* coef, void* stream);
int clear_param_l1l0(const void* param, const void* ref, int64_t n, double* partials, double* out2, int accumulate,
                     void* stream);
/* out[i] = bf16(*coef * sign(param[i] - ref[i])): backward of coef * ||param - ref||_1 (the L0 term has zero gradient) */
int clear_param_l1_grad(const void* param, const void* ref, int64_t n, const float* coef, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLEAR_B200_H_ */
