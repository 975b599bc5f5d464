// One-call loss head: clear_head_loss_workspace_bytes / clear_head_loss_fwd_bwd.
//
// Replaces, for one `model(**split)` call of the reference trainers, everything between "the trunk produced hidden
// states" and "autograd filled dH and lm_head.weight.grad" (mm/trainer.py:168-449, nlp/dataloader.py:144-459): rows are
// streamed in blocks of `block_rows`; per block: student projection (+ LSE partials) -> teacher projection or LSE
// partials of materialised teacher logits -> row pass (loss terms, dlogits in place) -> dH and dW contractions; then
// the fixed-order loss reduction.  The caller owns every buffer including the workspace; all launches go to one
// stream; nothing is allocated and nothing synchronises.  The normalisations are the reference's:
// CE mean over non-ignored rows, KL over all rows / N, mm NPO (2 / beta) x mean over N V, entropy raw sum,
// get_batch_loss per-sequence sums averaged over B.
//
// Row maps.  Two things select WHICH rows of the trunk's output the head evaluates, and both are kernel inputs here:
//   * `row_index` — `remove_image_tokens` (mm/trainer_utils.py:48-70; used four times per `dpo` step, mm/trainer.py:373-387)
//     as an index over hidden rows instead of a copy of [B, L, V] logits;
//   * `compact_ignored` — rows whose shifted label is -100 (padding, the 576 image positions, the question) contribute
//     nothing to a CE-type loss or to any gradient: valid rows are compacted ON THE DEVICE (clear_head_compact_rows) and
//     their count stays there — the contractions, the row pass, the gathers and the scatters all read it from device
//     memory (GemmParams::dyn_count), so the step has no host synchronisation.
// Both reduce to: gather the selected rows of hidden (zero-padded to the next multiple of 64 rows: the dW contraction's K
// block) -> run the block loop on the gathered rows -> scatter dH rows (others zero) and the loss terms back.
#include "common.h"

namespace clear {
namespace {

constexpr int64_t kDefaultBlockRows = 8192;
constexpr size_t kAlign = 256;

size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct Plan {
  int64_t blk, n_blocks, ld;
  int nt;
  bool needs_ref, project_teacher, write_grad, want_dh_ws, compact, indexed;
  size_t off_z, off_part, off_zr, off_rpart, off_inv, off_dhws;
  size_t off_pos, off_src, off_cnt, off_hg, off_rhg, off_lab, off_rw, off_lrows, off_dhg, total;
  float loss_scale, grad_scale;
  bool per_valid;
};

int make_plan(const ClearHeadArgs* a, Plan* p) {
  if (!a) return fail(CLEAR_ERR_NULL, "clear_head_loss: args is NULL");
  if (!a->hidden || !a->weight || !a->loss || !a->loss_rows) return fail(CLEAR_ERR_NULL, "clear_head_loss: hidden, weight, loss and loss_rows are required");
  if (a->n_rows <= 0 || a->n_rows > INT32_MAX || a->hidden_dim <= 0 || a->vocab <= 0)
    return fail(CLEAR_ERR_SHAPE, "clear_head_loss: bad sizes n_rows=%lld hidden=%d vocab=%d", (long long)a->n_rows, a->hidden_dim, a->vocab);
  if (a->kind < CLEAR_KIND_CE || a->kind > CLEAR_KIND_ENTROPY) return fail(CLEAR_ERR_KIND, "clear_head_loss: unknown kind %d", a->kind);
  p->needs_ref = a->kind == CLEAR_KIND_KL_FWD || a->kind == CLEAR_KIND_KL_REV || a->kind == CLEAR_KIND_NPO;
  p->project_teacher = p->needs_ref && !a->ref_logits;
  if (p->needs_ref && !a->ref_logits && (!a->ref_hidden || !a->ref_weight))
    return fail(CLEAR_ERR_NULL, "clear_head_loss: kind %d needs the frozen teacher: ref_hidden + ref_weight, or ref_logits", a->kind);
  if (a->kind == CLEAR_KIND_CE && !a->labels) return fail(CLEAR_ERR_NULL, "clear_head_loss: CE needs labels");
  if (a->kind == CLEAR_KIND_NPO && !(a->beta > 0.0)) return fail(CLEAR_ERR_SHAPE, "clear_head_loss: NPO beta must be > 0");
  if (a->n_peers < 0 || a->n_peers > 8) return fail(CLEAR_ERR_SHAPE, "clear_head_loss: 0 <= n_peers <= 8");
  if (a->n_peers > 0 && a->dweight) return fail(CLEAR_ERR_KIND, "clear_head_loss: pass dweight or dweight_peers, not both");
  p->compact = a->compact_ignored != 0;
  if (p->compact && a->kind != CLEAR_KIND_CE)
    return fail(CLEAR_ERR_KIND, "clear_head_loss: compact_ignored is defined for CE only (KL / NPO / entropy count every row in the reference)");
  p->indexed = p->compact || a->row_index;
  if (a->row_index && a->n_src_rows <= 0) return fail(CLEAR_ERR_SHAPE, "clear_head_loss: row_index needs n_src_rows (rows of hidden / dhidden)");
  if (p->indexed && (a->hidden_dim & 1)) return fail(CLEAR_ERR_SHAPE, "clear_head_loss: row maps need an even hidden_dim");
  p->write_grad = a->dhidden || a->dweight || a->n_peers > 0;
  p->blk = a->block_rows > 0 ? a->block_rows : kDefaultBlockRows;
  if (p->blk > a->n_rows) p->blk = a->n_rows;
  if (p->compact && p->blk < a->n_rows && (p->blk & 63))
    return fail(CLEAR_ERR_SHAPE, "clear_head_loss: with compact_ignored, block_rows must be a multiple of 64");
  p->n_blocks = (a->n_rows + p->blk - 1) / p->blk;
  p->ld = ((int64_t)a->vocab + 63) / 64 * 64;
  p->nt = clear_head_vtiles(a->vocab);
  p->want_dh_ws = a->dhidden && (a->hidden_dim % 8 == 0);

  // reference normalisations: (factor on the summed row terms, factor on the kernel's dterm, 1/#valid on top)
  const double n = (double)a->n_rows, v = (double)a->vocab, beta = a->beta;
  const double gfac = a->kind == CLEAR_KIND_NPO ? beta : 1.0;  // NPO's dterm is the derivative of (1/beta) sum log(1 + t)
  double ls, gs;
  p->per_valid = false;
  switch (a->reduction) {
    case CLEAR_REDUCE_SUM: ls = 1.0; gs = gfac; break;
    case CLEAR_REDUCE_SEQ_MEAN:
      if (a->n_seqs <= 0 || a->n_rows % a->n_seqs) return fail(CLEAR_ERR_SHAPE, "clear_head_loss: seq_mean needs n_seqs dividing n_rows");
      ls = 1.0 / a->n_seqs; gs = gfac / a->n_seqs; break;
    case CLEAR_REDUCE_REFERENCE:
      switch (a->kind) {
        case CLEAR_KIND_CE: ls = 1.0; gs = 1.0; p->per_valid = true; break;
        case CLEAR_KIND_KL_FWD: case CLEAR_KIND_KL_REV: ls = 1.0 / n; gs = 1.0 / n; break;
        case CLEAR_KIND_NPO: ls = 2.0 / (beta * n * v); gs = 2.0 / (n * v); break;
        default: ls = 1.0; gs = 1.0; break;  // entropy: raw sum
      }
      break;
    default: return fail(CLEAR_ERR_KIND, "clear_head_loss: unknown reduction %d", a->reduction);
  }
  if (a->seq_loss && (a->n_seqs <= 0 || a->n_rows % a->n_seqs)) return fail(CLEAR_ERR_SHAPE, "clear_head_loss: seq_loss needs n_seqs dividing n_rows");
  p->loss_scale = (float)(a->coef * ls);
  p->grad_scale = (float)(a->coef * gs);

  const size_t nr = (size_t)a->n_rows, hb = (size_t)a->hidden_dim * 2;
  size_t off = 0;
  p->off_z = off;     off += align_up((size_t)p->blk * p->ld * 2);
  p->off_part = off;  off += align_up((size_t)p->blk * p->nt * 8);
  p->off_zr = off;    if (p->project_teacher) off += align_up((size_t)p->blk * p->ld * 2);
  p->off_rpart = off; if (p->needs_ref) off += align_up((size_t)p->blk * p->nt * 8);
  p->off_inv = off;   off += kAlign;
  p->off_dhws = off;  if (p->want_dh_ws) off += align_up((size_t)p->blk * a->hidden_dim * 4);
  p->off_pos = off;   if (p->compact) off += align_up(nr * 4);
  p->off_src = off;   if (p->compact && a->row_index) off += align_up(nr * 4);
  p->off_cnt = off;   if (p->compact) off += kAlign;
  p->off_hg = off;    if (p->indexed) off += align_up(nr * hb);
  p->off_rhg = off;   if (p->indexed && p->project_teacher) off += align_up(nr * hb);
  p->off_lab = off;   if (p->compact) off += align_up(nr * 8);
  p->off_rw = off;    if (p->compact && a->row_weight) off += align_up(nr * 4);
  p->off_lrows = off; if (p->compact) off += align_up(nr * 4);
  p->off_dhg = off;   if (p->indexed && a->dhidden) off += align_up(nr * hb);
  p->total = off;
  return 0;
}

}  // namespace
}  // namespace clear

using namespace clear;

extern "C" {

int clear_head_loss_workspace_bytes(const ClearHeadArgs* a, size_t* out) {
  if (!out) return fail(CLEAR_ERR_NULL, "clear_head_loss_workspace_bytes: out is NULL");
  Plan p;
  if (int rc = make_plan(a, &p)) return rc;
  *out = p.total;
  return 0;
}

int clear_head_loss_fwd_bwd(const ClearHeadArgs* a, void* workspace, size_t ws_bytes, void* stream) {
  Plan p;
  if (int rc = make_plan(a, &p)) return rc;
  if (!workspace) return fail(CLEAR_ERR_NULL, "clear_head_loss_fwd_bwd: workspace is NULL");
  if (ws_bytes < p.total) return fail(CLEAR_ERR_WORKSPACE, "clear_head_loss_fwd_bwd: workspace of %zu bytes, %zu needed", ws_bytes, p.total);
  if (reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) return fail(CLEAR_ERR_ALIGN, "clear_head_loss_fwd_bwd: workspace must be 256-byte aligned");
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void* z = ws + p.off_z;
  float* part = reinterpret_cast<float*>(ws + p.off_part);
  void* zr = p.project_teacher ? ws + p.off_zr : nullptr;
  float* rpart = p.needs_ref ? reinterpret_cast<float*>(ws + p.off_rpart) : nullptr;
  float* inv_count = p.per_valid ? reinterpret_cast<float*>(ws + p.off_inv) : nullptr;
  float* dh_ws = p.want_dh_ws ? reinterpret_cast<float*>(ws + p.off_dhws) : nullptr;
  const int esz = 2;
  const int64_t H = a->hidden_dim;
  const int32_t hbytes = (int32_t)(H * esz);
  int rc;

  if (p.per_valid && (rc = clear_head_count_valid(a->labels, a->n_rows, inv_count, stream))) return rc;

  // ---- row maps: which rows of hidden are evaluated, and where their results go ----
  const void* hidden = a->hidden;
  const void* ref_hidden = a->ref_hidden;
  int64_t hidden_ld = a->hidden_ld, ref_hidden_ld = a->ref_hidden_ld;
  const int64_t* labels = a->labels;
  const float* row_weight = a->row_weight;
  float* loss_rows = a->loss_rows;
  void* dhidden = a->dhidden;
  int64_t dhidden_ld = a->dhidden_ld;
  const int32_t* pos = nullptr;     // compacted row -> row of this invocation (labels / loss_rows / row_weight)
  const int32_t* src = a->row_index;  // evaluated row -> row of hidden / dhidden
  const int32_t* count_dev = nullptr;
  if (p.compact) {
    int32_t* pos_w = reinterpret_cast<int32_t*>(ws + p.off_pos);
    int32_t* cnt_w = reinterpret_cast<int32_t*>(ws + p.off_cnt);
    if ((rc = clear_head_compact_rows(a->labels, a->n_rows, pos_w, cnt_w, stream))) return rc;
    pos = pos_w;
    count_dev = cnt_w;
    if (a->row_index) {
      int32_t* src_w = reinterpret_cast<int32_t*>(ws + p.off_src);
      if ((rc = clear_index_compose(pos_w, a->row_index, a->n_rows, cnt_w, src_w, stream))) return rc;
      src = src_w;
    } else {
      src = pos_w;
    }
    int64_t* lab_g = reinterpret_cast<int64_t*>(ws + p.off_lab);
    if ((rc = clear_rows_gather_dev(a->labels, 8, pos, a->n_rows, count_dev, 0, 8, lab_g, 8, stream))) return rc;
    labels = lab_g;
    if (a->row_weight) {
      float* rw_g = reinterpret_cast<float*>(ws + p.off_rw);
      if ((rc = clear_rows_gather_dev(a->row_weight, 4, pos, a->n_rows, count_dev, 0, 4, rw_g, 4, stream))) return rc;
      row_weight = rw_g;
    }
    loss_rows = reinterpret_cast<float*>(ws + p.off_lrows);
  }
  if (p.indexed) {
    void* hg = ws + p.off_hg;
    if ((rc = clear_rows_gather_dev(a->hidden, a->hidden_ld * esz, src, a->n_rows, count_dev, 64, hbytes, hg, hbytes, stream))) return rc;
    hidden = hg;
    hidden_ld = H;
    if (p.project_teacher) {
      void* rhg = ws + p.off_rhg;
      if ((rc = clear_rows_gather_dev(a->ref_hidden, a->ref_hidden_ld * esz, src, a->n_rows, count_dev, 64, hbytes, rhg, hbytes, stream))) return rc;
      ref_hidden = rhg;
      ref_hidden_ld = H;
    }
    if (a->dhidden) {
      dhidden = ws + p.off_dhg;
      dhidden_ld = H;
    }
  }

  const int ref_esz = a->ref_logits_f32 ? 4 : 2;
  auto stage_event = [&](int64_t b, int k) -> int {   // optional profiling hooks (ClearHeadArgs::stage_events)
    const int64_t i = 4 * b + k;
    if (a->stage_events && i < a->n_stage_events && a->stage_events[i])
      CLEAR_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(a->stage_events[i]), cs));
    return 0;
  };
  for (int64_t b = 0; b < p.n_blocks; ++b) {
    const int64_t r0 = b * p.blk, rows = (r0 + p.blk <= a->n_rows ? p.blk : a->n_rows - r0);
    ClearProjectArgs pa = {};
    pa.hidden = static_cast<const uint8_t*>(hidden) + r0 * hidden_ld * esz;
    pa.weight = a->weight;
    pa.logits = z;
    pa.lse_part = part;
    pa.hidden_ld = hidden_ld; pa.weight_ld = a->weight_ld; pa.logits_ld = p.ld;
    pa.n_rows = rows; pa.hidden_dim = a->hidden_dim; pa.vocab = a->vocab; pa.cta_group = a->cta_group;
    pa.n_rows_dev = count_dev; pa.n_rows_dev_offset = (int32_t)r0;
    if ((rc = stage_event(b, 0))) return rc;

    const void* ref_rows = nullptr;
    int64_t ref_ld = 0;
    int ref_f32 = 0;
    if (p.project_teacher) {   // student and teacher projections: two independent contractions, one launch
      ClearProjectArgs ta = pa;
      ta.hidden = static_cast<const uint8_t*>(ref_hidden) + r0 * ref_hidden_ld * esz;
      ta.weight = a->ref_weight;
      ta.logits = zr;
      ta.lse_part = rpart;
      ta.hidden_ld = ref_hidden_ld; ta.weight_ld = a->ref_weight_ld;
      if ((rc = clear_head_project_pair(&pa, &ta, stream))) return rc;
      ref_rows = zr;
      ref_ld = p.ld;
    } else if ((rc = clear_head_project(&pa, stream))) {
      return rc;
    }
    if (!p.project_teacher && p.needs_ref) {
      ref_rows = static_cast<const uint8_t*>(a->ref_logits) + r0 * a->ref_logits_ld * ref_esz;
      ref_ld = a->ref_logits_ld;
      ref_f32 = a->ref_logits_f32 ? 1 : 0;
      if (ref_f32) rc = clear_head_lse_partials_f32(ref_rows, ref_ld, rows, a->vocab, rpart, stream);
      else         rc = clear_head_lse_partials(ref_rows, ref_ld, rows, a->vocab, rpart, stream);
      if (rc) return rc;
    }

    if ((rc = stage_event(b, 1))) return rc;
    ClearRowLossArgs ra = {};
    ra.logits = z;
    ra.lse_part = part;
    ra.ref_logits = ref_rows;
    ra.ref_lse_part = rpart;
    ra.labels = (a->kind == CLEAR_KIND_CE) ? labels + r0 : nullptr;
    ra.row_weight = row_weight ? row_weight + r0 : nullptr;
    ra.scale_dev = inv_count;
    ra.loss_rows = loss_rows + r0;
    ra.logits_ld = p.ld; ra.ref_ld = ref_ld;
    ra.n_rows = rows; ra.vocab = a->vocab; ra.kind = a->kind;
    ra.grad_scale = p.grad_scale; ra.beta = (float)a->beta; ra.write_grad = p.write_grad ? 1 : 0;
    ra.ref_logits_f32 = ref_f32;
    ra.n_rows_dev = count_dev; ra.n_rows_dev_offset = (int32_t)r0;
    ra.status_flags = a->status_flags;
    if ((rc = clear_head_rowloss(&ra, stream))) return rc;
    if ((rc = stage_event(b, 2))) return rc;

    if (p.write_grad) {
      if (b == 0 && a->n_peers > 0 && a->dweight_ready_event)
        CLEAR_CUDA_OK(cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(a->dweight_ready_event), 0));
      ClearBackwardArgs ba = {};
      ba.dlogits = z;
      ba.hidden = static_cast<const uint8_t*>(hidden) + r0 * hidden_ld * esz;
      ba.weight = a->weight;
      ba.dhidden = dhidden ? static_cast<uint8_t*>(dhidden) + r0 * dhidden_ld * esz : nullptr;
      ba.dweight = a->dweight;
      ba.dlogits_ld = p.ld; ba.hidden_ld = hidden_ld; ba.weight_ld = a->weight_ld;
      ba.dhidden_ld = dhidden_ld; ba.dweight_ld = a->dweight_ld > 0 ? a->dweight_ld : a->hidden_dim;
      ba.n_rows = rows; ba.hidden_dim = a->hidden_dim; ba.vocab = a->vocab;
      ba.dweight_f32 = a->dweight_f32;
      ba.accumulate = (a->dweight_accumulate || b > 0) ? 1 : 0;
      ba.cta_group = a->cta_group;
      ba.dhidden_ws = dh_ws;
      ba.n_peers = a->n_peers;
      for (int r = 0; r < a->n_peers; ++r) ba.dweight_peers[r] = a->dweight_peers[r];
      ba.n_rows_dev = count_dev; ba.n_rows_dev_offset = (int32_t)r0;
      if ((rc = clear_head_backward(&ba, stream))) return rc;
    }
    if ((rc = stage_event(b, 3))) return rc;
  }

  // ---- results back to the caller's row space ----
  if (p.compact) {
    CLEAR_CUDA_OK(cudaMemsetAsync(a->loss_rows, 0, (size_t)a->n_rows * sizeof(float), cs));
    if ((rc = clear_rows_scatter_dev(loss_rows, 4, pos, a->n_rows, count_dev, 4, a->loss_rows, 4, stream))) return rc;
  }
  if (p.indexed && a->dhidden) {
    const int64_t out_rows = a->row_index ? a->n_src_rows : a->n_rows;
    CLEAR_CUDA_OK(cudaMemset2DAsync(a->dhidden, (size_t)a->dhidden_ld * esz, 0, (size_t)hbytes, (size_t)out_rows, cs));
    if ((rc = clear_rows_scatter_dev(dhidden, hbytes, src, a->n_rows, count_dev, hbytes, a->dhidden, a->dhidden_ld * esz, stream))) return rc;
  }
  if ((rc = clear_head_reduce_rows(a->loss_rows, a->n_rows, p.loss_scale, inv_count, a->loss, 0, stream))) return rc;
  if (a->seq_loss && (rc = clear_head_reduce_sequences(a->loss_rows, a->n_seqs, a->n_rows / a->n_seqs, a->seq_loss, stream))) return rc;
  return 0;
}

}  // extern "C"
