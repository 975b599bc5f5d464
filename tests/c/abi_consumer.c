/* A plain C consumer of include/clear_b200.h: proves the header is valid C, that the library links without Python or
 * torch, and that argument errors are reported before any launch (no GPU is needed for what this program calls). */
#include <stdio.h>
#include <string.h>

#include "clear_b200.h"

int main(void) {
  int failures = 0;
  if (clear_abi_version() != CLEAR_ABI_VERSION) { printf("abi version mismatch\n"); ++failures; }

  ClearProjectArgs p;
  memset(&p, 0, sizeof p);
  if (clear_head_project(&p, NULL) != CLEAR_ERR_NULL) { printf("project: expected CLEAR_ERR_NULL\n"); ++failures; }
  if (strlen(clear_last_error()) == 0) { printf("empty error string\n"); ++failures; }

  ClearRowLossArgs r;
  memset(&r, 0, sizeof r);
  if (clear_head_rowloss(&r, NULL) >= 0) { printf("rowloss: expected an argument error\n"); ++failures; }

  ClearBackwardArgs b;
  memset(&b, 0, sizeof b);
  if (clear_head_backward(&b, NULL) >= 0) { printf("backward: expected an argument error\n"); ++failures; }

  ClearAdamWArgs a;
  memset(&a, 0, sizeof a);
  if (clear_masked_adamw(&a, NULL) >= 0) { printf("adamw: expected an argument error\n"); ++failures; }

  int32_t r0 = -1, r1 = -1;
  if (clear_head_dweight_slab(32064, 8, 7, 0, &r0, &r1) != 0 || r1 != 32064 || r0 % 256 != 0) {
    printf("dweight slab: got [%d, %d)\n", (int)r0, (int)r1);
    ++failures;
  }
  if (clear_head_vtiles(32064) != 126) { printf("vtiles\n"); ++failures; }
  if (clear_saliency_chunk_elems() <= 0) { printf("chunk elems\n"); ++failures; }
  if (clear_mask_pack(NULL, 8, NULL, NULL) != CLEAR_ERR_NULL) { printf("mask_pack\n"); ++failures; }
  if (clear_rows_gather(NULL, 0, NULL, 0, 4, NULL, 0, NULL) != CLEAR_ERR_NULL) { printf("rows_gather\n"); ++failures; }
  /* ---- ABI 3 ---- */
  {
    static float dummy[4];      /* never dereferenced: every call below fails its argument check before any launch */
    static int64_t labels[4];
    size_t need = 0;
    ClearHeadArgs h;
    memset(&h, 0, sizeof h);
    h.hidden = dummy; h.weight = dummy; h.loss = dummy; h.loss_rows = dummy; h.labels = labels;
    h.n_rows = 4; h.hidden_dim = 8; h.vocab = 16; h.kind = CLEAR_KIND_CE; h.beta = 1.0; h.coef = 1.0;
    h.compact_ignored = 1;
    if (clear_head_loss_workspace_bytes(&h, &need) != 0 || need == 0) { printf("head workspace (compact_ignored)\n"); ++failures; }
    size_t plain = 0;
    h.compact_ignored = 0;
    if (clear_head_loss_workspace_bytes(&h, &plain) != 0 || plain >= need) { printf("row compaction needs its own workspace\n"); ++failures; }
    h.kind = CLEAR_KIND_ENTROPY; h.compact_ignored = 1;
    if (clear_head_loss_workspace_bytes(&h, &need) != CLEAR_ERR_KIND) { printf("compact_ignored is CE only\n"); ++failures; }
    h.kind = CLEAR_KIND_CE; h.compact_ignored = 0;
    h.row_index = (const int32_t*)labels; h.n_src_rows = 0;
    if (clear_head_loss_workspace_bytes(&h, &need) != CLEAR_ERR_SHAPE) { printf("row_index needs n_src_rows\n"); ++failures; }
    h.n_src_rows = 9; h.ref_logits_f32 = 1; h.status_flags = NULL; h.dweight_ready_event = NULL;
    if (clear_head_loss_workspace_bytes(&h, &need) != 0) { printf("row_index: %s\n", clear_last_error()); ++failures; }
  }
  if (clear_rows_gather_dev(NULL, 0, NULL, 0, NULL, 64, 4, NULL, 0, NULL) != CLEAR_ERR_NULL) { printf("rows_gather_dev\n"); ++failures; }
  if (clear_rows_scatter_dev(NULL, 0, NULL, 0, NULL, 4, NULL, 0, NULL) != CLEAR_ERR_NULL) { printf("rows_scatter_dev\n"); ++failures; }
  if (clear_index_compose(NULL, NULL, 4, NULL, NULL, NULL) != CLEAR_ERR_NULL) { printf("index_compose\n"); ++failures; }
  if (clear_head_lse_partials_f32(NULL, 8, 1, 8, NULL, NULL) != CLEAR_ERR_NULL) { printf("lse_partials_f32\n"); ++failures; }
  if (clear_saliency_hist_total(NULL, NULL, NULL, 1, NULL, NULL) != CLEAR_ERR_NULL) { printf("hist_total\n"); ++failures; }
  if (clear_saliency_select_global_tau(NULL, 1, NULL, 1, NULL, NULL) != CLEAR_ERR_NULL) { printf("select_global_tau\n"); ++failures; }
  if (clear_saliency_select_global_fill(NULL, NULL, NULL, 1, NULL, NULL, 1, NULL, NULL) != CLEAR_ERR_NULL) { printf("select_global_fill\n"); ++failures; }
  {
    ClearAdamW8Args a8;
    memset(&a8, 0, sizeof a8);
    if (clear_masked_adamw_8bit(&a8, NULL) != CLEAR_ERR_NULL) { printf("adamw 8-bit: expected CLEAR_ERR_NULL\n"); ++failures; }
    if (clear_adamw8_block_elems() != 256) { printf("adamw 8-bit block\n"); ++failures; }
  }
  printf(failures ? "FAILED %d\n" : "ok\n", failures);
  return failures;
}
