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
  printf(failures ? "FAILED %d\n" : "ok\n", failures);
  return failures;
}
