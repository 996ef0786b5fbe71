/* The C-ABI header must be plain C (a Rust bindgen / cgo / JNI binding parses it as C): compiled with gcc -std=c99 -pedantic by
 * tests/test_abi.py, linked against libh2sha_b200.so, and run without a GPU (plan-only engine: layout and handle queries). */
#include <stdio.h>
#include <string.h>

#include "../../include/h2sha_b200.h"

int main(void) {
  uint32_t sizes[2] = {128, 128};
  h2sha_config_t cfg;
  h2sha_engine_t* e = NULL;
  h2sha_layout_t lay;
  h2sha_lookup_info_t li;
  uint32_t input_len = 0, out_bytes[32], in_bytes[128];
  memset(&cfg, 0, sizeof cfg);
  cfg.n_digests = 2;
  cfg.max_variable_byte_sizes = sizes;
  cfg.is_input_range_check = 1;
  cfg.device = -1; /* plan-only: no GPU needed */
  if (h2sha_create(&cfg, &e) != H2SHA_OK) { printf("create: %s\n", h2sha_last_error()); return 1; }
  if (h2sha_get_layout(e, &lay) != H2SHA_OK || h2sha_get_lookup_info(e, &li) != H2SHA_OK) return 2;
  if (h2sha_get_handles(e, 1, &input_len, in_bytes, out_bytes) != H2SHA_OK) return 3;
  /* the reference's test circuit: 4 blocks in 3 advice columns (lib.rs:487-494) */
  printf("ok blocks=%u gate_cols=%u cells=%llu range_rows=%u\n", lay.n_blocks, lay.n_gate_cols, (unsigned long long)lay.cells_per_instance, li.range_table_rows);
  {
    h2sha_batch_t b;
    memset(&b, 0, sizeof b);
    b.n_instances = 1;
    if (h2sha_digest_batch(e, &b) != H2SHA_ECUDA) return 4; /* no CPU path */
  }
  h2sha_destroy(e);
  return (lay.n_blocks == 4 && lay.n_gate_cols == 3) ? 0 : 5;
}
