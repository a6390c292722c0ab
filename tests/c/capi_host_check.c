/* capi_host_check.c -- the C-ABI from plain C (the language of the reference's FFI layer).
 *
 * Compiles include/pflare_b200.h as C99, links libpflare_b200.so, and drives a host-only PLANNING context
 * (device = -1, no GPU needed) through the upload hook exactly as the PETSc-side shim would: a hand-written
 * 2-level hierarchy of the 1D upwind problem of tests/adv_1d.c (n = 6: C points 0,2,4 / F points 1,3,5).
 * Checks return codes, the error-string convention, the bit-exact round trip of the integer data, the work
 * model counters, and that every compute entry point refuses to run without a device (no CPU fallback).
 * Prints CAPI_HOST_OK.
 */
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include "pflare_b200.h"

#define REQUIRE(cond)                                                                 \
  do {                                                                                \
    if (!(cond)) { printf("FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, pflare_b200_last_error()); return 1; } \
  } while (0)

int main(void)
{
  void *h = NULL;
  /* level 1: A = [1; -1 1; ...] (6x6), F = {1,3,5}, C = {0,2,4} */
  const int is_f[3] = {1, 3, 5}, is_c[3] = {0, 2, 4}, smooth[1] = {2};
  /* A_ff = I (F points only couple to C points in 1D upwind), A_fc: row j -> -1 at C index j */
  const int aff_i[4] = {0, 1, 2, 3}, aff_j[3] = {0, 1, 2};
  const double aff_a[3] = {1, 1, 1};
  const int afc_i[4] = {0, 1, 2, 3}, afc_j[3] = {0, 1, 2};
  const double afc_a[3] = {-1, -1, -1};
  const double minv[3] = {1, 1, 1};                      /* diagonal inverse of A_ff */
  /* R = [Z I] (3x6): row k = C point 2k: identity at column 2k, Z entry at F column 2k-1 (value 1: Z = -A_cf M_ff) */
  const int r_i[4] = {0, 1, 3, 5}, r_j[5] = {0, 1, 2, 3, 4};
  const double r_a[5] = {1, 1, 1, 1, 1};
  /* P = [W; I] (6x3): C rows identity, F row 2k+1 -> one-point at C index k */
  const int p_i[7] = {0, 1, 2, 3, 4, 5, 6}, p_j[6] = {0, 0, 1, 1, 2, 2};
  const double p_a[6] = {1, 1, 1, 1, 1, 1};
  /* level 2 (coarsest, 3 rows): coarse matrix + assembled inverse */
  const int c_i[4] = {0, 1, 3, 5}, c_j[5] = {0, 0, 1, 1, 2};
  const double c_a[5] = {1, -1, 1, -1, 1};
  const int ci_i[4] = {0, 1, 3, 6}, ci_j[6] = {0, 0, 1, 0, 1, 2};
  const double ci_a[6] = {1, 1, 1, 1, 1, 1};
  double stats[8], b[6] = {1, 2, 3, 4, 5, 6}, x[6];
  int back[3], l_agg = 0, ng = -1;
  int64_t rows[2];

  REQUIRE(pflare_b200_create(&h, 0, 1, NULL, -1, 2) == 0);        /* device -1: host-only planning context */
  REQUIRE(h != NULL);
  REQUIRE(pflare_b200_set_level(h, 3, 0, 6, 3, is_f, 3, is_c, smooth, 1) != 0);   /* level out of range -> error code ... */
  REQUIRE(strlen(pflare_b200_last_error()) > 0);                                  /* ... and a message */
  REQUIRE(pflare_b200_set_level(h, 1, 0, 6, 3, is_f, 3, is_c, smooth, 1) == 0);
  REQUIRE(pflare_b200_set_csr(h, 1, PFLARE_B200_AFF, 3, 3, 0, aff_i, aff_j, aff_a, 0, NULL, NULL, NULL, NULL) == 0);
  REQUIRE(pflare_b200_set_csr(h, 1, PFLARE_B200_AFC, 3, 3, 0, afc_i, afc_j, afc_a, 0, NULL, NULL, NULL, NULL) == 0);
  REQUIRE(pflare_b200_set_diag(h, 1, PFLARE_B200_INV_AFF, 3, minv) == 0);
  REQUIRE(pflare_b200_set_csr(h, 1, PFLARE_B200_R, 3, 6, 0, r_i, r_j, r_a, 0, NULL, NULL, NULL, NULL) == 0);
  REQUIRE(pflare_b200_set_csr(h, 1, PFLARE_B200_P, 6, 3, 0, p_i, p_j, p_a, 0, NULL, NULL, NULL, NULL) == 0);
  REQUIRE(pflare_b200_set_level(h, 2, 0, 3, 0, NULL, 0, NULL, smooth, 0) == 0);
  REQUIRE(pflare_b200_set_csr(h, 2, PFLARE_B200_COARSE, 3, 3, 0, c_i, c_j, c_a, 0, NULL, NULL, NULL, NULL) == 0);
  REQUIRE(pflare_b200_set_csr(h, 2, PFLARE_B200_INV_AFF, 3, 3, 0, ci_i, ci_j, ci_a, 0, NULL, NULL, NULL, NULL) == 0);
  REQUIRE(pflare_b200_set_poly(h, 1, PFLARE_B200_INV_AFF, PFLARE_B200_INV_SAI, 3, minv, NULL, 0) != 0);  /* SAI has no matrix-free form */
  REQUIRE(pflare_b200_set_diag(h, 1, PFLARE_B200_INV_AFF, 3, minv) == 0);
  REQUIRE(pflare_b200_apply(h, b, x, 0) != 0);                      /* before finalize_setup */
  REQUIRE(pflare_b200_finalize_setup(h) == 0);                      /* host part only: layout, program, counters */

  /* integer data round trip (bit exact) */
  REQUIRE(pflare_b200_get_is(h, 1, 0, back) == 0 && memcmp(back, is_f, sizeof is_f) == 0);
  REQUIRE(pflare_b200_get_is(h, 1, 1, back) == 0 && memcmp(back, is_c, sizeof is_c) == 0);
  REQUIRE(pflare_b200_get_garray(h, 1, PFLARE_B200_AFC, NULL, &ng) == 0 && ng == 0);
  REQUIRE(pflare_b200_get_layout(h, &l_agg, rows, 2) == 0 && l_agg == 3 && rows[0] == 6 && rows[1] == 3);

  /* work model: nnz per cycle = its*(nnz(M)+nnz(A_ff)) + nnz(A_fc) + nnz(Z) + nnz(W) + nnz(M_coarse)
   * (diagonal A_ff forces the fused local smooth; smooth_order says 2 iterations) */
  REQUIRE(pflare_b200_get_stats(h, stats, 8) == 0);
  REQUIRE(stats[2] == 2 * (3 + 3) + 3 + 2 + 3 + 6);
  REQUIRE(stats[1] > 12.0 * stats[2]);                              /* bytes = 12 B per nonzero + vectors */

  /* no device bound -> every compute entry point refuses (there is no CPU fallback) */
  REQUIRE(pflare_b200_apply(h, b, x, 0) == 10);
  REQUIRE(strstr(pflare_b200_last_error(), "no CPU fallback") != NULL);
  REQUIRE(pflare_b200_inv_apply(h, 1, PFLARE_B200_INV_AFF, b, x, 0) == 10);
  REQUIRE(pflare_b200_fc_smooth(h, 1, b, x, 0) == 10);
  REQUIRE(pflare_b200_set_option(h, "no_such_option", 1.0) != 0);
  REQUIRE(pflare_b200_destroy(&h) == 0 && h == NULL);
  REQUIRE(pflare_b200_destroy(&h) == 0);                            /* idempotent */
  printf("CAPI_HOST_OK\n");
  return 0;
}
