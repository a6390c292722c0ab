/* pflare_b200.h -- C-ABI of the B200-native AIRG V-cycle apply.
 *
 * Drop-in boundary for the apply path of PFLARE's PCAIR / PCPFLAREINV (reference paths are
 * relative to the PFLARE source tree):
 *
 *   pflare_b200_apply()       replaces PCApply(pcmg) inside PCApply_AIR_Shell
 *                             (src/PCAIR_Shell.F90:170-188, reached from PCApply_AIR_c,
 *                             src/PCAIR.c:150-166): one Kaskade V-cycle = restrict b down,
 *                             coarse solve (mg_coarse_shell_apply, src/FC_Smooth.F90:29-49),
 *                             interpolate + mg_FC_point_richardson (src/FC_Smooth.F90:421-640) up.
 *   pflare_b200_inv_apply()   replaces MatMult(mat_inverse) in PCApply_PFLAREINV_c
 *                             (src/PCPFLAREINV.c:618-626) and the polynomial MatShell mults
 *                             (src/Gmres_Poly.F90:1375-1518, src/Gmres_Poly_Newton.F90:716-912,
 *                             src/Neumann_Poly.F90:19-55).
 *   pflare_b200_set_*()       the upload hook at the end of setup_air_pcmg
 *                             (src/AIR_MG_Setup.F90:1178-1211): the reference's own setup
 *                             builds every operator; they are handed over once, as host CSR
 *                             in PETSc MPIAIJ layout (diag block + off-diag block + garray,
 *                             as MatMPIAIJGetSeqAIJ returns them, e.g.
 *                             src/Grid_Transferk.kokkos.cxx:30-42).
 *   pflare_b200_destroy()     release hook (reset_air_data, src/AIR_Data_Type_Routines.F90:105;
 *                             PCReset_AIR_c, src/PCAIR.c:134-148).
 *
 * Conventions (mirroring src/C_PETSc_Interfaces.F90:156-198): opaque handle passed by value,
 * `void **` for create/destroy; plain pointers and sizes only; every function returns 0 on
 * success and a non-zero error code otherwise (map to PetscErrorCode); no exceptions cross the
 * boundary; all calls are collective over the ranks given to pflare_b200_create().  Host arrays
 * are borrowed only for the duration of the call (copied).  PetscInt = 32-bit, PetscScalar =
 * real double (the only configuration the reference's load tests support, Makefile:84-86).
 * There is no CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef PFLARE_B200_H
#define PFLARE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Which operator of a level (names follow air_multigrid_data, src/AIR_Data_Type.F90:284-360). */
enum {
  PFLARE_B200_AFF = 0,      /* A_ff(our_level)                                     */
  PFLARE_B200_AFC = 1,      /* A_fc(our_level)                                     */
  PFLARE_B200_ACF = 2,      /* A_cf(our_level)            (C-point smoothing only) */
  PFLARE_B200_ACC = 3,      /* A_cc(our_level)            (C-point smoothing only) */
  PFLARE_B200_INV_AFF = 4,  /* inv_A_ff(our_level); on our_level == no_levels the coarse solver */
  PFLARE_B200_INV_ACC = 5,  /* inv_A_cc(our_level)                                 */
  PFLARE_B200_R = 6,        /* restrictors(our_level)  = [Z I], natural numbering  */
  PFLARE_B200_P = 7,        /* prolongators(our_level) = [W; I], natural numbering */
  PFLARE_B200_COARSE = 8    /* coarse_matrix(no_levels) (matrix-free coarse solver only) */
};

/* PCPFLAREINVType, same values as include/pflare.h:36-46 of the reference. */
enum {
  PFLARE_B200_INV_POWER = 0,
  PFLARE_B200_INV_ARNOLDI = 1,
  PFLARE_B200_INV_NEWTON = 2,
  PFLARE_B200_INV_NEWTON_NO_EXTRA = 3,
  PFLARE_B200_INV_NEUMANN = 4,
  PFLARE_B200_INV_SAI = 5,
  PFLARE_B200_INV_ISAI = 6,
  PFLARE_B200_INV_WJACOBI = 7,
  PFLARE_B200_INV_JACOBI = 8
};

/* Size of the opaque communicator id exchanged between ranks (== NCCL_UNIQUE_ID_BYTES). */
#define PFLARE_B200_UNIQUE_ID_BYTES 128

/* Fill `id` (PFLARE_B200_UNIQUE_ID_BYTES bytes) on rank 0; broadcast it to the other ranks
 * with the host communicator (MPI_Bcast in a PETSc build, torch.distributed in the tests). */
int pflare_b200_get_unique_id(void *id);

/* Create the per-PC device context (one per PCAIR/PCPFLAREINV instance -- the per-instance
 * handle pattern of air_data%kokkos_is_views_handle, src/AIR_Data_Type.F90:344-349).
 *   rank/nranks : position in the PC's communicator; nranks == 1 needs no unique id (NULL).
 *   device      : CUDA device ordinal to bind.
 *   no_levels   : air_data%no_levels (1 for PCPFLAREINV). */
int pflare_b200_create(void **handle, int rank, int nranks, const void *unique_id, int device, int no_levels);

/* Setup-time host communicator (multi-rank only).  By default the ranks exchange their setup
 * data (ownership ranges, ghost requests, agglomerated coarse levels) over NCCL; a host program that
 * already owns an MPI communicator (a PETSc build: the PC's comm) hands it in as two callbacks with
 * MPI_Alltoall / MPI_Alltoallv semantics on bytes, both returning 0 on success:
 *   int alltoall (void *ctx, const int64_t *send, int64_t *recv);                 one int64 per rank
 *   int alltoallv(void *ctx, const char *sbuf, const int64_t *scnt, const int64_t *sdsp,
 *                 char *rbuf, const int64_t *rcnt, const int64_t *rdsp);
 * With callbacks set, a context created with device == -1 can run finalize_setup WITHOUT a GPU: it
 * builds the ghost plans and the kernel program only ("planning context"; apply fails loudly). */
int pflare_b200_set_host_exchange(void *handle, void *alltoall_fn, void *alltoallv_fn, void *ctx);

/* Per-level metadata.  is_fine / is_coarse are the LOCAL (global - rstart) sorted index lists
 * of IS_fine_index / IS_coarse_index (src/VecISCopyLocalk.kokkos.cxx:73-132); smooth_order is
 * smooth_order_levels(our_level)%array (+k = k F smooths, -k = k C smooths, 0 terminates).
 * rstart = first global row owned by this rank on this level (MatGetOwnershipRange).
 * For our_level == no_levels pass n_fine = n_coarse = n_smooth = 0. */
int pflare_b200_set_level(void *handle, int our_level, int64_t rstart, int n_local, int n_fine, const int *is_fine,
                          int n_coarse, const int *is_coarse, const int *smooth_order, int n_smooth);

/* One assembled AIJ operator in MPIAIJ layout.  (di,dj,da) = diag block CSR with local column
 * indices; (oi,oj,oa) = off-diag block CSR over `n_ghost` compressed ghost columns whose global
 * column numbers are garray[0..n_ghost).  cstart = first global column owned by this rank
 * (MatGetOwnershipRangeColumn).  For nranks == 1 pass n_ghost = 0 and NULL off-diag arrays. */
int pflare_b200_set_csr(void *handle, int our_level, int which, int m, int n_local_cols, int64_t cstart, const int *di,
                        const int *dj, const double *da, int n_ghost, const int *oi, const int *oj, const double *oa,
                        const int64_t *garray);

/* PetscInt = 64-bit builds (--with-64-bit-indices; the reference's Makefile:75): the two upload calls above with 64-bit
 * index arrays.  A rank's LOCAL block is narrowed to 32 bits on entry (rows, local columns and nonzeros of one rank
 * must stay below 2^31 -- error 3 otherwise); global sizes are unrestricted (rstart, cstart and garray are 64-bit in
 * both variants). */
int pflare_b200_set_level_i64(void *handle, int our_level, int64_t rstart, int64_t n_local, int64_t n_fine, const int64_t *is_fine,
                              int64_t n_coarse, const int64_t *is_coarse, const int64_t *smooth_order, int64_t n_smooth);
int pflare_b200_set_csr_i64(void *handle, int our_level, int which, int64_t m, int64_t n_local_cols, int64_t cstart,
                            const int64_t *di, const int64_t *dj, const double *da, int64_t n_ghost, const int64_t *oi,
                            const int64_t *oj, const double *oa, const int64_t *garray);

/* MATDIAGONAL inverse (Jacobi / weighted Jacobi / 0th-order sparsity polynomial):
 * MatMult = pointwise multiply by d (src/Weighted_Jacobi.F90:76-85). which = INV_AFF | INV_ACC. */
int pflare_b200_set_diag(void *handle, int our_level, int which, int n, const double *d);

/* Matrix-free polynomial inverse (the MatShell of src/Gmres_Poly.F90:1568-1659 and
 * src/Gmres_Poly_Newton.F90:1931-2010).  coeffs_re[0..ncoef) are the power/Arnoldi/Neumann
 * coefficients, or the real parts of the Newton roots; coeffs_im the imaginary parts (NULL
 * unless Newton); conjugate pairs adjacent, positive imaginary part first, as
 * poly_data%coefficients(:,1:2).  diag_scale = -pc_air_diag_scale_polys.  The polynomial is
 * applied to A_ff (INV_AFF), A_cc (INV_ACC) or coarse_matrix (INV_AFF on the coarsest level),
 * which must be set with pflare_b200_set_csr. */
int pflare_b200_set_poly(void *handle, int our_level, int which, int inverse_type, int ncoef, const double *coeffs_re,
                         const double *coeffs_im, int diag_scale);

/* Build the device layout: nested CF ordering, fused-operator splitting (R -> Z, P -> W),
 * ghost exchange plans, the kernel program and its CUDA graph.  Collective. */
int pflare_b200_finalize_setup(void *handle);  /* collective over the ranks */

/* One AIRG V-cycle: x = PCApply(b).  b, x have n_local(level 1) entries in PETSc's natural
 * local ordering.  on_device != 0: b and x are device pointers and the call is asynchronous on
 * the handle's stream (pflare_b200_get_stream); on_device == 0: host pointers, the call copies
 * in, runs, copies out and returns after completion. */
int pflare_b200_apply(void *handle, const double *b, double *x, int on_device);

/* y = inverse * x for one approximate inverse of the hierarchy (PCPFLAREINV's PCApply with
 * no_levels == 1, our_level == 1, which == INV_AFF).  Same pointer convention as apply. */
int pflare_b200_inv_apply(void *handle, int our_level, int which, const double *x, double *y, int on_device);

/* One mg_FC_point_richardson sweep on a level (seam 2 of SURVEY.md section 8b): x is updated in place
 * from b; natural ordering of that level; same pointer convention. */
int pflare_b200_fc_smooth(void *handle, int our_level, const double *b, double *x, int on_device);

/* Outer Krylov method on the device: the KSPSolve the reference's drivers run around PCApply
 * (tests/Makefile:537-546, 1128-1134, 1322-1323; tests/adv_diff_fd.c: KSPGMRES, restart 30, b = 0, x0 = 1).
 * pflare_b200_ksp_set_operator hands over the system matrix (level-1 MPIAIJ blocks, natural numbering, same
 * conventions as pflare_b200_set_csr); it must be called before finalize_setup.  pflare_b200_ksp_solve runs
 *   ksp_type 0: KSPGMRES (restart <= 30, classical Gram-Schmidt), pc_side 0 = left, 1 = right
 *   ksp_type 1: KSPRICHARDSON (unpreconditioned norm)
 * with KSPConvergedDefault (||r|| <= max(rtol ||b||, atol); zero rhs and nonzero guess: rtol ||r0||), preconditioned
 * by the handle's V-cycle (no_levels >= 2) or its inverse (PCPFLAREINV handle, no_levels == 1).  x holds the
 * initial guess on entry and the solution on exit.  *its = iterations PETSc would report, *reason = 2 (rtol) /
 * 3 (atol) / -3 (max_it reached) as in KSPConvergedReason, *rnorm = last residual norm estimate.  Vectors, products,
 * V-cycles and all vector updates stay on the device; only the Gram-Schmidt scalars travel to the host. */
int pflare_b200_ksp_set_operator(void *handle, int m, int n_local_cols, int64_t cstart, const int *di, const int *dj,
                                 const double *da, int n_ghost, const int *oi, const int *oj, const double *oa,
                                 const int64_t *garray);
int pflare_b200_ksp_solve(void *handle, int ksp_type, int pc_side, double rtol, double atol, int max_it, int restart,
                          const double *b, double *x, int on_device, int *its, int *reason, double *rnorm);

/* The CUDA stream (cudaStream_t) all device work of this handle is ordered on. */
int pflare_b200_get_stream(void *handle, void **stream);
int pflare_b200_synchronize(void *handle);

/* Round trip of the integer data handed over at upload (bit-exact check). which_is: 0 = fine,
 * 1 = coarse; out must hold n_fine / n_coarse ints.  pflare_b200_get_garray returns the ghost
 * map of one uploaded operator. */
int pflare_b200_get_is(void *handle, int our_level, int which_is, int *out);
int pflare_b200_get_garray(void *handle, int our_level, int which, int64_t *out, int *n_ghost);

/* The ghost-exchange plan built for one uploaded operator (multi-rank; which = the selectors above,
 * R -> its Z block, P -> its W block): per peer rank how many entries this rank sends / receives,
 * where each peer's chunk starts in this rank's ghost buffer (ghost order == garray order), and the
 * positions (in the local vector segment the operator reads) this rank packs, peer after peer.
 * Any output pointer may be NULL.  The operator's column maps themselves round-trip bit-exactly
 * through pflare_b200_get_garray. */
int pflare_b200_get_ghost_plan(void *handle, int our_level, int which, int *send_count, int *recv_count, int *recv_off,
                               int *send_idx, int *n_send_idx);

/* Multi-rank layout chosen at finalize: *l_agg = first level that was agglomerated onto rank 0
 * (no_levels + 1 if none; option "agg_rows" = global-row threshold, the stand-in for the reference's
 * processor agglomeration, src/AIR_MG_Setup.F90:645-907), global_rows[l-1] = global rows of level l. */
int pflare_b200_get_layout(void *handle, int *l_agg, int64_t *global_rows, int n_levels);

/* Counters for measurement (per apply): stats[0] = kernel launches (graph nodes that are
 * kernels of this library), [1] = algorithmic HBM bytes of one V-cycle under the model of
 * SURVEY.md section 8d, [2] = nnz traversed per cycle (nnzs_air_v of src/AIR_MG_Stats.F90:79-252),
 * [3] = device bytes held, [4] = ghost bytes sent per cycle by this rank, [5] = algorithmic
 * bytes of the largest single kernel, [6] = number of levels run by the single-CTA tail kernel,
 * [7] = ghost / agglomeration exchange groups per cycle. */
int pflare_b200_get_stats(void *handle, double *stats, int nstats);

/* Per-op timing of one apply with CUDA events (disables the graph for that apply).
 * Fills up to max_ops entries: ms[i], bytes[i] (algorithmic), level[i], kind[i]; returns the
 * number of ops in *n_ops. */
int pflare_b200_profile_apply(void *handle, const double *b_dev, double *x_dev, int max_ops, float *ms, double *bytes,
                              int *level, int *kind, int *n_ops);

/* Switches.  Options of the reference that change the apply:
 *   "full_smoothing_up_and_down" (before finalize_setup)  -pc_air_full_smoothing_up_and_down: PCMG multiplicative V(1,1), one
 *                  Richardson sweep with inv_A_ff(level) on ALL unknowns down and up, residual restriction R (b - A x)
 *                  (src/AIR_MG_Setup.F90:978-1074); the hook hands over coarse_matrix(level) + inv_A_ff(level) instead of A_ff / A_fc.
 *   "mg_coarse_ksp_max_it" N (default 1)  -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it N: the coarse solve runs N Richardson
 *                  sweeps x_L += inv_A_ff(L) (b_L - A_L x_L) from a zero guess (KSP_NORM_NONE: exactly N, src/AIR_MG_Setup.F90:1094-1102;
 *                  tests/Makefile:132-136 uses 5); needs coarse_matrix(no_levels); may be changed on a finalized handle.
 * Execution switches (results stay within the 1e-12 parity bar):
 *   "graph" (0/1) CUDA graph of the cycle; "pdl" programmatic dependent launch; "fuse" (0: one kernel per PETSc call instead of
 *   the fused epilogues); "epi_classes" (0: run-time branched epilogue instead of the compiled classes);
 *   "kernel" (before finalize_setup) 0 = CSR stream kernel (also the fallback for rows longer than a warp tile), 1 = round-1 TMA
 *                  kernel with CTA tiles (A/B baseline), 2 = warp-tile storage (default);
 *   "wt_format" / "fmt_split" (before finalize_setup) warp-tile storage per operator: 0 = by mean row length (< fmt_split, default 8:
 *                  chunk format, else row-aligned lanes), 1 / 2 = force one of them;
 *   "engine" kernel of the row-aligned storage: 1 = direct (default), 0 = TMA ring (A/B);
 *   "ctas_per_sm" / "max_ctas" caps on the persistent grids (tests: many tiles per warp);
 *   "dense_rows" levels with <= this many rows are collapsed into one dense operator built from the same kernels at setup (0 = off);
 *   "fuse_perm" (before finalize_setup) 0 (default) / 1 / 2: entry / exit permutation fused into the level-1 ops (measured slower);
 *   "agg_rows" (multi-rank, before finalize_setup) levels with <= this many global rows are agglomerated onto rank 0 (0 = never);
 *   "p2p" (multi-rank, before finalize_setup) ghost exchange: 2 (default) = fused: the consuming SpMV kernel pushes this rank's
 *                  boundary entries straight into the peers' ghost buffers (CUDA IPC peer memory, one buffer per exchange of the
 *                  cycle, ready flags per exchange, one "entered the cycle" flag per cycle), multiplies its interior tiles and
 *                  waits for the peers' entries only at its first tile with ghost columns -- no launch for the exchange at all;
 *                  1 = push kernel + flags + acknowledgements; 0 = pack kernel + grouped ncclSend/ncclRecv ("overlap": on a side
 *                  stream under the interior tiles).  A wait for a peer that never arrives gives up after 120 s and the next
 *                  host-buffer apply returns error 27 (the GPU never hangs). */
int pflare_b200_set_option(void *handle, const char *key, double value);

const char *pflare_b200_last_error(void);

int pflare_b200_destroy(void **handle);

/* In-process rank group: `nranks` logical ranks of a partitioned hierarchy driven by ONE process on
 * one device and one stream, executed in lockstep; ghost exchanges are device-to-device copies.
 * (One process per GPU uses pflare_b200_create with a unique id and NCCL instead.)  Each rank's
 * operators are uploaded through its own handle (pflare_b200_cluster_rank + the set_* calls);
 * finalize / apply / destroy go through the group.  b[r] / x[r] / y[r] are rank r's local rows.
 * device == -1 gives a host-only planning group (no GPU needed; apply fails loudly). */
int pflare_b200_cluster_create(void **cluster, int nranks, int device, int no_levels);
int pflare_b200_cluster_rank(void *cluster, int rank, void **handle);
int pflare_b200_cluster_finalize(void *cluster);
int pflare_b200_cluster_apply(void *cluster, const double *const *b, double *const *x, int on_device);
int pflare_b200_cluster_inv_apply(void *cluster, int our_level, int which, const double *const *x, double *const *y,
                                  int on_device);
int pflare_b200_cluster_set_option(void *cluster, const char *key, double value);
int pflare_b200_cluster_get_stream(void *cluster, void **stream);
int pflare_b200_cluster_destroy(void **cluster);

#ifdef __cplusplus
}
#endif
#endif /* PFLARE_B200_H */
