/* air_oracle.c -- CPU restatement of PFLARE's AIRG V-cycle apply (PCApply of PCAIR) and of
 * the PCPFLAREINV matrix-free polynomial apply.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this library, and only as the checker or as
 * the reported CPU baseline; the product (pflare_b200/) never links or calls it.
 *
 * The reference itself (Fortran + PETSc >= 3.25) cannot be compiled here: no gfortran, MPI
 * or PETSc in the image, and PETSc -- which owns MatMult, the Vec
 * AXPY family and the PCMG Kaskade loop -- is not vendored under /root/reference.  This file
 * therefore restates, operation by operation and in the reference's order:
 *
 *   pcmg_kaskade()            PETSc PCMG, PC_MG_KASKADE, as wired by
 *                             src/AIR_MG_Setup.F90:967-1156 (restrict b down, x_L = 0,
 *                             coarse PREONLY solve, interpolate + 1 Richardson sweep up)
 *   pcmg_multiplicative()     -pc_air_full_smoothing_up_and_down: PETSc PCMG, PC_MG_MULTIPLICATIVE
 *                             V-cycle, one Richardson/PCMAT sweep with inv_A_ff(level) on ALL
 *                             unknowns down and up, true residual restriction R (b - A x)
 *                             (src/AIR_MG_Setup.F90:978-1074, src/AIR_Operators_Setup.F90:115-119)
 *   mg_coarse_shell_apply()   src/FC_Smooth.F90:29-49
 *   coarse_solve()            the coarse KSP around it (src/AIR_MG_Setup.F90:1094-1102): PREONLY, or N Richardson
 *                             sweeps with -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it N (tests/Makefile:132-136)
 *   mg_fc_point_richardson()  src/FC_Smooth.F90:421-495
 *   f_smooths()               src/FC_Smooth.F90:499-568
 *   c_smooths()               src/FC_Smooth.F90:572-640
 *   vec_is_copy()             VecISCopy / src/VecISCopyLocalk.kokkos.cxx:137-195
 *   petsc_horner()            src/Gmres_Poly.F90:1418-1484
 *   right-scaled Horner       src/Gmres_Poly.F90:1341-1414
 *   petsc_newton()            src/Gmres_Poly_Newton.F90:763-875 (+ right-scaled :716-759)
 *   Neumann I - D^-1 A        src/Neumann_Poly.F90:19-55
 *   MATDIAGONAL inverse       src/Weighted_Jacobi.F90:76-85, src/AIR_MG_Setup.F90:481-522
 *   PCApply_PFLAREINV         src/PCPFLAREINV.c:618-626
 *
 * PETSc semantics restated from PETSc's documented behaviour (source not in the container):
 *   MatMult_SeqAIJ: y_i = sum_j a_ij x_j in stored (sorted-column) order;
 *   VecAXPY(y,a,x): y += a x;  VecAYPX(y,a,x): y = x + a y;  VecAXPBY(y,a,b,x): y = a x + b y;
 *   VecPointwiseDivide(w,x,y): w = x / y.
 *
 * PARITY PINNING: the reference holds no vector-level golden outputs for PCApply
 * (SURVEY.md section 8c).  What it does hold are known-answer iteration bounds (tests/Makefile),
 * and those are what tests/test_oracle_pins.py checks this oracle against -- including the runs on the
 * reference's own data fixtures (mat_stream_2364 with its rhs, its ParILU factors, 1138_bus, e05r0100).  Vector-level
 * parity against a real PFLARE build is therefore UNPINNED ("parity unpinned").
 *
 * Rows are processed with OpenMP `parallel for`; per-row summation order is unchanged.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXLEV 320

/* `which` selectors shared with include/pflare_b200.h */
enum { W_AFF = 0, W_AFC = 1, W_ACF = 2, W_ACC = 3, W_INV_AFF = 4, W_INV_ACC = 5, W_R = 6, W_P = 7, W_COARSE = 8, W_COUNT = 9 };
/* PCPFLAREINVType (include/pflare.h:36-46) */
enum { T_POWER = 0, T_ARNOLDI, T_NEWTON, T_NEWTON_NO_EXTRA, T_NEUMANN, T_SAI, T_ISAI, T_WJACOBI, T_JACOBI };

/* PFLARE_TOL_ZERO: the single-precision literal 1e-12 widened to double (src/Pflare_Parameters.F90:206) */
static const double PFLARE_TOL_ZERO = (double)1e-12f;

typedef struct { int m, n; int64_t nnz; int *ia; int *ja; double *a; } csr_t;

typedef struct {
  int kind; /* 0 none, 1 assembled AIJ, 2 MATDIAGONAL, 3 matrix-free MatShell */
  csr_t mat;
  double *diag;
  int inverse_type, ncoef, diag_scale;
  double *re, *im;
  int which_A; /* which matrix the shell applies: W_AFF, W_ACC or W_COARSE */
  double *Adiag; /* MF_VEC_DIAG */
  double *t1, *t2, *t3, *rhs; /* mf_temp_vec(MF_VEC_TEMP..), MF_VEC_RHS */
} inv_t;

typedef struct {
  int n, nf, nc;
  int *is_f, *is_c;
  int nsmooth; int smooth[16];
  csr_t M[W_COUNT];
  inv_t inv_ff, inv_cc;
  double *tf[5], *tc[5]; /* temp_vecs_fine(1:4), temp_vecs_coarse(1:4) */
  double *b, *x;         /* PCMG level vectors */
  double *r, *z;         /* PCMG residual / Richardson work vectors (full smoothing only) */
} level_t;

typedef struct { int no_levels; int full_smoothing; int coarse_its; level_t L[MAXLEV]; } hier_t;

/* ---------------------------------------------------------------- PETSc primitives */
static void MatMult(const csr_t *A, const double *x, double *y) {
  const int m = A->m;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m; ++i) {
    double s = 0.0;
    for (int p = A->ia[i]; p < A->ia[i + 1]; ++p) s += A->a[p] * x[A->ja[p]];
    y[i] = s;
  }
}
static void VecAXPY(int n, double *y, double a, const double *x) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}
static void VecAYPX(int n, double *y, double a, const double *x) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] = x[i] + a * y[i];
}
static void VecAXPBY(int n, double *y, double a, double b, const double *x) {
  if (b == 0.0) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) y[i] = a * x[i];
  } else {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) y[i] = a * x[i] + b * y[i];
  }
}
static void VecPointwiseDivide(int n, double *w, const double *x, const double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) w[i] = x[i] / y[i];
}
static void VecPointwiseMult(int n, double *w, const double *x, const double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) w[i] = x[i] * y[i];
}
static void VecCopy(int n, const double *x, double *y) { memcpy(y, x, sizeof(double) * (size_t)n); }
static void VecSet(int n, double *y, double v) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] = v;
}
/* VecISCopy: SCATTER_REVERSE = gather reduced[i] = full[is[i]]; SCATTER_FORWARD = scatter */
static void vec_is_copy(int nis, const int *is, double *full, int reverse, double *reduced) {
  if (reverse) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nis; ++i) reduced[i] = full[is[i]];
  } else {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nis; ++i) full[is[i]] = reduced[i];
  }
}

/* ---------------------------------------------------------------- polynomial MatShells */
typedef struct { const csr_t *A; const double *diag; int mode; /* 0: A, 1: D^-1 A, 2: I - D^-1 A */ } shellA_t;

static void shell_mult(const shellA_t *S, const double *x, double *y) {
  MatMult(S->A, x, y);
  if (S->mode >= 1) VecPointwiseDivide(S->A->m, y, y, S->diag);          /* Gmres_Poly.F90:1366-1369 */
  if (S->mode == 2) VecAXPBY(S->A->m, y, 1.0, -1.0, x);                    /* Neumann_Poly.F90:45-52 */
}

/* src/Gmres_Poly.F90:1418-1484 */
static void petsc_horner(const shellA_t *S, int ncoef, const double *c, double *temp, const double *x, double *y) {
  const int n = S->A->m;
  VecAXPBY(n, y, c[ncoef - 1], 0.0, x);
  if (ncoef > 1) {
    for (int order = ncoef - 2; order >= 0; --order) {
      if (c[order] == 0.0) continue;
      VecCopy(n, y, temp);
      shell_mult(S, temp, y);
      VecAXPBY(n, y, c[order], 1.0, x);
    }
  }
}

/* src/Gmres_Poly_Newton.F90:763-875 */
static void petsc_newton(const shellA_t *S, int nr, const double *re, const double *im, double *t, double *t2, double *t3,
                         const double *x, double *y) {
  const int n = S->A->m;
  VecCopy(n, x, t);
  VecSet(n, y, 0.0);
  int i = 1; /* 1-based like the reference */
  while (i <= nr - 1) {
    if (im[i - 1] == 0.0) {
      if (fabs(re[i - 1]) < PFLARE_TOL_ZERO) { i += 1; continue; }
      VecAXPY(n, y, 1.0 / re[i - 1], t);
      shell_mult(S, t, t2);
      VecAXPY(n, t, -1.0 / re[i - 1], t2);
      i += 1;
    } else {
      const double sq = re[i - 1] * re[i - 1] + im[i - 1] * im[i - 1];
      if (sq < PFLARE_TOL_ZERO) { i += 2; continue; }
      shell_mult(S, t, t2);
      VecAXPBY(n, t2, 2.0 * re[i - 1], -1.0, t);
      VecAXPY(n, y, 1.0 / sq, t2);
      if (i <= nr - 2) {
        shell_mult(S, t2, t3);
        VecAXPY(n, t, -1.0 / sq, t3);
      }
      i += 2;
    }
  }
  if (im[nr - 1] == 0.0) {
    if (fabs(re[nr - 1]) > PFLARE_TOL_ZERO) VecAXPBY(n, y, 1.0 / re[nr - 1], 1.0, t);
  }
}

static void ensure(double **p, int n) { if (!*p) *p = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

/* MatMult on an approximate inverse (assembled AIJ | MATDIAGONAL | polynomial MatShell) */
static void inv_mult(inv_t *I, const csr_t *A, const double *x, double *y) {
  if (I->kind == 1) { MatMult(&I->mat, x, y); return; }
  if (I->kind == 2) { VecPointwiseMult(I->mat.m, y, x, I->diag); return; } /* MATDIAGONAL MatMult */
  const int n = A->m;
  ensure(&I->t1, n); ensure(&I->t2, n); ensure(&I->t3, n); ensure(&I->rhs, n);
  if (!I->Adiag) { /* MatGetDiagonal(matrix, MF_VEC_DIAG) */
    I->Adiag = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    for (int i = 0; i < n; ++i)
      for (int p = A->ia[i]; p < A->ia[i + 1]; ++p)
        if (A->ja[p] == i) I->Adiag[i] = A->a[p];
  }
  shellA_t S = { A, I->Adiag, 0 };
  const double *rhs = x;
  if (I->inverse_type == T_NEUMANN) {              /* q(I - D^-1 A) D^-1 x ; Neumann_Poly.F90:108-175 */
    S.mode = 2;
    VecPointwiseDivide(n, I->rhs, x, I->Adiag);     /* Gmres_Poly.F90:1406-1407 */
    rhs = I->rhs;
  } else if (I->diag_scale) {                       /* q(D^-1 A) D^-1 x */
    S.mode = 1;
    VecPointwiseDivide(n, I->rhs, x, I->Adiag);
    rhs = I->rhs;
  }
  if (I->inverse_type == T_NEWTON || I->inverse_type == T_NEWTON_NO_EXTRA)
    petsc_newton(&S, I->ncoef, I->re, I->im, I->t1, I->t2, I->t3, rhs, y);
  else
    petsc_horner(&S, I->ncoef, I->re, I->t1, rhs, y);
}

/* ---------------------------------------------------------------- F/C smoothing */
/* src/FC_Smooth.F90:499-568 */
static void f_smooths(level_t *L, double *b, double *x, int first_smooth, int its) {
  const int nf = L->nf, nc = L->nc;
  vec_is_copy(nf, L->is_f, b, 1, L->tf[4]);
  if (first_smooth) {
    vec_is_copy(nf, L->is_f, x, 1, L->tf[1]);
    vec_is_copy(nc, L->is_c, x, 1, L->tc[1]);
  }
  MatMult(&L->M[W_AFC], L->tc[1], L->tf[2]);
  VecAXPY(nf, L->tf[4], -1.0, L->tf[2]);
  for (int f = 1; f <= its; ++f) {
    MatMult(&L->M[W_AFF], L->tf[1], L->tf[3]);
    VecAYPX(nf, L->tf[3], -1.0, L->tf[4]);
    inv_mult(&L->inv_ff, &L->M[W_AFF], L->tf[3], L->tf[2]);
    VecAXPY(nf, L->tf[1], 1.0, L->tf[2]);
  }
  vec_is_copy(nf, L->is_f, x, 0, L->tf[1]);
}

/* src/FC_Smooth.F90:572-640 */
static void c_smooths(level_t *L, double *b, double *x, int first_smooth, int its) {
  const int nf = L->nf, nc = L->nc;
  vec_is_copy(nc, L->is_c, b, 1, L->tc[4]);
  if (first_smooth) {
    vec_is_copy(nf, L->is_f, x, 1, L->tf[1]);
    vec_is_copy(nc, L->is_c, x, 1, L->tc[1]);
  }
  MatMult(&L->M[W_ACF], L->tf[1], L->tc[2]);
  VecAXPY(nc, L->tc[4], -1.0, L->tc[2]);
  for (int c = 1; c <= its; ++c) {
    MatMult(&L->M[W_ACC], L->tc[1], L->tc[3]);
    VecAYPX(nc, L->tc[3], -1.0, L->tc[4]);
    inv_mult(&L->inv_cc, &L->M[W_ACC], L->tc[3], L->tc[2]);
    VecAXPY(nc, L->tc[1], 1.0, L->tc[2]);
  }
  vec_is_copy(nc, L->is_c, x, 0, L->tc[1]);
}

/* src/FC_Smooth.F90:421-495 (maxits == 1; guess_zero ignored; r never updated) */
static void mg_fc_point_richardson(level_t *L, double *b, double *x) {
  int first_smooth = 1;
  for (int i = 0; i < L->nsmooth; ++i) {
    const int s = L->smooth[i];
    if (s == 0) break;
    if (s > 0) f_smooths(L, b, x, first_smooth, s);
    else c_smooths(L, b, x, first_smooth, -s);
    first_smooth = 0;
  }
}

/* ---------------------------------------------------------------- the cycle */
/* Coarse solve of both cycles.  Default: KSPPREONLY around the PCSHELL mg_coarse_shell_apply (src/FC_Smooth.F90:29-49,
 * src/AIR_MG_Setup.F90:1094-1102), i.e. x_L = inv_A_ff(L) b_L.  With -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it N the
 * same KSP (KSP_NORM_NONE, zero initial guess, scale 1) runs exactly N Richardson sweeps:
 *   x = 0 ; x += M b ; then N - 1 times  r = b - A_L x ; x += M r            (tests/Makefile:132-136 uses N = 5) */
static void coarse_solve(hier_t *H) {
  level_t *Lc = &H->L[H->no_levels];
  const int n = Lc->n;
  VecSet(n, Lc->x, 0.0);
  inv_mult(&Lc->inv_ff, &Lc->M[W_COARSE], Lc->b, Lc->x);                                /* FC_Smooth.F90:47 */
  for (int it = 1; it < H->coarse_its; ++it) {
    ensure(&Lc->r, n); ensure(&Lc->z, n);
    MatMult(&Lc->M[W_COARSE], Lc->x, Lc->r);
    VecAYPX(n, Lc->r, -1.0, Lc->b);
    inv_mult(&Lc->inv_ff, &Lc->M[W_COARSE], Lc->r, Lc->z);
    VecAXPY(n, Lc->x, 1.0, Lc->z);
  }
}

/* PETSc PCMG in PC_MG_KASKADE mode as configured by src/AIR_MG_Setup.F90:967-1156 */
static void pcmg_kaskade(hier_t *H, const double *b_in, double *x_out) {
  const int NL = H->no_levels;
  level_t *L = H->L;
  VecCopy(L[1].n, b_in, L[1].b);
  for (int l = 1; l <= NL - 1; ++l) MatMult(&L[l].M[W_R], L[l].b, L[l + 1].b);          /* MatRestrict */
  coarse_solve(H);
  for (int l = NL - 1; l >= 1; --l) {
    MatMult(&L[l].M[W_P], L[l + 1].x, L[l].x);                                          /* MatInterpolate */
    mg_fc_point_richardson(&L[l], L[l].b, L[l].x);
  }
  VecCopy(L[1].n, L[1].x, x_out);
}

/* MatMultAdd_SeqAIJ(A, x, y, y): the row sum starts from y_i (PETSc: sum = y[i]; sum += a_ij x_j ...) */
static void MatMultAddInPlace(const csr_t *A, const double *x, double *y) {
  const int m = A->m;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m; ++i) {
    double s = y[i];
    for (int p = A->ia[i]; p < A->ia[i + 1]; ++p) s += A->a[p] * x[A->ja[p]];
    y[i] = s;
  }
}

/* -pc_air_full_smoothing_up_and_down (src/AIR_MG_Setup.F90:978-1074): PCMG keeps its default
 * PC_MG_MULTIPLICATIVE V-cycle (PCMGMCycle_Private of PETSc's mg.c), the level smoothers are
 * KSPRICHARDSON (1 iteration, KSP_NORM_NONE, scale 1) preconditioned by PCMAT with
 * inv_A_ff(our_level) -- here an approximate inverse of the WHOLE level matrix
 * coarse_matrix(our_level) (src/AIR_Operators_Setup.F90:115-119, 373-377):
 *   down:  x_l = 0 ; x_l += M_l b_l            (pre-smooth; zero guess: r = b)
 *          r_l = b_l - A_l x_l                  (PCMGResidualDefault)
 *          b_{l+1} = R_l r_l ; x_{l+1} = 0      (MatRestrict)
 *   coarsest: x_L = inv_A_ff(L) b_L             (mg_coarse_shell_apply)
 *   up:    x_l = x_l + P_l x_{l+1}              (MatInterpolateAdd)
 *          r = b_l - A_l x_l ; x_l += M_l r     (post-smooth; nonzero guess) */
static void pcmg_multiplicative(hier_t *H, const double *b_in, double *x_out) {
  const int NL = H->no_levels;
  level_t *L = H->L;
  VecCopy(L[1].n, b_in, L[1].b);
  for (int l = 1; l <= NL - 1; ++l) {
    const int n = L[l].n;
    ensure(&L[l].r, n); ensure(&L[l].z, n);
    VecSet(n, L[l].x, 0.0);
    inv_mult(&L[l].inv_ff, &L[l].M[W_COARSE], L[l].b, L[l].z);
    VecAXPY(n, L[l].x, 1.0, L[l].z);
    MatMult(&L[l].M[W_COARSE], L[l].x, L[l].r);
    VecAYPX(n, L[l].r, -1.0, L[l].b);
    MatMult(&L[l].M[W_R], L[l].r, L[l + 1].b);
  }
  coarse_solve(H);
  for (int l = NL - 1; l >= 1; --l) {
    const int n = L[l].n;
    MatMultAddInPlace(&L[l].M[W_P], L[l + 1].x, L[l].x);
    MatMult(&L[l].M[W_COARSE], L[l].x, L[l].r);
    VecAYPX(n, L[l].r, -1.0, L[l].b);
    inv_mult(&L[l].inv_ff, &L[l].M[W_COARSE], L[l].r, L[l].z);
    VecAXPY(n, L[l].x, 1.0, L[l].z);
  }
  VecCopy(L[1].n, L[1].x, x_out);
}

/* ---------------------------------------------------------------- construction API */
static void csr_copy(csr_t *d, int m, int n, const int *ia, const int *ja, const double *a) {
  free(d->ia); free(d->ja); free(d->a);
  d->m = m; d->n = n; d->nnz = ia[m];
  d->ia = (int *)malloc(sizeof(int) * (size_t)(m + 1));
  d->ja = (int *)malloc(sizeof(int) * (size_t)(d->nnz > 0 ? d->nnz : 1));
  d->a = (double *)malloc(sizeof(double) * (size_t)(d->nnz > 0 ? d->nnz : 1));
  memcpy(d->ia, ia, sizeof(int) * (size_t)(m + 1));
  memcpy(d->ja, ja, sizeof(int) * (size_t)d->nnz);
  memcpy(d->a, a, sizeof(double) * (size_t)d->nnz);
}

void *oracle_create(int no_levels) {
  if (no_levels < 1 || no_levels >= MAXLEV) return NULL;
  hier_t *H = (hier_t *)calloc(1, sizeof(hier_t));
  H->no_levels = no_levels;
  return H;
}

int oracle_set_level(void *h, int our_level, int n, int nf, const int *is_f, int nc, const int *is_c, const int *smooth,
                     int nsmooth) {
  hier_t *H = (hier_t *)h;
  level_t *L = &H->L[our_level];
  L->n = n; L->nf = nf; L->nc = nc;
  L->is_f = (int *)malloc(sizeof(int) * (size_t)(nf > 0 ? nf : 1));
  L->is_c = (int *)malloc(sizeof(int) * (size_t)(nc > 0 ? nc : 1));
  if (nf) memcpy(L->is_f, is_f, sizeof(int) * (size_t)nf);
  if (nc) memcpy(L->is_c, is_c, sizeof(int) * (size_t)nc);
  L->nsmooth = nsmooth > 16 ? 16 : nsmooth;
  for (int i = 0; i < L->nsmooth; ++i) L->smooth[i] = smooth[i];
  for (int k = 1; k <= 4; ++k) { L->tf[k] = (double *)calloc((size_t)(nf > 0 ? nf : 1), 8); L->tc[k] = (double *)calloc((size_t)(nc > 0 ? nc : 1), 8); }
  L->b = (double *)calloc((size_t)(n > 0 ? n : 1), 8);
  L->x = (double *)calloc((size_t)(n > 0 ? n : 1), 8);
  return 0;
}

static inv_t *pick_inv(level_t *L, int which) { return which == W_INV_ACC ? &L->inv_cc : &L->inv_ff; }

int oracle_set_csr(void *h, int our_level, int which, int m, int n, const int *ia, const int *ja, const double *a) {
  hier_t *H = (hier_t *)h;
  level_t *L = &H->L[our_level];
  if (which == W_INV_AFF || which == W_INV_ACC) {
    inv_t *I = pick_inv(L, which);
    I->kind = 1;
    csr_copy(&I->mat, m, n, ia, ja, a);
  } else {
    csr_copy(&L->M[which], m, n, ia, ja, a);
  }
  return 0;
}

int oracle_set_diag(void *h, int our_level, int which, int n, const double *d) {
  hier_t *H = (hier_t *)h;
  inv_t *I = pick_inv(&H->L[our_level], which);
  I->kind = 2; I->mat.m = n; I->mat.n = n;
  I->diag = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  memcpy(I->diag, d, sizeof(double) * (size_t)n);
  return 0;
}

int oracle_set_poly(void *h, int our_level, int which, int inverse_type, int ncoef, const double *re, const double *im,
                    int diag_scale) {
  hier_t *H = (hier_t *)h;
  inv_t *I = pick_inv(&H->L[our_level], which);
  I->kind = 3; I->inverse_type = inverse_type; I->ncoef = ncoef; I->diag_scale = diag_scale;
  I->re = (double *)calloc((size_t)ncoef, 8); I->im = (double *)calloc((size_t)ncoef, 8);
  memcpy(I->re, re, 8 * (size_t)ncoef);
  if (im) memcpy(I->im, im, 8 * (size_t)ncoef);
  return 0;
}

/* PCApply_AIR_c -> PCApply_AIR_Shell -> PCApply(pcmg): src/PCAIR.c:150-166, src/PCAIR_Shell.F90:170-188 */
int oracle_pcapply(void *h, const double *b, double *x) {
  hier_t *H = (hier_t *)h;
  if (H->no_levels == 1) return 1; /* the reference falls back to PCJACOBI here (AIR_MG_Setup.F90:1167-1174) */
  if (!H->L[H->no_levels].b) return 2; /* oracle_set_level must be called for every level incl. the coarsest */
  if (H->full_smoothing) pcmg_multiplicative(H, b, x);
  else pcmg_kaskade(H, b, x);
  return 0;
}

/* -pc_air_full_smoothing_up_and_down: inv_A_ff(l) then inverts coarse_matrix(l) (set as W_COARSE on every level) */
void oracle_set_full_smoothing(void *h, int flag) { ((hier_t *)h)->full_smoothing = flag; }
/* -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it n */
void oracle_set_coarse_its(void *h, int n) { ((hier_t *)h)->coarse_its = n < 1 ? 1 : n; }

/* PCApply_PFLAREINV_c: y = mat_inverse * x (src/PCPFLAREINV.c:618-626); also used by tests to
 * exercise any level's inverse on its own.  which = W_INV_AFF (matrix W_AFF, or W_COARSE on
 * the coarsest level) or W_INV_ACC (matrix W_ACC). */
int oracle_inv_apply(void *h, int our_level, int which, const double *x, double *y) {
  hier_t *H = (hier_t *)h;
  level_t *L = &H->L[our_level];
  inv_t *I = pick_inv(L, which);
  const csr_t *A = which == W_INV_ACC ? &L->M[W_ACC] : ((our_level == H->no_levels || H->full_smoothing) ? &L->M[W_COARSE] : &L->M[W_AFF]);
  if (I->kind == 0) return 1;
  inv_mult(I, A, x, y);
  return 0;
}

/* y = A x for an uploaded matrix (test helper) */
int oracle_matmult(void *h, int our_level, int which, const double *x, double *y) {
  hier_t *H = (hier_t *)h;
  MatMult(&H->L[our_level].M[which], x, y);
  return 0;
}

/* one mg_FC_point_richardson on a level (test helper for the minimum slice) */
int oracle_fc_smooth(void *h, int our_level, const double *b, double *x) {
  hier_t *H = (hier_t *)h;
  level_t *L = &H->L[our_level];
  VecCopy(L->n, b, L->b);
  mg_fc_point_richardson(L, L->b, x);
  return 0;
}

static void inv_free(inv_t *I) {
  free(I->mat.ia); free(I->mat.ja); free(I->mat.a); free(I->diag); free(I->re); free(I->im);
  free(I->Adiag); free(I->t1); free(I->t2); free(I->t3); free(I->rhs);
}

void oracle_destroy(void *h) {
  hier_t *H = (hier_t *)h;
  if (!H) return;
  for (int l = 0; l < MAXLEV; ++l) {
    level_t *L = &H->L[l];
    free(L->is_f); free(L->is_c); free(L->b); free(L->x); free(L->r); free(L->z);
    for (int k = 0; k < 5; ++k) { free(L->tf[k]); free(L->tc[k]); }
    for (int w = 0; w < W_COUNT; ++w) { free(L->M[w].ia); free(L->M[w].ja); free(L->M[w].a); }
    inv_free(&L->inv_ff); inv_free(&L->inv_cc);
  }
  free(H);
}

/* bench.py under torchrun inherits OMP_NUM_THREADS=1: the CPU baseline sets its thread count explicitly */
void oracle_set_omp_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_omp_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
