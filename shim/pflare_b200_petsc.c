/* pflare_b200_petsc.c -- PETSc-side uploader for libpflare_b200 (add to PFLARE's src/, link -lpflare_b200).
 *
 * NOT COMPILED IN THIS REPOSITORY (no PETSc here).  Mirrors the access patterns of the reference:
 *   MatMPIAIJGetSeqAIJ + CSR access      src/Grid_Transferk.kokkos.cxx:30-66, src/MatDiagDom.F90:167-187
 *   local IS lists (global - rstart)     src/VecISCopyLocalk.kokkos.cxx:73-132
 *   C functions callable from Fortran    src/C_PETSc_Routines.c (PETSC_INTERN void f(...), PetscCallVoid)
 */
#include <petscmat.h>
#include <petscis.h>
#include <stdint.h>
#include "pflare_b200.h"

#define B200_CHECK(call)                                                                      \
  do {                                                                                        \
    if ((call) != 0) SETERRABORT(PETSC_COMM_WORLD, PETSC_ERR_LIB, pflare_b200_last_error()); \
  } while (0)

/* ---- setup-time communicator = the PC's MPI communicator ------------------------------------------------- */
static int b200_alltoall(void *ctx, const int64_t *send, int64_t *recv)
{
  return MPI_Alltoall((void *)send, 1, MPI_INT64_T, recv, 1, MPI_INT64_T, *(MPI_Comm *)ctx) != MPI_SUCCESS;
}

static int b200_alltoallv(void *ctx, const char *sbuf, const int64_t *scnt, const int64_t *sdsp, char *rbuf,
                          const int64_t *rcnt, const int64_t *rdsp)
{
  MPI_Comm comm = *(MPI_Comm *)ctx;
  int      P, rc;
  MPI_Comm_size(comm, &P);
  int *sc = (int *)malloc(4 * sizeof(int) * (size_t)P), *sd = sc + P, *rc_ = sd + P, *rd = rc_ + P;
  for (int p = 0; p < P; ++p) { /* setup messages are far below 2 GiB per peer */
    sc[p] = (int)scnt[p]; sd[p] = (int)sdsp[p]; rc_[p] = (int)rcnt[p]; rd[p] = (int)rdsp[p];
  }
  rc = MPI_Alltoallv((void *)sbuf, sc, sd, MPI_BYTE, rbuf, rc_, rd, MPI_BYTE, comm);
  free(sc);
  return rc != MPI_SUCCESS;
}

/* One static communicator slot per handle would live in air_data; shown here for a single PC. */
static MPI_Comm b200_comm;

PETSC_INTERN void pflare_b200_create_c(void **handle, Mat *A_top, PetscInt no_levels)
{
  MPI_Comm    comm;
  PetscMPIInt rank, size;
  int         device = 0, ndev = 1;
  char        uid[PFLARE_B200_UNIQUE_ID_BYTES];

  PetscCallVoid(PetscObjectGetComm((PetscObject)*A_top, &comm));
  PetscCallMPIAbort(comm, MPI_Comm_rank(comm, &rank));
  PetscCallMPIAbort(comm, MPI_Comm_size(comm, &size));
  /* one rank per GPU: PETSc has already bound this rank to its device (-device_select / PetscDeviceContext) */
  cudaGetDeviceCount(&ndev);
  cudaGetDevice(&device);
  if (size > 1) {
    if (rank == 0) B200_CHECK(pflare_b200_get_unique_id(uid));
    PetscCallMPIAbort(comm, MPI_Bcast(uid, PFLARE_B200_UNIQUE_ID_BYTES, MPI_BYTE, 0, comm));
  }
  B200_CHECK(pflare_b200_create(handle, rank, size, size > 1 ? uid : NULL, device, (int)no_levels));
  if (size > 1) {
    b200_comm = comm;
    B200_CHECK(pflare_b200_set_host_exchange(*handle, (void *)b200_alltoall, (void *)b200_alltoallv, &b200_comm));
  }
}

/* IS_fine_index / IS_coarse_index (global, sorted) -> local lists; smooth_order = smooth_order_levels(l)%array */
PETSC_INTERN void pflare_b200_set_level_c(void **handle, PetscInt our_level, Mat *A_level, IS *is_fine, IS *is_coarse,
                                          const int *smooth_order, int n_smooth)
{
  PetscInt        rstart, rend, nf = 0, nc = 0;
  const PetscInt *f = NULL, *c = NULL;
  int            *lf, *lc;

  PetscCallVoid(MatGetOwnershipRange(*A_level, &rstart, &rend));
  if (is_fine) { PetscCallVoid(ISGetLocalSize(*is_fine, &nf)); PetscCallVoid(ISGetIndices(*is_fine, &f)); }
  if (is_coarse) { PetscCallVoid(ISGetLocalSize(*is_coarse, &nc)); PetscCallVoid(ISGetIndices(*is_coarse, &c)); }
  PetscCallVoid(PetscMalloc2(nf, &lf, nc, &lc));
  for (PetscInt i = 0; i < nf; ++i) lf[i] = (int)(f[i] - rstart);
  for (PetscInt i = 0; i < nc; ++i) lc[i] = (int)(c[i] - rstart);
  B200_CHECK(pflare_b200_set_level(*handle, (int)our_level, (int64_t)rstart, (int)(rend - rstart), (int)nf, lf, (int)nc, lc,
                                   smooth_order, n_smooth));
  if (is_fine) PetscCallVoid(ISRestoreIndices(*is_fine, &f));
  if (is_coarse) PetscCallVoid(ISRestoreIndices(*is_coarse, &c));
  PetscCallVoid(PetscFree2(lf, lc));
}

/* One assembled AIJ operator (SEQAIJ or MPIAIJ, host CSR) -> pflare_b200_set_csr. */
PETSC_INTERN void pflare_b200_upload_mat_c(void **handle, PetscInt our_level, int which, Mat *A)
{
  Mat                Ad = *A, Ao = NULL;
  const PetscInt    *garray = NULL, *di, *dj, *oi = NULL, *oj = NULL;
  const PetscScalar *da, *oa = NULL;
  PetscInt           m, n, cstart, nd, no = 0, ng = 0;
  PetscBool          mpi, done;
  int64_t           *g64 = NULL;

  PetscCallVoid(PetscObjectBaseTypeCompare((PetscObject)*A, MATMPIAIJ, &mpi));
  if (mpi) PetscCallVoid(MatMPIAIJGetSeqAIJ(*A, &Ad, &Ao, &garray));
  PetscCallVoid(MatGetLocalSize(*A, &m, &n));
  PetscCallVoid(MatGetOwnershipRangeColumn(*A, &cstart, NULL));
  PetscCallVoid(MatGetRowIJ(Ad, 0, PETSC_FALSE, PETSC_FALSE, &nd, &di, &dj, &done));
  PetscCallVoid(MatSeqAIJGetArrayRead(Ad, &da));
  if (Ao) {
    PetscCallVoid(MatGetSize(Ao, NULL, &ng)); /* compressed ghost columns */
    PetscCallVoid(MatGetRowIJ(Ao, 0, PETSC_FALSE, PETSC_FALSE, &no, &oi, &oj, &done));
    PetscCallVoid(MatSeqAIJGetArrayRead(Ao, &oa));
    PetscCallVoid(PetscMalloc1(ng, &g64));
    for (PetscInt k = 0; k < ng; ++k) g64[k] = (int64_t)garray[k];
  }
  /* PetscInt must be 32-bit (the only configuration the reference's load tests cover, Makefile:84-86) */
  B200_CHECK(pflare_b200_set_csr(*handle, (int)our_level, which, (int)m, (int)n, (int64_t)cstart, (const int *)di, (const int *)dj,
                                 da, (int)ng, (const int *)oi, (const int *)oj, oa, g64));
  PetscCallVoid(MatSeqAIJRestoreArrayRead(Ad, &da));
  PetscCallVoid(MatRestoreRowIJ(Ad, 0, PETSC_FALSE, PETSC_FALSE, &nd, &di, &dj, &done));
  if (Ao) {
    PetscCallVoid(MatSeqAIJRestoreArrayRead(Ao, &oa));
    PetscCallVoid(MatRestoreRowIJ(Ao, 0, PETSC_FALSE, PETSC_FALSE, &no, &oi, &oj, &done));
    PetscCallVoid(PetscFree(g64));
  }
}

/* MATDIAGONAL inverse (src/Weighted_Jacobi.F90:76-85, src/AIR_MG_Setup.F90:481-522) */
PETSC_INTERN void pflare_b200_upload_diag_c(void **handle, PetscInt our_level, int which, Mat *D)
{
  Vec                d;
  const PetscScalar *a;
  PetscInt           n;
  PetscCallVoid(MatCreateVecs(*D, &d, NULL));
  PetscCallVoid(MatGetDiagonal(*D, d));
  PetscCallVoid(VecGetLocalSize(d, &n));
  PetscCallVoid(VecGetArrayRead(d, &a));
  B200_CHECK(pflare_b200_set_diag(*handle, (int)our_level, which, (int)n, a));
  PetscCallVoid(VecRestoreArrayRead(d, &a));
  PetscCallVoid(VecDestroy(&d));
}

/* PCApply: x, y are the Vecs of PCApply_AIR_Shell (src/PCAIR_Shell.F90:170-188) */
PETSC_INTERN void pflare_b200_apply_c(void **handle, Vec *x, Vec *y)
{
  const PetscScalar *xa;
  PetscScalar       *ya;
  PetscMemType       mx, my;
  PetscCallVoid(VecGetArrayReadAndMemType(*x, &xa, &mx));
  PetscCallVoid(VecGetArrayWriteAndMemType(*y, &ya, &my));
  const int on_device = PetscMemTypeDevice(mx) && PetscMemTypeDevice(my);
  B200_CHECK(pflare_b200_apply(*handle, xa, ya, on_device));
  if (on_device) B200_CHECK(pflare_b200_synchronize(*handle)); /* or order the library stream with PETSc's device context */
  PetscCallVoid(VecRestoreArrayReadAndMemType(*x, &xa));
  PetscCallVoid(VecRestoreArrayWriteAndMemType(*y, &ya));
}

/* PCPFLAREINV: y = mat_inverse * x (src/PCPFLAREINV.c:618-626) */
PETSC_INTERN void pflare_b200_inv_apply_c(void **handle, Vec *x, Vec *y)
{
  const PetscScalar *xa;
  PetscScalar       *ya;
  PetscMemType       mx, my;
  PetscCallVoid(VecGetArrayReadAndMemType(*x, &xa, &mx));
  PetscCallVoid(VecGetArrayWriteAndMemType(*y, &ya, &my));
  const int on_device = PetscMemTypeDevice(mx) && PetscMemTypeDevice(my);
  B200_CHECK(pflare_b200_inv_apply(*handle, 1, PFLARE_B200_INV_AFF, xa, ya, on_device));
  if (on_device) B200_CHECK(pflare_b200_synchronize(*handle));
  PetscCallVoid(VecRestoreArrayReadAndMemType(*x, &xa));
  PetscCallVoid(VecRestoreArrayWriteAndMemType(*y, &ya));
}

PETSC_INTERN void pflare_b200_finalize_c(void **handle) { B200_CHECK(pflare_b200_finalize_setup(*handle)); }
PETSC_INTERN void pflare_b200_destroy_c(void **handle) { if (*handle) B200_CHECK(pflare_b200_destroy(handle)); }
