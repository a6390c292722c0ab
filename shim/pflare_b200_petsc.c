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

/* Per-PC state kept next to the library handle in air_data (two c_ptr members: b200_handle, b200_aux): every PCAIR /
 * PCPFLAREINV instance owns its own communicator slot and ordering events (per-instance handle rule,
 * src/AIR_Data_Type.F90:344-349; regression tests/ex6_two_airg.c). */
typedef struct {
  MPI_Comm    comm;          /* the PC's communicator: setup-time exchanges of the library */
  cudaEvent_t ev_in, ev_out; /* ordering between PETSc's stream and the library's stream */
} B200Aux;

PETSC_INTERN void pflare_b200_create_c(void **handle, void **aux_out, Mat *A_top, PetscInt no_levels)
{
  MPI_Comm    comm;
  B200Aux    *aux;
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
  PetscCallVoid(PetscNew(&aux));
  aux->comm = comm;
  cudaEventCreateWithFlags(&aux->ev_in, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&aux->ev_out, cudaEventDisableTiming);
  *aux_out = aux;
  if (size > 1) B200_CHECK(pflare_b200_set_host_exchange(*handle, (void *)b200_alltoall, (void *)b200_alltoallv, &aux->comm));
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
#if defined(PETSC_USE_64BIT_INDICES)
  /* --with-64-bit-indices (Makefile:75): the _i64 entry point narrows this rank's local block (must stay below 2^31) */
  B200_CHECK(pflare_b200_set_csr_i64(*handle, (int)our_level, which, m, n, (int64_t)cstart, (const int64_t *)di, (const int64_t *)dj, da, ng,
                                     (const int64_t *)oi, (const int64_t *)oj, oa, g64));
#else
  B200_CHECK(pflare_b200_set_csr(*handle, (int)our_level, which, (int)m, (int)n, (int64_t)cstart, (const int *)di, (const int *)dj,
                                 da, (int)ng, (const int *)oi, (const int *)oj, oa, g64));
#endif
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

/* Richardson sweeps of the PCMG coarse solve (option "mg_coarse_ksp_max_it"): KSPPREONLY (the default of
 * src/AIR_MG_Setup.F90:1094-1102) = 1, -mg_coarse_ksp_type richardson -mg_coarse_ksp_max_it N = N. */
PETSC_INTERN void pflare_b200_coarse_its_c(PC *pcmg, int *its)
{
  KSP       coarse;
  PetscBool is_rich;
  PetscInt  max_it;
  *its = 1;
  PetscCallVoid(PCMGGetCoarseSolve(*pcmg, &coarse));
  PetscCallVoid(PetscObjectTypeCompare((PetscObject)coarse, KSPRICHARDSON, &is_rich));
  if (is_rich) {
    PetscCallVoid(KSPGetTolerances(coarse, NULL, NULL, NULL, &max_it));
    *its = (int)max_it;
  }
}

/* Stream ordering around a device-pointer call.  The library runs on its own non-blocking stream
 * (pflare_b200_get_stream); PETSc's kernels that produced x run on the PetscDeviceContext's stream.  BEFORE the call
 * the library stream waits for PETSc's stream, AFTER it PETSc's stream waits for the library's: no host
 * synchronisation, and a device KSP never reads a half-written vector. */
static void b200_order_in(void *handle, B200Aux *aux, cudaStream_t *petsc_stream)
{
  PetscDeviceContext dctx;
  void              *hdl = NULL, *lib = NULL;
  PetscCallVoid(PetscDeviceContextGetCurrentContext(&dctx));
  PetscCallVoid(PetscDeviceContextGetStreamHandle(dctx, &hdl));
  *petsc_stream = *(cudaStream_t *)hdl;
  B200_CHECK(pflare_b200_get_stream(handle, &lib));
  cudaEventRecord(aux->ev_in, *petsc_stream);
  cudaStreamWaitEvent((cudaStream_t)lib, aux->ev_in, 0);
}
static void b200_order_out(void *handle, B200Aux *aux, cudaStream_t petsc_stream)
{
  void *lib = NULL;
  B200_CHECK(pflare_b200_get_stream(handle, &lib));
  cudaEventRecord(aux->ev_out, (cudaStream_t)lib);
  cudaStreamWaitEvent(petsc_stream, aux->ev_out, 0);
}

/* x (read) and y (write) either both on the device or both on the host: a mixed pair is brought to the host (PETSc
 * synchronises and copies inside VecGetArray*), never passed with the wrong pointer kind */
static int b200_get_arrays(Vec x, Vec y, const PetscScalar **xa, PetscScalar **ya)
{
  PetscMemType mx, my;
  PetscCallAbort(PETSC_COMM_SELF, VecGetArrayReadAndMemType(x, xa, &mx));
  PetscCallAbort(PETSC_COMM_SELF, VecGetArrayWriteAndMemType(y, ya, &my));
  if (PetscMemTypeDevice(mx) && PetscMemTypeDevice(my)) return 1;
  if (PetscMemTypeHost(mx) && PetscMemTypeHost(my)) return 0;
  PetscCallAbort(PETSC_COMM_SELF, VecRestoreArrayReadAndMemType(x, xa));
  PetscCallAbort(PETSC_COMM_SELF, VecRestoreArrayWriteAndMemType(y, ya));
  PetscCallAbort(PETSC_COMM_SELF, VecGetArrayRead(x, xa));   /* host copies */
  PetscCallAbort(PETSC_COMM_SELF, VecGetArrayWrite(y, ya));
  return 2;
}
static void b200_restore_arrays(Vec x, Vec y, int kind, const PetscScalar **xa, PetscScalar **ya)
{
  if (kind == 2) {
    PetscCallVoid(VecRestoreArrayRead(x, xa));
    PetscCallVoid(VecRestoreArrayWrite(y, ya));
  } else {
    PetscCallVoid(VecRestoreArrayReadAndMemType(x, xa));
    PetscCallVoid(VecRestoreArrayWriteAndMemType(y, ya));
  }
}

/* PCApply: x, y are the Vecs of PCApply_AIR_Shell (src/PCAIR_Shell.F90:170-188) */
PETSC_INTERN void pflare_b200_apply_c(void **handle, void **aux_p, Vec *x, Vec *y)
{
  const PetscScalar *xa;
  PetscScalar       *ya;
  cudaStream_t       ps = NULL;
  B200Aux           *aux = (B200Aux *)*aux_p;
  const int          kind = b200_get_arrays(*x, *y, &xa, &ya);
  if (kind == 1) b200_order_in(*handle, aux, &ps);
  B200_CHECK(pflare_b200_apply(*handle, xa, ya, kind == 1));
  if (kind == 1) b200_order_out(*handle, aux, ps);
  b200_restore_arrays(*x, *y, kind, &xa, &ya);
}

/* PCPFLAREINV: y = mat_inverse * x (src/PCPFLAREINV.c:618-626) */
PETSC_INTERN void pflare_b200_inv_apply_c(void **handle, void **aux_p, Vec *x, Vec *y)
{
  const PetscScalar *xa;
  PetscScalar       *ya;
  cudaStream_t       ps = NULL;
  B200Aux           *aux = (B200Aux *)*aux_p;
  const int          kind = b200_get_arrays(*x, *y, &xa, &ya);
  if (kind == 1) b200_order_in(*handle, aux, &ps);
  B200_CHECK(pflare_b200_inv_apply(*handle, 1, PFLARE_B200_INV_AFF, xa, ya, kind == 1));
  if (kind == 1) b200_order_out(*handle, aux, ps);
  b200_restore_arrays(*x, *y, kind, &xa, &ya);
}

/* ---- hierarchy container dump (pflare_b200/petsc_io.py documents the object order): one binary viewer, MatView /
 * ISView / VecView back to back.  Called from the upload hook when -pc_air_b200_dump <file> is given; together with a
 * PCApply input / output pair written by the same viewer it is the vector-level parity pin of this library. */
PETSC_INTERN void pflare_b200_dump_ints_c(PetscViewer *viewer, PetscInt n, const PetscInt *vals)
{
  IS is;
  PetscCallVoid(ISCreateGeneral(PETSC_COMM_SELF, n, vals, PETSC_USE_POINTER, &is));
  PetscCallVoid(ISView(is, *viewer));
  PetscCallVoid(ISDestroy(&is));
}
PETSC_INTERN void pflare_b200_dump_mat_c(PetscViewer *viewer, Mat *A) { PetscCallVoid(MatView(*A, *viewer)); }
PETSC_INTERN void pflare_b200_dump_vec_c(PetscViewer *viewer, Vec *v) { PetscCallVoid(VecView(*v, *viewer)); }

PETSC_INTERN void pflare_b200_finalize_c(void **handle) { B200_CHECK(pflare_b200_finalize_setup(*handle)); }
PETSC_INTERN void pflare_b200_destroy_c(void **handle, void **aux_p)
{
  if (*handle) B200_CHECK(pflare_b200_destroy(handle));
  if (*aux_p) {
    B200Aux *aux = (B200Aux *)*aux_p;
    cudaEventDestroy(aux->ev_in);
    cudaEventDestroy(aux->ev_out);
    PetscCallVoid(PetscFree(aux));
    *aux_p = NULL;
  }
}
