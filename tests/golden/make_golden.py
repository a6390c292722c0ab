"""Generate the committed fixtures of tests/golden/ (run in the build container, where
/root/reference is mounted; the fixtures travel to the GPU box, the reference does not).

  * <case>.npz          : a whole hierarchy (every operator the upload hook hands over), the seeded
                          rhs and the oracle's PCApply output, for the cases in tests/cases.py:GOLDEN.
  * <case>.petsc        : the same container as PETSc binary objects (pflare_b200/petsc_io.py), the format a PFLARE build
                          dumps through shim/pflare_b200_petsc.c; holds b and the oracle's x as the trailing Vec pair.
  * ilu_mat_stream.npz  : BASELINE.json configs[4] -- L and U from the ParILU(0) sweep of
                          /root/reference/tests/ilu_factors.c:484-567 on the reference's own fixture
                          /root/reference/tests/data/mat_stream_2364 (PETSc binary), the Newton-basis
                          roots of PCPFLAREINV (order 6, matrix-free, ilu_factors.c:122-126), rhs and
                          the oracle's PCApply output.

  * mat_stream_2364_system.npz : the reference's fixture tests/data/mat_stream_2364 (AIJ Mat + rhs Vec, PETSc binary) converted to
                          npz: the matrix and rhs of the `ex12f` / `ex6 -f data/mat_stream_2364 ... -ksp_max_it 5` runs of
                          tests/Makefile:89-99,113,208 (cases `ms2364_*` in tests/cases.py).  (`make_golden.py mat_stream`)
  * e05r0100_system.npz : the reference's fixture tests/data/e05r0100_petsc converted to npz (tests/Makefile:157: `ex6 -f
                          data/e05r0100_petsc -b_in_f 0 -pc_air_a_drop 1e-3 -pc_air_inverse_type power -ksp_max_it 26`).
  * bus1138_newton.npz  : the reference's fixture tests/data/1138_bus with the high-order Newton-basis GMRES polynomials of
                          tests/Makefile:199-205 (PCPFLAREINV newton, matrix-free, order 60 and 120 "with added roots",
                          src/Gmres_Poly_Newton.F90:630-700): matrix, roots, a seeded initial guess, the oracle's apply.
                          (`python tests/golden/make_golden.py bus1138` regenerates only this file.)

The reference holds no vector-level golden outputs for PCApply, so the stored
outputs are the ORACLE's (pinned by the iteration-count bounds of tests/test_oracle_pins.py); they
freeze the oracle against drift and give the GPU tests fixed inputs independent of hiergen's RNG.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import hiergen  # noqa: E402
import oracle  # noqa: E402
from hiergen import io as hio, poly  # noqa: E402


def make_bus1138():
    mats, _ = hiergen.read_petsc_binary("/root/reference/tests/data/1138_bus")
    A = mats[0].tocsr()
    A.sort_indices()
    n = A.shape[0]
    out = {"indptr": A.indptr.astype(np.int32), "indices": A.indices.astype(np.int32), "data": A.data, "x0": np.random.default_rng(0).random(n),
           "v": cases.rhs(n)}
    for order in (60, 120):
        H = hiergen.build_pflareinv(A, poly.NEWTON, order, 1, True)
        O = hiergen.feed(H, oracle.OracleAIR(1))
        out["roots_%d" % order] = np.asarray(H.inv_coarse.coeffs)
        out["y_oracle_%d" % order] = O.inv_apply(1, oracle.INV_AFF, out["v"])
        print("1138_bus order", order, "roots", out["roots_%d" % order].shape)
    np.savez_compressed(os.path.join(HERE, "bus1138_newton.npz"), **out)


def make_e05r():
    mats, _ = hiergen.read_petsc_binary("/root/reference/tests/data/e05r0100_petsc")
    A = mats[0].tocsr()
    A.sort_indices()
    np.savez_compressed(os.path.join(HERE, "e05r0100_system.npz"), indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32), data=A.data)
    print("e05r0100_system", A.shape[0], A.nnz)


def make_mat_stream():
    mats, vecs = hiergen.read_petsc_binary("/root/reference/tests/data/mat_stream_2364")
    A = mats[0].tocsr()
    A.sort_indices()
    np.savez_compressed(os.path.join(HERE, "mat_stream_2364_system.npz"), indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32),
                        data=A.data, b=np.asarray(vecs[0], dtype=np.float64))
    print("mat_stream_2364_system", A.shape[0], A.nnz)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "bus1138":
        return make_bus1138()
    if len(sys.argv) > 1 and sys.argv[1] == "mat_stream":
        return make_mat_stream()
    if len(sys.argv) > 1 and sys.argv[1] == "e05r":
        return make_e05r()
    make_bus1138()
    make_mat_stream()
    make_e05r()
    for name in cases.GOLDEN:
        A, H = cases.build(name)
        O = hiergen.feed(H, oracle.OracleAIR(H.no_levels))
        b = cases.rhs(A.shape[0])
        x = O.apply(b)
        hio.save(os.path.join(HERE, name + ".npz"), H, b=b, x_oracle=x)
        print(name, A.shape[0], H.no_levels, os.path.getsize(os.path.join(HERE, name + ".npz")))
    # the same container in PETSc binary (pflare_b200/petsc_io.py: what a PFLARE build would dump), incl. a full-smoothing case
    from pflare_b200 import petsc_io
    for name in ("fd2d_25", "fd2d_full_mf"):
        A, H = cases.build(name)
        O = hiergen.feed(H, oracle.OracleAIR(H.no_levels))
        b = cases.rhs(A.shape[0])
        petsc_io.save_hierarchy(os.path.join(HERE, name + ".petsc"), H, b, O.apply(b))
        print(name + ".petsc", os.path.getsize(os.path.join(HERE, name + ".petsc")))
    ref = "/root/reference/tests/data/mat_stream_2364"
    mats, _ = hiergen.read_petsc_binary(ref)
    L, U, idu, sweeps = hiergen.parilu_factors(mats[0])
    out = {"sweeps": np.asarray(sweeps)}
    b = cases.rhs(L.shape[0])
    out["b"] = b
    for nm, T in (("L", L), ("U", U)):
        H = hiergen.build_pflareinv(T, poly.NEWTON, 6, 1, True)
        O = hiergen.feed(H, oracle.OracleAIR(1))
        out[nm + "_indptr"], out[nm + "_indices"], out[nm + "_data"] = T.indptr, T.indices, T.data
        out[nm + "_roots"] = H.inv_coarse.coeffs
        out[nm + "_y_oracle"] = O.inv_apply(1, oracle.INV_AFF, b)
    np.savez_compressed(os.path.join(HERE, "ilu_mat_stream.npz"), **out)
    print("ilu_mat_stream", L.shape[0], sweeps)


if __name__ == "__main__":
    main()
