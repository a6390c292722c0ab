"""GPU suite, multi-rank path: a partitioned hierarchy (PETSc MPIAIJ layout: diag/off-diag blocks +
garray per rank) must reproduce the serial oracle.  Runs on ONE GPU through the in-process rank group
(lockstep execution, ghost exchange = device copies), and -- when the box has >= 2 GPUs -- through
NCCL with one process per GPU (tests/dist_nccl_check.py under torchrun)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import hiergen
import oracle
import pflare_b200

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _oracle(H):
    return hiergen.feed(H, oracle.OracleAIR(H.no_levels))


def _cluster_apply(H, nranks, b, **opts):
    parts = hiergen.partition(H, nranks)
    cl = pflare_b200.ClusterAIR(H.no_levels, nranks)
    for k, v in opts.items():
        cl.set_option(k, v)
    cl.upload(parts)
    xs = cl.apply(hiergen.scatter_vector(b, parts[0].rangesV[0]))
    info = (cl.ranks[0].layout()[0], [r.stats() for r in cl.ranks])
    cl.close()
    return np.concatenate(xs), info


@pytest.mark.parametrize("nranks", [2, 3, 4])
@pytest.mark.parametrize("name", ["fd2d_64", "fd2d_mf_newton", "fd2d_fcf", "fd2d_idealW", "fd3d_10_lump", "dg_mf",
                                  "fd2d_diagAff", "fd2d_trunc_newton", "adv1d_makefile", "fd2d_ffcc_mf", "fd2d_mf_neumann",
                                  "fd2d_full", "fd2d_full_mf_newton"])
def test_partitioned_vcycle_matches_serial_oracle(built_libs, name, nranks):
    A, H = cases.build(name)
    b = cases.rhs(A.shape[0])
    xo = _oracle(H).apply(b)
    for agg_rows in (0, 600, 10 ** 9):       # fully distributed | coarse levels on rank 0 | everything below level 1 on rank 0
        x, (l_agg, stats) = _cluster_apply(H, nranks, b, agg_rows=agg_rows, dense_rows=0 if nranks == 3 else 4096, p2p={4: 1, 2: 2}.get(nranks, 0))
        assert cases.rel_l2(x, xo) <= TOL, (name, nranks, agg_rows)
        if agg_rows == 0:
            assert l_agg == H.no_levels + 1
        if agg_rows == 10 ** 9:
            assert l_agg == 2
        assert all(s["exchange_groups"] > 0 for s in stats)


@pytest.mark.parametrize("opts", [dict(graph=0), dict(kernel=0), dict(fuse=0), dict(kernel=1), dict(dense_rows=0), dict(pdl=0), dict(p2p=1), dict(p2p=1, graph=0), dict(overlap=0), dict(overlap=0, graph=0),
                                  dict(epi_classes=0), dict(engine=0, p2p=1), dict(max_ctas=1), dict(kernel=1, p2p=1),
                                  dict(p2p=2), dict(p2p=2, graph=0), dict(p2p=2, engine=0), dict(p2p=2, kernel=1), dict(p2p=2, kernel=0), dict(p2p=2, max_ctas=1), dict(p2p=2, fuse=0)], ids=str)
def test_partitioned_execution_modes(built_libs, opts):
    A, H = cases.build("fd2d_64")
    b = cases.rhs(A.shape[0], seed=3)
    xo = _oracle(H).apply(b)
    x, _ = _cluster_apply(H, 3, b, agg_rows=300, **opts)
    assert cases.rel_l2(x, xo) <= TOL


def test_repeated_applies_reuse_ghost_buffers_safely(built_libs):
    """The peer-memory exchange reuses every ghost buffer across exchanges and cycles (epoch flags + acks)."""
    A, H = cases.build("fd2d_mf_newton")      # matrix-free Newton smoothing: the A_ff plan is exchanged many times per cycle
    parts = hiergen.partition(H, 4)
    cl = pflare_b200.ClusterAIR(H.no_levels, 4)
    cl.set_option("agg_rows", 300)
    cl.set_option("p2p", 1)
    cl.upload(parts)
    O = _oracle(H)
    for seed in range(5):
        b = cases.rhs(A.shape[0], seed=seed)
        xs = cl.apply(hiergen.scatter_vector(b, parts[0].rangesV[0]))
        assert cases.rel_l2(np.concatenate(xs), O.apply(b)) <= TOL
    cl.close()


def test_repeated_applies_with_the_fused_push(built_libs):
    """p2p=2: one ghost buffer per exchange instance, ready[] flags per instance and one started[] flag per cycle (no acks);
    in an in-process group the pushes run as stand-alone launches of the same device code (kernels.cuh: ghost_push)."""
    A, H = cases.build("fd2d_mf_newton")
    parts = hiergen.partition(H, 4)
    cl = pflare_b200.ClusterAIR(H.no_levels, 4)
    cl.set_option("agg_rows", 300)
    cl.set_option("p2p", 2)
    cl.upload(parts)
    O = _oracle(H)
    for seed in range(5):
        b = cases.rhs(A.shape[0], seed=seed)
        xs = cl.apply(hiergen.scatter_vector(b, parts[0].rangesV[0]))
        assert cases.rel_l2(np.concatenate(xs), O.apply(b)) <= TOL
    cl.close()


def test_more_ranks_than_coarse_rows(built_libs):
    """Ranks that own zero rows on the coarse levels (PETSc allows empty ranks; the reference's processor
    agglomeration produces them on purpose, src/AIR_Data_Type.F90:56-76)."""
    A, H = cases.build("fd2d_25")
    b = cases.rhs(A.shape[0])
    xo = _oracle(H).apply(b)
    for agg_rows in (0, 100):
        for p2p in (0, 2):
            x, _ = _cluster_apply(H, 7, b, agg_rows=agg_rows, p2p=p2p)
            assert cases.rel_l2(x, xo) <= TOL


@pytest.mark.parametrize("name", ["inv_newton_10_o50", "inv_arnoldi_asm", "inv_neumann_mf", "inv_power_mf"])
def test_partitioned_pflareinv(built_libs, name):
    A, H = cases.build_inv(name)
    x = cases.rhs(A.shape[0])
    yo = _oracle(H).inv_apply(1, oracle.INV_AFF, x)
    parts = hiergen.partition(H, 3)
    cl = pflare_b200.ClusterAIR(1, 3)
    cl.upload(parts)
    ys = cl.inv_apply(1, pflare_b200.INV_AFF, hiergen.scatter_vector(x, parts[0].rangesV[0]))
    assert cases.rel_l2(np.concatenate(ys), yo) <= TOL
    cl.close()


def test_garray_round_trip_bit_exact(built_libs):
    A, H = cases.build("fd2d_64")
    parts = hiergen.partition(H, 3)
    cl = pflare_b200.ClusterAIR(H.no_levels, 3)
    cl.upload(parts)
    for r, lh in enumerate(parts):
        for l, lv in enumerate(lh.levels, start=1):
            for which, op in lv["ops"].items():
                assert np.array_equal(cl.ranks[r].get_garray(l, which), op.garray)
    cl.close()


def test_nccl_two_processes(built_libs):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "tests", "dist_nccl_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "NCCL_DIST_OK" in out.stdout
