"""CPU suite, part 3: the multi-rank HOST logic of the product (no GPU): MPIAIJ partitioning, ghost
plans, agglomeration layout -- through the in-process rank group (threads) and through a world_size-2
gloo group (torchrun) in host-only planning mode."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import cases
import hiergen
import pflare_b200
from dist_emul import distributed_products

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _reassemble(parts, l, which, shape):
    blocks = []
    for lh in parts:
        op = lh.levels[l - 1]["ops"][which]
        d = op.diag.tocoo()
        rows, cols, vals = [d.row], [d.col + op.cstart], [d.data]
        if op.offdiag is not None:
            o = op.offdiag.tocoo()
            rows.append(o.row); cols.append(op.garray[o.col]); vals.append(o.data)
        blocks.append(sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                    shape=(op.diag.shape[0], shape[1])))
    return sp.vstack(blocks).tocsr()


@pytest.mark.parametrize("nranks", [2, 3, 5])
def test_partition_reassembles_bit_exactly(nranks):
    A, H = cases.build("fd2d_fcf")
    parts = hiergen.partition(H, nranks)
    for l, lv in enumerate(H.levels, start=1):
        for which, M in ((pflare_b200.AFF, lv.A_ff), (pflare_b200.AFC, lv.A_fc), (pflare_b200.ACF, lv.A_cf),
                         (pflare_b200.ACC, lv.A_cc), (pflare_b200.R, lv.R), (pflare_b200.P, lv.P)):
            B = _reassemble(parts, l, which, M.shape)
            assert (B != M).nnz == 0 and B.nnz == M.nnz
        # index sets: local lists + rstart reproduce the serial lists
        isf = np.concatenate([lh.levels[l - 1]["is_fine"] + lh.levels[l - 1]["rstart"] for lh in parts])
        assert np.array_equal(isf, lv.is_fine)
        # garray sorted, strictly increasing, never a local column (PETSc's convention)
        for lh in parts:
            for op in lh.levels[l - 1]["ops"].values():
                g = op.garray
                assert np.all(np.diff(g) > 0)
                assert not np.any((g >= op.cstart) & (g < op.cstart + op.diag.shape[1]))


@pytest.mark.parametrize("name,nranks,agg_rows", [("fd2d_64", 2, 0), ("fd2d_64", 4, 700), ("fd3d_10_lump", 3, 0),
                                                   ("fd2d_fcf", 3, 200), ("dg_mf", 2, 10 ** 9), ("fd2d_25", 7, 0)])
def test_ghost_plans_in_process_group(built_libs, name, nranks, agg_rows):
    """Host-only planning group: plans are symmetric and drive correct distributed products."""
    A, H = cases.build(name)
    parts = hiergen.partition(H, nranks)
    cl = pflare_b200.ClusterAIR(H.no_levels, nranks, device=-1)
    cl.set_option("agg_rows", agg_rows)
    cl.upload(parts)
    l_agg, rows = cl.ranks[0].layout()
    assert list(rows) == H.sizes()
    if agg_rows == 0:
        assert l_agg == H.no_levels + 1
    else:
        assert l_agg == next((l for l in range(2, H.no_levels + 1) if H.sizes()[l - 1] <= agg_rows), H.no_levels + 1)
    nl = min(l_agg, H.no_levels)
    for l in range(1, nl):
        for which in (pflare_b200.AFF, pflare_b200.AFC, pflare_b200.R, pflare_b200.P):
            plans = [r.ghost_plan(l, which) for r in cl.ranks]
            for p in range(nranks):
                for q in range(nranks):
                    assert plans[p]["send_count"][q] == plans[q]["recv_count"][p]
                assert plans[p]["send_idx"].size == plans[p]["send_count"].sum()

    # the plans drive correct distributed products (exchange emulated in-process: rank r receives
    # exactly what rank p packed for it)
    _PACKS.clear()   # keyed by id(cl): must not survive from a previous (garbage-collected) group
    worst = max(_products_with_lookup(H, parts, r, nranks, cl, l_agg) for r in range(nranks))
    assert worst < 1e-13
    # no GPU bound: apply must refuse
    with pytest.raises(pflare_b200.PflareB200Error):
        cl.apply([np.zeros(p.local_rows()) for p in parts])
    cl.close()


def _products_with_lookup(H, parts, rank, world, cl, l_agg):
    """distributed_products with an exchange that looks the peers' packed chunks up directly."""
    state = {}

    def exchange(sendbufs, recvcounts):
        # called once per (level, operator) in a fixed order; replay the same order on every peer
        k = state.setdefault("k", 0)
        state["k"] = k + 1
        out = []
        for p in range(world):
            if p == rank or not recvcounts[p]:
                out.append(np.zeros(0))
                continue
            out.append(_PACKS[(id(cl), p)][k][rank])
        return out

    # make sure every peer's packs exist
    for p in range(world):
        key = (id(cl), p)
        if key not in _PACKS:
            rec = []
            distributed_products(H, parts, p, world, cl.ranks[p], l_agg,
                                 exchange=lambda sendbufs, recvcounts, rec=rec: (rec.append(sendbufs), [np.zeros(int(c)) for c in recvcounts])[1])
            _PACKS[key] = rec
    return distributed_products(H, parts, rank, world, cl.ranks[rank], l_agg, exchange)


_PACKS = {}


def test_work_model_matches_reference_nnz_count(built_libs):
    """pflare_b200_get_stats[2] == nnzs_air_v of src/AIR_MG_Stats.F90:79-252 minus the identity blocks of R and P
    (never stored here); checked on a host-only planning context."""
    for name in ("fd2d_64", "fd2d_mf_arnoldi", "fd2d_fcf", "fd2d_jacobi"):
        A, H = cases.build(name)
        tot = 0
        for lv in H.levels:
            for s in lv.smooth_order:
                if s == 0:
                    break
                its = abs(s)
                inv, Ad, Aoff = (lv.inv_A_ff, lv.A_ff, lv.A_fc) if s > 0 else (lv.inv_A_cc, lv.A_cc, lv.A_cf)
                if inv.kind == "poly":
                    co = np.asarray(inv.coeffs)
                    ninv = int(np.count_nonzero(co[:-1, 0])) * Ad.nnz         # Horner: one product per non-zero lower coefficient
                else:
                    ninv = inv.nnz()
                tot += its * (ninv + Ad.nnz) + Aoff.nnz
            tot += (lv.R.nnz - lv.is_coarse.size) + (lv.P.nnz - lv.is_coarse.size)
        tot += H.inv_coarse.nnz()
        cl = pflare_b200.ClusterAIR(H.no_levels, 1, device=-1)
        hiergen.feed(H, cl.ranks[0])
        cl.finalize()
        assert cl.ranks[0].stats()["nnz_per_cycle"] == tot, name
        cl.close()


def test_gloo_two_processes(built_libs):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_gloo_check.py")]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "GLOO_DIST_OK" in out.stdout


def test_group_finalize_fails_instead_of_hanging_when_one_rank_is_incomplete(built_libs):
    """A rank that returns early from finalize (level never set) must not leave the others blocked in a setup collective."""
    A, H = cases.build("fd2d_25")
    parts = hiergen.partition(H, 3)
    cl = pflare_b200.ClusterAIR(H.no_levels, 3, device=-1)
    for r, lh in enumerate(parts):
        if r == 1:
            continue            # rank 1 uploads nothing
        lh.feed(cl.ranks[r])
    with pytest.raises(pflare_b200.PflareB200Error) as e:
        cl.finalize()
    assert "rank 1" in str(e.value)
    cl.close()


@pytest.mark.parametrize("nranks", [1, 3])
def test_resetup_on_a_live_handle_host_logic(built_libs, nranks):
    """The reference's re-setup (src/PCAIR_Shell.F90:148-162, SAME_NONZERO_PATTERN): the upload hook runs again on the SAME handle
    and finalize_setup rebuilds from the new host operators (host-only planning group: the plans, counters and programme of the
    second setup equal those of a fresh handle fed with the same operators; the device side of the path is covered by test_gpu_parity.py::test_resetup_on_a_live_handle)."""
    A, H = cases.build("fd2d_64")
    H2 = hiergen.build_hierarchy(1.7 * A, cases.CASES["fd2d_64"]()[1])
    ref = pflare_b200.ClusterAIR(H2.no_levels, nranks, device=-1)
    ref.upload(hiergen.partition(H2, nranks))
    want = [r.stats() for r in ref.ranks]
    ref.close()
    cl = pflare_b200.ClusterAIR(H.no_levels, nranks, device=-1)
    cl.upload(hiergen.partition(H, nranks))
    first = [r.stats() for r in cl.ranks]
    cl.upload(hiergen.partition(H2, nranks))            # second setup on the live handles
    again = [r.stats() for r in cl.ranks]
    for a, w in zip(again, want):
        for k in ("kernel_launches", "exchange_groups", "nnz_per_cycle", "algorithmic_bytes", "ghost_bytes_sent"):
            assert a[k] == w[k], k
    assert sum(a["nnz_per_cycle"] for a in again) != sum(f["nnz_per_cycle"] for f in first)      # the new operators really replaced the old ones
    cl.close()
