"""torchrun worker: one process per GPU, NCCL ghost exchange; every rank checks its rows of the
distributed V-cycle against the serial oracle.  Prints NCCL_DIST_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import hiergen  # noqa: E402
import oracle  # noqa: E402
import pflare_b200  # noqa: E402
from pflare_b200 import _capi  # noqa: E402
import ctypes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")

    def fresh_uid():
        """one NCCL unique id per communicator (= per PC), made on rank 0 and broadcast by the host communicator"""
        uid = [None]
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            _capi.check(_capi.lib().pflare_b200_get_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
            uid[0] = buf.raw
        dist.broadcast_object_list(uid, src=0)
        return uid[0]

    worst = 0.0
    for name in ("fd2d_64", "fd2d_mf_newton", "fd2d_fcf", "dg_mf"):
        A, H = cases.build(name)
        b = cases.rhs(A.shape[0])
        xo = hiergen.feed(H, oracle.OracleAIR(H.no_levels)).apply(b)
        parts = hiergen.partition(H, world)
        rg = parts[0].rangesV[0]
        # p2p=1: CUDA-IPC peer-memory exchange (push kernel + acks); p2p=2: push fused into the consuming SpMV kernel; p2p=0: NCCL send/recv, with (overlap=1) or without the side-stream split phase
        for agg_rows, p2p, overlap in ((0, 2, 1), (600, 2, 1), (0, 1, 1), (600, 1, 1), (600, 0, 1), (0, 0, 1), (600, 0, 0)):
            pc = pflare_b200.PC(rank=rank, nranks=world, unique_id=fresh_uid(), device=local).setType("air").setHierarchy(parts[rank])
            pc.setOption("agg_rows", agg_rows)
            pc.setOption("p2p", p2p)
            pc.setOption("overlap", overlap)
            for rep in range(3):                                 # repeated cycles reuse the ghost buffers (epochs / acks)
                x = pc.apply(b[rg[rank]:rg[rank + 1]] * (rep + 1))
                err = np.linalg.norm(x - (rep + 1) * xo[rg[rank]:rg[rank + 1]]) / np.linalg.norm((rep + 1) * xo)
                worst = max(worst, err)
            dist.barrier()
            pc.destroy()
    # outer Krylov method across the ranks (pflare_b200_ksp_solve: the cycle's fused exchange, the system matrix over NCCL
    # send/recv, the Gram-Schmidt scalars through ncclAllReduce): iteration count equal (+-1) to a host GMRES around the oracle
    from krylov import gmres
    from hiergen.partition import _split_cols
    for name, rtol, p2p in (("fd2d_64", 1e-8, 2), ("fd2d_mf_newton", 1e-8, 0)):
        A, H = cases.build(name)
        n = A.shape[0]
        parts = hiergen.partition(H, world)
        rg = parts[0].rangesV[0]
        O = hiergen.feed(H, oracle.OracleAIR(H.no_levels))
        _, its_cpu, conv_cpu = gmres(A, np.zeros(n), np.ones(n), O.apply, rtol=rtol, side="right")
        d = pflare_b200.DeviceAIR(H.no_levels, rank=rank, nranks=world, unique_id=fresh_uid(), device=local)
        d.set_option("p2p", p2p)
        dg, od, ga = _split_cols(A.tocsr()[rg[rank]:rg[rank + 1]], rg[rank], rg[rank + 1])
        d.ksp_set_operator(dg, od, ga, cstart=rg[rank])
        parts[rank].feed(d)
        xl, its, conv, rn = d.ksp_solve(np.zeros(rg[rank + 1] - rg[rank]), np.ones(rg[rank + 1] - rg[rank]), ksp_type="gmres", side="right", rtol=rtol)
        pieces = [None] * world
        dist.all_gather_object(pieces, xl)
        x = np.concatenate(pieces)
        res = np.linalg.norm(A @ x) / np.linalg.norm(A @ np.ones(n))
        if rank == 0:
            print("ksp %s: its %d (host around the oracle: %d), converged %s, relative residual %.2e" % (name, its, its_cpu, conv, res))
        assert conv and conv_cpu and abs(its - its_cpu) <= 1 and res <= 10 * rtol, (name, its, its_cpu, res)
        dist.barrier()
        d.close()
    t = torch.tensor([worst], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("worst relative error %.3e" % t.item())
        assert t.item() <= 1e-12
        print("NCCL_DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
