"""torchrun worker (CPU, gloo, world_size 2): every rank drives a host-only PLANNING context of the
product library (device = -1) whose setup-time communicator is torch.distributed/gloo behind the
MPI_Alltoall(v)-style callbacks of pflare_b200_set_host_exchange, then checks the resulting ghost plans
(a) against the in-process rank group and (b) by emulating one distributed SpMV per operator class with
gloo point-to-point traffic and comparing with the serial product.  Prints GLOO_DIST_OK on rank 0."""
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
import pflare_b200  # noqa: E402
from dist_emul import distributed_products, serial_products  # noqa: E402


def gloo_alltoall(send):
    world = dist.get_world_size()
    mine = torch.from_numpy(np.asarray(send, dtype=np.int64).copy())
    allv = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allv, mine)
    r = dist.get_rank()
    return np.array([int(allv[p][r]) for p in range(world)], dtype=np.int64)


def gloo_alltoallv(data, scnt, sdsp, rcnt, rdsp):
    world, r = dist.get_world_size(), dist.get_rank()
    # gather everybody's (counts, displacements, payload) -- fine for setup-sized messages
    obj = [None] * world
    dist.all_gather_object(obj, (list(scnt), list(sdsp), bytes(data)))
    tot = max([rdsp[p] + rcnt[p] for p in range(world)] + [0])
    out = bytearray(tot)
    for p in range(world):
        sc, sd, payload = obj[p]
        assert sc[r] == rcnt[p]
        out[rdsp[p]:rdsp[p] + rcnt[p]] = payload[sd[r]:sd[r] + sc[r]]
    return bytes(out)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    for name, agg_rows in (("fd2d_64", 0), ("fd2d_fcf", 300), ("fd3d_10_lump", 0), ("dg_mf", 500)):
        A, H = cases.build(name)
        parts = hiergen.partition(H, world)
        dev = pflare_b200.DeviceAIR(H.no_levels, rank=rank, nranks=world, device=-1)
        dev.set_option("agg_rows", agg_rows)
        dev.set_host_exchange(gloo_alltoall, gloo_alltoallv)
        parts[rank].feed(dev)                          # finalize_setup: collective over the gloo group
        # (a) the same plans as the in-process (threads) rank group
        cl = pflare_b200.ClusterAIR(H.no_levels, world, device=-1)
        cl.set_option("agg_rows", agg_rows)
        cl.upload(parts)
        l_agg = dev.layout()[0]
        assert l_agg == cl.ranks[rank].layout()[0]
        nl = min(l_agg, H.no_levels)
        for l in range(1, nl):
            for which in (pflare_b200.AFF, pflare_b200.AFC, pflare_b200.R, pflare_b200.P):
                p1, p2 = dev.ghost_plan(l, which), cl.ranks[rank].ghost_plan(l, which)
                for k in p1:
                    assert np.array_equal(p1[k], p2[k]), (name, l, which, k)
        # (b) the plans drive correct distributed products (exchange = gloo send/recv)
        err = distributed_products(H, parts, rank, world, dev, l_agg,
                                   exchange=lambda sendbufs, recvcounts: p2p_exchange(sendbufs, recvcounts))
        worst = max(worst, err)
        dev.close()
        cl.close()
    t = torch.tensor([worst], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("worst relative error of the emulated distributed products %.3e" % t.item())
        assert t.item() < 1e-13
        print("GLOO_DIST_OK")
    dist.destroy_process_group()


def p2p_exchange(sendbufs, recvcounts):
    """sendbufs[p] = float64 array for rank p; returns the arrays received from every rank."""
    world, r = dist.get_world_size(), dist.get_rank()
    reqs, out = [], [np.zeros(0)] * world
    bufs = []
    for p in range(world):
        if p == r:
            continue
        if len(sendbufs[p]):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(sendbufs[p])), p))
        if recvcounts[p]:
            t = torch.zeros(int(recvcounts[p]), dtype=torch.float64)
            bufs.append((p, t))
            reqs.append(dist.irecv(t, p))
    for q in reqs:
        q.wait()
    for p, t in bufs:
        out[p] = t.numpy()
    return out


if __name__ == "__main__":
    main()
