"""Host-side logic of the z-slab decomposition on CPU: partition arithmetic, and (world_size 2, gloo) halo exchange,
scalar all-reduce and level replication -- the same SlabComm code that runs over NCCL on the GPUs.  The numerical
stand-in for the kernels is scipy on the oracle's CSR rows of each slab (tests only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pymoto_b200.slab import SlabPartition, SlabComm


def test_partition_arithmetic():
    # 512x256x256 on 8 ranks: 32 planes per rank, levels split while >= 4 planes per rank and even boundaries
    parts = [SlabPartition(256, 8, r, n_levels=7) for r in range(8)]
    p = parts[3]
    assert p.m == 32 and p.n_dist == 4
    assert p.planes(0) == (96, 128) and p.planes(1) == (48, 64) and p.planes(3) == (12, 16)
    assert p.planes(4) == (0, 17)  # replicated: whole 16-element grid
    assert parts[7].planes(0) == (224, 257) and parts[7].elem_layers(0) == (224, 256)
    for lvl in range(4):  # slabs tile the planes of every distributed level, coarse plane K lives with fine plane 2K
        edges = [parts[r].planes(lvl) for r in range(8)]
        assert edges[0][0] == 0 and edges[-1][1] == (256 >> lvl) + 1
        assert all(edges[r][1] == edges[r + 1][0] for r in range(7))
        assert all(e[0] % 2 == 0 for e in edges)
    assert p.slab_planes(4) == (6, 8)
    # dof threshold keeps small levels replicated
    dofs = [3 * ((512 >> l) + 1) * ((256 >> l) + 1) * ((256 >> l) + 1) for l in range(7)]
    assert SlabPartition(256, 8, 0, n_levels=7, level_dofs=dofs, min_dofs=1_000_000).n_dist == 3
    assert SlabPartition(256, 8, 0, n_levels=7, level_dofs=dofs, min_dofs=4_000_000).n_dist == 2
    # the coarsest level is never split; single rank owns everything
    assert SlabPartition(16, 2, 0, n_levels=2).n_dist == 1
    s = SlabPartition(32, 1, 0, n_levels=3)
    assert s.planes(0) == (0, 33) and not s.is_distributed(0)
    with pytest.raises(ValueError):
        SlabPartition(30, 4, 0)
    with pytest.raises(ValueError):
        SlabPartition(6, 2, 0, n_levels=3)  # 3 planes per rank: odd boundary cannot carry a coarse level


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nx, ny, nz, ndof):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from oracle import Grid

        part = SlabPartition(nz, world, rank, n_levels=2)
        comm = SlabComm(part)
        k0, k1 = part.planes(0)
        plane = (nx + 1) * (ny + 1) * ndof
        n_loc = (k1 - k0) * plane
        rng = np.random.default_rng(0)  # same global data on every rank
        gr = Grid(nx, ny, nz)
        Ke = rng.standard_normal((8 * ndof, 8 * ndof))
        Ke = Ke + Ke.T
        K = oracle.assembly.Assembler(gr, Ke)(rng.random(gr.nel))
        xg = rng.standard_normal(gr.nnodes * ndof)

        # halo exchange: padded local vector, one plane on each side
        base = torch.zeros(n_loc + 2 * plane, dtype=torch.float64)
        base[plane:plane + n_loc] = torch.from_numpy(xg[k0 * plane:k1 * plane])
        comm.exchange(base, plane, n_loc, plane)
        lo, hi = max(k0 - 1, 0) * plane, min(k1 + 1, nz + 1) * plane
        window = base[plane - (k0 * plane - lo): plane + n_loc + (hi - k1 * plane)].numpy()
        assert np.array_equal(window, xg[lo:hi])
        # slab SpMV == rows of the global product (what pmb_spmv computes on each rank)
        y_loc = K[k0 * plane:k1 * plane, lo:hi] @ window
        assert np.allclose(y_loc, (K @ xg)[k0 * plane:k1 * plane], rtol=1e-13, atol=1e-13)
        assert K[k0 * plane:k1 * plane, :lo].nnz == 0 and K[k0 * plane:k1 * plane, hi:].nnz == 0  # one halo plane suffices

        # one-sided exchanges (restriction needs only the lower halo, prolongation only the upper one)
        base2 = torch.full_like(base, -7.0)
        base2[plane:plane + n_loc] = base[plane:plane + n_loc]
        comm.exchange(base2, plane, n_loc, plane, lower=True, upper=False)
        if part.lower is not None:
            assert np.array_equal(base2[:plane].numpy(), xg[(k0 - 1) * plane:k0 * plane])
        assert torch.all(base2[plane + n_loc:] == -7.0)
        # wide halos (density filter, radius 2 -> 2 element layers)
        lay = nx * ny
        e0, e1 = part.elem_layers(0)
        xe = rng.random(gr.nel)
        eb = torch.zeros((e1 - e0 + 4) * lay, dtype=torch.float64)
        eb[2 * lay:(2 + e1 - e0) * lay] = torch.from_numpy(xe[e0 * lay:e1 * lay])
        comm.exchange(eb, 2 * lay, (e1 - e0) * lay, lay, width=2)
        a, b = max(e0 - 2, 0), min(e1 + 2, nz)
        got = eb[(2 - (e0 - a)) * lay:(2 + e1 - e0 + (b - e1)) * lay].numpy()
        assert np.array_equal(got, xe[a * lay:b * lay])

        # global dot product = all-reduce of the slab partials
        d = torch.tensor([float(xg[k0 * plane:k1 * plane] @ xg[k0 * plane:k1 * plane])], dtype=torch.float64)
        comm.allreduce_(d)
        assert abs(d.item() - xg @ xg) <= 1e-12 * (xg @ xg)
        # replication of a coarse-level vector: every rank contributes its slab
        kc0, kc1 = part.slab_planes(1)
        planec = (nx // 2 + 1) * (ny // 2 + 1) * ndof
        vc = rng.standard_normal(planec * (nz // 2 + 1))
        full = torch.empty(vc.size, dtype=torch.float64)
        comm.gather_full(torch.from_numpy(vc[kc0 * planec:kc1 * planec].copy()), full, kc0 * planec)
        assert np.array_equal(full.numpy(), vc)
        assert comm.exchanges == 3 and comm.allreduces == 2
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,ndof", [((4, 3, 8), 3), ((5, 2, 4), 1)])
def test_halo_exchange_and_collectives_gloo_world2(shape, ndof):
    mp.spawn(_worker, args=(2, _free_port(), shape[0], shape[1], shape[2], ndof), nprocs=2, join=True)
