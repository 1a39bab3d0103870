"""N > 1 host path on CPU: world_size-2 (and 3) gloo groups exercise the partition plan of each rank the way the
GPU path uses it - slab ownership, halo = [lower | owned | upper] local layout, send lists packed by the sender and
received as one contiguous range per peer - by running the halo exchange with gloo send/recv on a global test vector
and a distributed oracle SpMV that must equal the serial one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dune_hdd_b200 import grids, parallel
        from oracle import oracle as o
        g = (grids.simplex if kind == "alu" else grids.cube)(n, partitions=(4, 4))
        nl = g.n_loc
        off = parallel.rank_cell_offsets(g, world)
        halo, send = parallel.partition_plan(g, world, rank)
        c0, c1 = int(off[rank]), int(off[rank + 1])
        own = np.arange(c0, c1)
        local_cells = np.concatenate([halo[halo < c0], own, halo[halo >= c1]])
        assert np.all(np.diff(local_cells) > 0)  # local order == global order => block order is preserved
        own0 = int((halo < c0).sum())
        # global test vector: value = global DoF id; owned part known, halo part to be received
        x_local = np.full(local_cells.shape[0] * nl, np.nan)
        x_local[own0 * nl:(own0 + len(own)) * nl] = np.arange(c0 * nl, c1 * nl, dtype=float)
        owner = np.searchsorted(off, local_cells, side="right") - 1
        reqs = []
        recv_bufs = {}
        for peer in sorted(set(owner) - {rank}):
            sel = np.where(owner == peer)[0]
            assert np.all(np.diff(sel) == 1)  # one contiguous receive range per peer
            recv_bufs[peer] = (sel[0], torch.empty(len(sel) * nl, dtype=torch.float64))
            reqs.append(dist.irecv(recv_bufs[peer][1], src=peer))
        for peer, cells in send.items():
            loc = cells - c0 + own0
            idx = (loc[:, None] * nl + np.arange(nl)[None, :]).ravel()
            reqs.append(dist.isend(torch.from_numpy(x_local[idx].copy()), dst=peer))  # K11 pack
        for r in reqs:
            r.wait()
        for peer, (first, buf) in recv_bufs.items():
            x_local[first * nl:(first + buf.numel() // nl) * nl] = buf.numpy()
        expect = (local_cells[:, None] * nl + np.arange(nl)[None, :]).ravel().astype(float)
        assert np.array_equal(x_local, expect)
        # distributed SpMV with the oracle matrix rows of this rank == rows of the serial product
        m = o.Mesh(o.SIMPLEX if kind == "alu" else o.CUBE, g.xy, g.cell_verts, g.cell_neigh)
        rp, col = o.pattern(m)
        A = o.assemble_lhs(m, o.const(1.0), None, rp, col)
        xg = np.cos(np.arange(g.n_dofs, dtype=float))
        y_serial = o.spmv(rp, col, A, xg)
        g2l = -np.ones(g.n_cells, dtype=np.int64)
        g2l[local_cells] = np.arange(len(local_cells))
        xl = xg[(local_cells[:, None] * nl + np.arange(nl)[None, :]).ravel()]
        y = np.zeros((c1 - c0) * nl)
        for row in range(c0 * nl, c1 * nl):
            cols = col[rp[row]:rp[row + 1]]
            lc = g2l[cols // nl]
            assert np.all(lc >= 0)  # every column of an owned row is owned or in the halo
            y[row - c0 * nl] = A[rp[row]:rp[row + 1]] @ xl[lc * nl + cols % nl]
        assert np.allclose(y, y_serial[c0 * nl:c1 * nl], rtol=1e-14, atol=1e-14)
        # dot products: local partial + all_reduce == serial
        part = torch.tensor([float(xg[c0 * nl:c1 * nl] @ y)], dtype=torch.float64)
        dist.all_reduce(part)
        assert abs(part.item() - xg @ y_serial) <= 1e-12 * abs(xg @ y_serial)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,n,world", [("alu", 4, 2), ("sgrid", 8, 2), ("sgrid", 8, 3)])
def test_partition_and_halo_exchange_over_gloo(kind, n, world):
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, kind, n, out), nprocs=world, join=True)
    assert sorted(out.keys()) == list(range(world))


@pytest.mark.parametrize("kind,n,parts", [("alu", 8, (4, 4)), ("alu", 16, (8, 8)), ("sgrid", 32, (8, 8)), ("sgrid", 24, (4, 4)),
                                          ("sgrid", 12, (1, 1)), ("alu", 6, (3, 2))])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_owned_side_halo_walk_equals_the_full_sweep(kind, n, parts, world):
    """hdd_mesh_create on N > 1 ranks finds halo and send lists from the owned side (walk over the partition boundary);
    the device-free full sweep of hdd_partition_plan is the specification"""
    from dune_hdd_b200 import grids, parallel
    g = (grids.simplex if kind == "alu" else grids.cube)(n, partitions=parts)
    if g.n_subdomains < world:
        # uneven hand-made ranges (any contiguous split is allowed by the plan functions)
        off = np.linspace(0, g.n_cells, world + 1).astype(np.int64)
    else:
        off = parallel.rank_cell_offsets(g, world)
    for rank in range(world):
        halo, send = parallel.partition_plan(g, world, rank, offsets=off)
        halo_l, send_l = parallel.partition_plan(g, world, rank, offsets=off, local=True)
        assert np.array_equal(halo, halo_l), (rank, len(halo), len(halo_l))
        assert sorted(send) == sorted(send_l)
        for peer in send:
            assert np.array_equal(send[peer], send_l[peer]), (rank, peer)
