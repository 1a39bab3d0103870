"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for the rendezvous, NCCL inside libhdd_b200 for the
data path (coupling-face halo of the CG direction + dot-product all-reduces).

BlockSWIPDG's subdomains are dealt to the ranks in contiguous slabs (subdomain-major cell numbering makes every slab a
contiguous cell / DoF range, discretizations/block-swipdg.hh:1042)."""
import ctypes as C

import numpy as np

from . import capi


def subdomain_slabs(n_subdomains, world_size):
    """subdomain offsets per rank: rank r owns subdomains [slabs[r], slabs[r+1])"""
    return [n_subdomains * r // world_size for r in range(world_size + 1)]


def rank_cell_offsets(grid, world_size):
    """cell offsets per rank for `grid` (whole subdomains per rank)"""
    off = grid.subdomain_cell_offsets()
    slabs = subdomain_slabs(len(off) - 1, world_size)
    return np.array([off[s] for s in slabs], dtype=np.int64)


def partition_plan(grid, world_size, rank, offsets=None, local=False):
    """host-side halo / send plan of `rank` (hdd_partition_plan): (halo_cells, {peer: send_cells}).
    local: the owned-side walk hdd_mesh_create uses on N > 1 ranks (hdd_partition_plan_local) - same result"""
    L = capi.lib()
    off = rank_cell_offsets(grid, world_size) if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
    halo_p, send_p = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
    n_halo = C.c_int64()
    send_off = np.zeros(world_size + 1, np.int64)
    if local:
        capi.check(L.hdd_partition_plan_local(grid.kind, C.c_int64(grid.n_cells), capi.ptr(grid.cell_verts, C.c_int32),
                                              capi.ptr(grid.cell_neigh, C.c_int32), world_size, capi.ptr(off, C.c_int64),
                                              rank, C.byref(halo_p), C.byref(n_halo), C.byref(send_p),
                                              capi.ptr(send_off, C.c_int64)))
    else:
        capi.check(L.hdd_partition_plan(grid.kind, C.c_int64(grid.n_cells), C.c_int64(grid.n_verts),
                                        capi.ptr(grid.cell_verts, C.c_int32), world_size, capi.ptr(off, C.c_int64), rank,
                                        C.byref(halo_p), C.byref(n_halo), C.byref(send_p), capi.ptr(send_off, C.c_int64)))
    try:
        halo = np.array([halo_p[i] for i in range(n_halo.value)], dtype=np.int32)
        flat = np.array([send_p[i] for i in range(send_off[-1])], dtype=np.int32)
    finally:
        L.hdd_free(halo_p)
        L.hdd_free(send_p)
    send = {r: flat[send_off[r]:send_off[r + 1]] for r in range(world_size) if send_off[r + 1] > send_off[r]}
    return halo, send


class Comm:
    """one NCCL communicator per process (hdd_comm); pass it to every discretization of this process"""

    def __init__(self, uid, rank, world_size, device):
        self.rank, self.world_size, self.device = rank, world_size, device
        self.handle = C.c_void_p()
        capi.check(capi.lib().hdd_comm_create(uid, rank, world_size, device, C.byref(self.handle)))

    def __del__(self):
        try:
            if self.handle:
                capi.lib().hdd_comm_destroy(self.handle)
        except Exception:
            pass


def init_comm(rank, world_size, device=0):
    """NCCL unique id from rank 0, broadcast over the already initialised torch.distributed group.
    Returns the Comm the discretization constructors take, or None for a single process."""
    if world_size == 1:
        return None
    import torch
    import torch.distributed as dist
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        capi.check(capi.lib().hdd_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        uid = uid.cuda(device)
    dist.broadcast(uid, 0)
    return Comm(bytes(uid.cpu().tolist()), rank, world_size, device)
