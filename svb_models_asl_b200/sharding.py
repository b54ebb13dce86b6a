"""
Voxel sharding over the GPUs of one box (one process per GPU, torch.distributed).

Voxel order is C-order over the mask with z fastest (aslrest.py:438-443), so a contiguous range of voxels is
an x-slab.  Voxel-wise priors need nothing else.  The spatial ("M") prior couples each voxel with its 6
neighbours, so a shard additionally holds a *halo*: the local arrays cover the contiguous global range

        [lo - halo_lo | lo .. hi | hi + halo_hi)

(`global id = lo - halo_lo + local index`, also for halo voxels - that is what lets every rank draw a halo
voxel's Philox stream itself).  Per iteration the halo's posterior STATE rows are exchanged with the two
neighbouring ranks (a few floats per boundary voxel); samples are never exchanged.

Everything here is host logic on torch tensors and works identically on CPU tensors with the gloo backend,
which is how the CPU test tier exercises the N > 1 path (tests/test_sharding_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as td


def shard_bounds(n, rank, world):
    """Contiguous voxel range [lo, hi) of `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world_info():
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


class ShardPlan:
    """Layout of one rank's shard.  `neighbours` is the GLOBAL table [W,6] (-1 = none), a callable
    `neighbours(lo, hi) -> [hi-lo, 6]` giving the rows of a voxel range only (DataModel.neighbour_table; a rank then
    never builds the table of the whole volume), or None."""

    def __init__(self, n_vox, rank, world, neighbours=None):
        self.n_global, self.rank, self.world = int(n_vox), rank, world
        self.lo, self.hi = shard_bounds(n_vox, rank, world)
        self.n_own = self.hi - self.lo
        self.halo_lo = self.halo_hi = 0
        self.neighbours_local = None
        if neighbours is not None:
            rows = neighbours if callable(neighbours) else (lambda lo, hi: np.asarray(neighbours)[lo:hi])
            nb = np.asarray(rows(self.lo, self.hi))
            self.halo_lo, self.halo_hi = ShardPlan._halo(nb, self.lo, self.hi)
            base = self.lo - self.halo_lo
            local = np.where(nb >= 0, nb - base, -1).astype(np.int32)
            table = -np.ones((self.ld, 6), dtype=np.int32)
            table[self.halo_lo:self.halo_lo + self.n_own] = local
            self.neighbours_local = np.ascontiguousarray(table.T)           # [6, ld]
        # halo sizes of the adjacent ranks (what they need from us)
        self.prev_halo_hi = self.next_halo_lo = 0
        if world > 1 and neighbours is not None:
            if rank > 0:
                lo_p = shard_bounds(n_vox, rank - 1, world)
                self.prev_halo_hi = ShardPlan._halo(np.asarray(rows(*lo_p)), *lo_p)[1]
            else:
                lo_p = (0, 0)
            if rank < world - 1:
                lo_n = shard_bounds(n_vox, rank + 1, world)
                self.next_halo_lo = ShardPlan._halo(np.asarray(rows(*lo_n)), *lo_n)[0]
            else:
                lo_n = (0, 0)
            if self.prev_halo_hi > self.n_own or self.next_halo_lo > self.n_own:
                raise ValueError("shard of %d voxels is thinner than the neighbouring rank's halo "
                                 "(use fewer GPUs for this volume)" % self.n_own)
            if self.halo_lo > lo_p[1] - lo_p[0] or self.halo_hi > lo_n[1] - lo_n[0]:
                raise ValueError("halo reaches beyond the adjacent rank (use fewer GPUs for this volume)")

    @staticmethod
    def _halo(nb, lo, hi):
        """(lower, upper) halo sizes of the voxel range [lo, hi) whose neighbour rows are `nb`."""
        below = nb[(nb >= 0) & (nb < lo)]
        above = nb[nb >= hi]
        return (int(lo - below.min()) if below.size else 0, int(above.max() + 1 - hi) if above.size else 0)

    @property
    def ld(self):
        return self.halo_lo + self.n_own + self.halo_hi

    @property
    def global_offset(self):
        """Global index of local index 0."""
        return self.lo - self.halo_lo

    @property
    def own(self):
        return slice(self.halo_lo, self.halo_lo + self.n_own)

    def take(self, array, axis=0):
        """Slice a global per-voxel array (numpy) to this rank's local range, halo included."""
        idx = [slice(None)] * np.ndim(array)
        idx[axis] = slice(self.global_offset, self.global_offset + self.ld)
        return np.ascontiguousarray(np.asarray(array)[tuple(idx)])

    # ------------------------------------------------------------------
    def exchange_halo(self, state):
        """Fill the halo columns of `state` [rows, ld] from the adjacent ranks' owned voxels.
        Grouped point-to-point sends/receives (ncclSend/ncclRecv under NCCL, plain gloo on CPU)."""
        if self.world == 1 or (self.halo_lo == 0 and self.halo_hi == 0 and self.prev_halo_hi == 0
                               and self.next_halo_lo == 0):
            return
        ops, keep, recv = [], [], []
        a, b = self.halo_lo, self.halo_lo + self.n_own
        if self.rank > 0:
            if self.prev_halo_hi:                       # rank-1 needs my first voxels as its upper halo
                buf = state[:, a:a + self.prev_halo_hi].contiguous()
                keep.append(buf)
                ops.append(td.P2POp(td.isend, buf, self.rank - 1))
            if self.halo_lo:                            # my lower halo = rank-1's last voxels
                buf = torch.empty((state.shape[0], self.halo_lo), dtype=state.dtype, device=state.device)
                recv.append((buf, slice(0, self.halo_lo)))
                ops.append(td.P2POp(td.irecv, buf, self.rank - 1))
        if self.rank < self.world - 1:
            if self.next_halo_lo:                       # rank+1 needs my last voxels as its lower halo
                buf = state[:, b - self.next_halo_lo:b].contiguous()
                keep.append(buf)
                ops.append(td.P2POp(td.isend, buf, self.rank + 1))
            if self.halo_hi:
                buf = torch.empty((state.shape[0], self.halo_hi), dtype=state.dtype, device=state.device)
                recv.append((buf, slice(b, b + self.halo_hi)))
                ops.append(td.P2POp(td.irecv, buf, self.rank + 1))
        if not ops:
            return
        for req in td.batch_isend_irecv(ops):
            req.wait()
        for buf, sl in recv:
            state[:, sl] = buf

    @staticmethod
    def allreduce_sum(tensor):
        """Global sums that every rank needs identically: cost (reporting) and d(cost)/d(log ak)."""
        if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
            td.all_reduce(tensor, op=td.ReduceOp.SUM)
        return tensor

    def gather_owned(self, local, axis=0):
        """Concatenate the owned part of per-shard numpy arrays on every rank (result assembly)."""
        if self.world == 1:
            return local
        parts = [None] * self.world
        td.all_gather_object(parts, local)
        return np.concatenate(parts, axis=axis)
