"""Row sharding of the grouping path over the GPUs of one box.

The reference is single-process (no NCCL/MPI anywhere); this is new.  One process per GPU, every
rank holds the full embedding matrix (2-12 GB against 180 GB of HBM) and computes a row shard of
every phase.  Exchange steps, all through torch.distributed (NCCL over NVLink; gloo in CPU tests):

  K-means prologue   all-reduce  column sums / sums of squares            2*d fp64
  k-means++ round    all-reduce  candidate potentials (T fp64)
                     all-gather  closest-distance shard (N fp32 in total)
  Lloyd iteration    all-reduce  centroid sums + counts + changed         (k*d + k + 1) fp64
  dedup              all-reduce(min) of the packed (distance, argmin) keys  N int64
  silhouette / CH / DB  all-reduce of n_labellings (or 2k + k*d) fp64 partial sums
"""
import torch


class Comm:
    """Thin wrapper: a no-op for one rank, torch.distributed collectives otherwise."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.active = dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group) if self.active else 0
        self.world = dist.get_world_size(group) if self.active else 1
        if self.world == 1:
            self.active = False

    @classmethod
    def single(cls):
        """A communicator of one rank regardless of the process group (engine.solo_if)."""
        c = cls.__new__(cls)
        import torch.distributed as dist
        c._dist, c.active, c.group, c.rank, c.world = dist, False, None, 0, 1
        return c

    def allreduce_sum(self, t):
        if self.active:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t

    def broadcast_object(self, box, src=0):
        """`box` is a one-element list; afterwards every rank holds rank `src`'s element (pickled, tiny)."""
        if self.active:
            self._dist.broadcast_object_list(box, src=src, group=self.group)
        return box

    def broadcast(self, t, src=0):
        """Every rank's tensor `t` takes rank `src`'s contents."""
        if self.active:
            self._dist.broadcast(t, src=self._global(src), group=self.group)
        return t

    def allreduce_min(self, t):
        if self.active:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.MIN, group=self.group)
        return t

    def allgather_rows(self, full, bounds):
        """`full` is a [N, ...] tensor of which this rank filled rows bounds[rank]:bounds[rank+1];
        afterwards every rank holds all rows.  Shards may be ragged.

        NCCL: equal shards gather IN PLACE (all_gather_into_tensor with the rank's own slice of `full` as input: no
        staging copy, one collective); ragged shards go through one padded buffer (one collective + `world` small
        copies) instead of `world` serial broadcasts.  Other back ends (gloo in the CPU tests) keep the list form."""
        if not self.active:
            return full
        sizes = [bounds[r + 1] - bounds[r] for r in range(self.world)]
        nccl = full.is_cuda and full.is_contiguous()
        if nccl and len(set(sizes)) == 1 and sizes[0] > 0:
            mine = full[bounds[self.rank]:bounds[self.rank + 1]]
            self._dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1), group=self.group)
            return full
        if nccl and max(sizes) > 0:
            m = max(sizes)
            buf = torch.empty((self.world, m) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
            buf[self.rank, :sizes[self.rank]] = full[bounds[self.rank]:bounds[self.rank + 1]]
            self._dist.all_gather_into_tensor(buf.view(-1), buf[self.rank].reshape(-1), group=self.group)
            for r in range(self.world):
                if r != self.rank and sizes[r]:
                    full[bounds[r]:bounds[r + 1]] = buf[r, :sizes[r]]
            return full
        parts = [full[bounds[r]:bounds[r + 1]] for r in range(self.world)]
        if len(set(sizes)) == 1:
            self._dist.all_gather(parts, parts[self.rank].clone(), group=self.group)
        else:
            for r in range(self.world):
                self._dist.broadcast(parts[r], src=self._global(r), group=self.group)
        return full

    def allgather_pieces(self, full, bounds, lo, hi):
        """Equal shards only: every rank filled rows bounds[rank] + [lo, hi) of `full`; afterwards every rank holds those rows
        of every shard (one piece of a pipelined allgather_rows; the pieces of the ranks are not adjacent, hence the list)."""
        if not self.active:
            return full
        outs = [full[bounds[r] + lo:bounds[r] + hi] for r in range(self.world)]
        self._dist.all_gather(outs, outs[self.rank], group=self.group)
        return full

    def _global(self, r):
        return r if self.group is None else self._dist.get_global_rank(self.group, r)

    def barrier(self):
        if self.active:
            self._dist.barrier(group=self.group)


def shard_bounds(n, world, align=1):
    """Contiguous shards of `n` units, each boundary a multiple of `align` (last one ragged)."""
    units = (n + align - 1) // align
    base, rem = divmod(units, world)
    bounds, u = [0], 0
    for r in range(world):
        u += base + (1 if r < rem else 0)
        bounds.append(min(n, u * align))
    return bounds


def cyclic_tiles(n_tiles, rank, world):
    """Tiles rank, rank+world, ...: balances the triangular dedup work (hard part 6)."""
    return list(range(rank, n_tiles, world))
