"""Multi-GPU driver: time-slab sharding, one process per GPU.

Every (level, time) column of the reference's `AA` matrix is an independent right-hand side
(sph_zonal_mean.py:244-251) and every stencil acts inside one time step (tem_util.py:154,192,232), so
the record is split into contiguous time slabs, one per rank, with the basis replicated.  There is
NO collective on the data path; the only exchange is one all-gather of the small (lat, plev, time)
outputs at the end (NCCL on GPUs; gloo works for CPU tensors and is what the CPU tests use).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(T, world):
    """Contiguous, balanced time slabs: returns [(t0, t1)] * world; the first T % world ranks get one extra step."""
    base, extra = divmod(int(T), int(world))
    out, t = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((t, t + n))
        t += n
    return out


def gather_time_sharded(local, T, group=None):
    """All-gather per-rank results shaped (..., T_local) (time last, the reference's output layout
    (lat, plev, time)) into the full (..., T) array on every rank.  Uneven slabs are zero-padded to the
    largest slab for the equal-count all-gather and trimmed afterwards."""
    world = dist.get_world_size(group)
    bounds = shard_bounds(T, world)
    tmax = max(b - a for a, b in bounds)
    rank = dist.get_rank(group)
    a, b = bounds[rank]
    assert local.shape[-1] == b - a, (local.shape, bounds[rank])
    lead = tuple(local.shape[:-1])
    # time-major staging so each rank's block is contiguous in the gathered buffer
    mine = torch.zeros((tmax,) + lead, dtype=local.dtype, device=local.device)
    mine[:b - a] = local.movedim(-1, 0)
    full = torch.empty((world, tmax) + lead, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full.view(world * tmax, *lead), mine, group=group)
    parts = [full[r, :bb - aa] for r, (aa, bb) in enumerate(bounds)]
    return torch.cat(parts, 0).movedim(0, -1).contiguous()


class ShardedTEM:
    """TEMDiagnostics over a time-sharded record: each rank passes the FULL-record inputs' local slab
    (or the full arrays plus `time_axis`) and gets full-record outputs back.

        tem = ShardedTEM(ua, va, ta, wap, p, lat, T=T_total, L=..., dims=('time', 'lev', 'ncol'))
        vtem = tem.gather('vtem')         # (lat, plev, T_total) on every rank
    """

    def __init__(self, ua, va, ta, wap, *args, T=None, group=None, **kw):
        from .tem import TEMDiagnostics
        self.group = group
        self.local = TEMDiagnostics(ua, va, ta, wap, *args, **kw)
        self.T = int(T) if T is not None else None
        if self.T is None:
            n = torch.tensor([self.local.NT], dtype=torch.int64, device=self.local.ZM._engine.device)
            dist.all_reduce(n, group=group)
            self.T = int(n.item())

    def gather(self, name):
        t = self.local._dev_results[name].permute(2, 1, 0).contiguous()     # (M, K, T_local) on the device
        return gather_time_sharded(t, self.T, self.group)
