"""Multi-GPU driver: time-slab sharding, one process per GPU.

Every (level, time) column of the reference's `AA` matrix is an independent right-hand side
(sph_zonal_mean.py:244-251) and every stencil acts inside one time step (tem_util.py:154,192,232), so
the record is split into contiguous time slabs, one per rank, with the basis replicated.  There is
NO collective on the data path; the only exchange is ONE all-gather of the small [n_out][time][lev][lat]
output planes at the end.  Two transports:

  * `torch.distributed` (NCCL on GPUs; gloo for CPU tensors, which is what the CPU tests use);
  * `TemdComm`: NCCL called from inside libtemd.so (`temd_comm_*`, `temd_allgather_outputs`), so that a caller of
    the C ABI without torch.distributed has the same multi-GPU path.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

PUBLIC_OUTPUTS = ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv', 'utendepfd', 'utendvtem', 'utendwtem')
TRACER_PUBLIC = ('etfy', 'etfz', 'etdiv', 'qtendetfd', 'qtendvtem', 'qtendwtem')


def shard_bounds(T, world, weights=None):
    """Contiguous time slabs: returns [(t0, t1)] * world.  Without weights the slabs are balanced (the first T % world
    ranks get one extra step).  `weights` (one positive number per rank, identical on every rank) makes rank r's slab
    proportional to weights[r] - e.g. the host->device bandwidth each GPU really gets (`h2d_weights`): on boxes whose
    GPUs do not share the host links evenly an equal split makes everybody wait for the slowest link."""
    T, world = int(T), int(world)
    if weights is None:
        base, extra = divmod(T, world)
        out, t = [], 0
        for r in range(world):
            n = base + (1 if r < extra else 0)
            out.append((t, t + n))
            t += n
        return out
    w = np.asarray(weights, dtype=np.float64)
    if w.shape != (world,) or not np.all(w > 0):
        raise ValueError('weights must be %d positive numbers' % world)
    # minimax apportionment: floor of the proportional share, then the remaining steps one at a time to the rank
    # whose finishing time (steps + 1) / weight stays smallest - minimises max_r steps_r / weights_r
    n = np.floor(T * w / w.sum()).astype(np.int64)
    for _ in range(T - int(n.sum())):
        n[int(np.argmin((n + 1) / w))] += 1
    edges = np.concatenate([[0], np.cumsum(n)])
    return [(int(edges[r]), int(edges[r + 1])) for r in range(world)]


def h2d_weights(device, group=None, nbytes=256 << 20, reps=3):
    """Host->device bandwidth (GB/s) every rank gets while ALL ranks copy at once, as a list identical on every rank:
    the `weights` for `shard_bounds` / `ShardedTEM` when the inputs live in host memory.  (8xB200 box of this pool:
    55.6 GB/s per GPU alone, 23.3 / 35.4 GB/s for GPUs 0-3 / 4-7 when all eight copy.)"""
    device = torch.device(device)
    n = int(nbytes) // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    h.zero_()
    d = torch.empty(n, dtype=torch.float64, device=device)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(device)
    dist.barrier(group)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(device):
        a.record()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        b.record()
        torch.cuda.synchronize(device)
    rate = torch.tensor([reps * n * 8 / (a.elapsed_time(b) * 1e-3) / 1e9], dtype=torch.float64, device=device)
    out = torch.empty(dist.get_world_size(group), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, rate, group=group)
    return [float(x) for x in out.cpu()]


def gather_time_major(local, T, group=None, comm=None, weights=None):
    """ONE all-gather of per-rank blocks shaped (P, T_local, ...) (time second: the device layout
    [plane][time][lev][lat]) into the full (P, T, ...) array on every rank.  Uneven slabs (and empty ones,
    T_local = 0) are zero-padded to the largest slab for the equal-count collective and trimmed afterwards.
    `comm`: a `TemdComm` to run the collective inside libtemd instead of torch.distributed."""
    world = comm.nranks if comm is not None else dist.get_world_size(group)
    rank = comm.rank if comm is not None else dist.get_rank(group)
    bounds = shard_bounds(T, world, weights)
    tmax = max(b - a for a, b in bounds)
    a, b = bounds[rank]
    assert local.shape[1] == b - a, (tuple(local.shape), bounds[rank])
    P, rest = local.shape[0], tuple(local.shape[2:])
    if b - a == tmax and local.is_contiguous():
        mine = local
    else:
        mine = torch.zeros((P, tmax) + rest, dtype=local.dtype, device=local.device)
        mine[:, :b - a] = local
    full = torch.empty((world, P, tmax) + rest, dtype=local.dtype, device=local.device)
    if comm is not None:
        comm.allgather(mine, full)
    else:
        dist.all_gather_into_tensor(full.view((world * P, tmax) + rest), mine, group=group)
    if all(bb - aa == tmax for aa, bb in bounds):
        return full.transpose(0, 1).reshape((P, world * tmax) + rest)
    return torch.cat([full[r, :, :bb - aa] for r, (aa, bb) in enumerate(bounds) if bb > aa], 1)


def gather_time_sharded(local, T, group=None):
    """All-gather per-rank results shaped (..., T_local) (time last, the reference's output layout
    (lat, plev, time)) into the full (..., T) array on every rank."""
    lead = tuple(local.shape[:-1])
    blk = local.movedim(-1, 0).unsqueeze(0)                 # (1, T_local, ...)
    full = gather_time_major(blk.contiguous(), T, group)    # (1, T, ...)
    return full[0].movedim(0, -1).reshape(lead + (T,)).contiguous()


class TemdComm:
    """NCCL communicator owned by libtemd (`temd_comm_init`), for callers of the C ABI.  The 128-byte unique id made
    on rank 0 (`TemdComm.unique_id()`) must reach every rank by any side channel (here: a torch.distributed /
    file / socket broadcast by the caller)."""

    def __init__(self, nranks, rank, unique_id, device):
        from . import _lib
        self.lib = _lib.load()
        self.nranks, self.rank = int(nranks), int(rank)
        self.device = torch.device(device)
        self._comm = C.c_void_p(0)
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _lib.check(self.lib.temd_comm_init(self.device.index or 0, self.nranks, self.rank, buf, C.byref(self._comm)),
                   'temd_comm_init')

    @staticmethod
    def unique_id():
        from . import _lib
        lib = _lib.load()
        buf = (C.c_char * 128)()
        _lib.check(lib.temd_comm_unique_id(buf), 'temd_comm_unique_id')
        return bytes(buf.raw)

    def allgather(self, send, recv):
        from . import _lib
        assert send.is_contiguous() and recv.is_contiguous() and send.dtype == torch.float64
        assert recv.numel() == self.nranks * send.numel()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.temd_allgather_outputs(self._comm, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()),
                                                   send.numel(), stream), 'temd_allgather_outputs')

    def close(self):
        if self._comm.value:
            self.lib.temd_comm_destroy(self._comm)
            self._comm = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedTEM:
    """TEMDiagnostics over a time-sharded record: each rank passes its LOCAL time slab of the inputs (possibly empty)
    and gets full-record outputs back.

        tem = ShardedTEM(ua, va, ta, wap, p, lat, T=T_total, L=..., dims=('time', 'lev', 'ncol'))
        out = tem.gather_all()            # {'vtem': (lat, plev, T_total), ...}: ONE collective for all outputs
        vtem = tem.gather('vtem')         # a single output

    `comm=TemdComm(...)` routes the collective through libtemd's own NCCL call instead of torch.distributed.
    `weights=` (identical on every rank, e.g. `h2d_weights(device)`) sizes the slabs as `shard_bounds(T, world, weights)`.
    """

    def __init__(self, ua, va, ta, wap, *args, T=None, group=None, comm=None, weights=None, **kw):
        from .tem import TEMDiagnostics
        from . import arrays as ar
        self.group, self.comm, self.weights = group, comm, weights
        # local number of time steps, before TEMDiagnostics sees the arrays (an empty slab builds nothing)
        r = ar.raw(ua)
        tname = (kw.get('dim_names') or {}).get('time', 'time')
        if ar.is_dataarray(ua):
            dims = tuple(ua.dims)
        else:
            dims = kw.get('dims') or ('ncol', 'plev', 'time')[:r.ndim]
        tdim = [i for i, d in enumerate(dims) if d in ('time', tname)]
        nt_local = int(r.shape[tdim[0]]) if tdim else 1
        import time as _time
        t_start = _time.perf_counter()
        self.local = TEMDiagnostics(ua, va, ta, wap, *args, **kw) if nt_local > 0 else None
        # wall time of this rank's own work (upload + kernels; the constructor ends with a synchronising NaN screen):
        # steps / local_seconds is the rate to feed back into `weights` when slabs should finish together
        self.local_seconds = _time.perf_counter() - t_start
        self.local_steps = nt_local
        dev = kw.get('device')
        if self.local is not None:
            dev = self.local.ZM._engine.device
        elif dev is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(dev)
        # shapes an empty rank cannot know, and the total record length
        meta = self._allgather_meta([nt_local, self.local.NLEV if self.local else 0,
                                     self.local.ZM_N if self.local else 0, self.local.ntrac if self.local else 0])
        self.T = int(T) if T is not None else int(meta[:, 0].sum())
        self.K, self.M, self.ntrac = (int(meta[:, j].max()) for j in (1, 2, 3))
        a, b = shard_bounds(self.T, self._world(), self.weights)[self._rank()]
        if b - a != nt_local:
            raise RuntimeError('rank {} holds {} time steps but shard_bounds({}, {}{}) assigns it [{}, {})'.format(
                self._rank(), nt_local, self.T, self._world(), '' if self.weights is None else ', weights', a, b))

    def _allgather_meta(self, vals):
        """[world][len(vals)] integer table, through whichever transport this object uses."""
        world = self._world()
        if self.comm is not None:
            send = torch.tensor(vals, dtype=torch.float64, device=self.device)
            recv = torch.empty((world, len(vals)), dtype=torch.float64, device=self.device)
            self.comm.allgather(send, recv)
            return recv.cpu().numpy().astype(np.int64)
        try:
            backend = dist.get_backend(self.group)
        except Exception:
            backend = 'nccl'
        mdev = self.device if backend == 'nccl' else torch.device('cpu')     # NCCL reduces CUDA tensors, gloo CPU ones
        send = torch.tensor(vals, dtype=torch.int64, device=mdev)
        recv = torch.empty((world, len(vals)), dtype=torch.int64, device=mdev)
        dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        return recv.cpu().numpy()

    def _world(self):
        return self.comm.nranks if self.comm is not None else dist.get_world_size(self.group)

    def _rank(self):
        return self.comm.rank if self.comm is not None else dist.get_rank(self.group)

    def _stack(self, names, tracer=None):
        """[len(names)][T_local][K][M] contiguous device tensor of this rank's planes."""
        if self.local is None:
            return torch.zeros((len(names), 0, self.K, self.M), dtype=torch.float64, device=self.device)
        src = self.local._dev_results if tracer is None else self.local._dev_tracer[tracer]
        return torch.stack([src[n] for n in names])

    def gather_all(self, names=PUBLIC_OUTPUTS, tracers=True, layout='reference'):
        """One all-gather for every requested output (plus the six tracer diagnostics of every tracer).
        layout='reference': (lat, plev, time) views like the reference's methods; 'device': [time][lev][lat];
        'stacked': the tuple (tensor [P][time][lev][lat], [(name, tracer index or None)] * P)."""
        blocks, keys = [self._stack(names)], [(n, None) for n in names]
        if tracers:
            for i in range(self.ntrac):
                blocks.append(self._stack(TRACER_PUBLIC, tracer=i))
                keys += [(n, i) for n in TRACER_PUBLIC]
        local = torch.cat(blocks, 0) if len(blocks) > 1 else blocks[0]
        full = gather_time_major(local, self.T, self.group, self.comm, self.weights)         # [P][T][K][M]
        if layout == 'stacked':      # the gathered block itself (one device->host copy moves everything) + plane keys
            return full, keys
        out = {}
        for j, (n, i) in enumerate(keys):
            t = full[j] if layout == 'device' else full[j].permute(2, 1, 0)
            if i is None:
                out[n] = t
            else:
                out.setdefault(n, [None] * self.ntrac)[i] = t
        return out

    def gather(self, name):
        return self.gather_all((name,), tracers=False)[name].contiguous()
