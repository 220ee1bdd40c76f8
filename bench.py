#!/usr/bin/env python
"""Benchmark of the TEM-suite hot path (BASELINE.json metric: column.level.steps/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config config2]

One "step" = the full TEM suite (project -> fused eddy/flux/project -> output-grid synthesis ->
stencil epilogue) over ONE time slab of the named BASELINE config, with the four input fields
already resident in HBM (`value`), and the same thing through the public API
`TEMDiagnostics(ua, va, ta, wap, p, lat)` with pinned HOST arrays (`e2e`).  With N > 1 (torchrun)
every rank owns its own time slab (weak scaling, no collective on the data path) and the public
outputs are all-gathered with NCCL at the end of every step.

`--impl reference` times the reference's CPU algorithm (the NumPy oracle port, factored form: the
literal N x N operator of sph_zonal_mean.py:251 needs 956 GB at this grid) on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

if '--impl' in sys.argv and 'reference' in sys.argv:
    # the reference arm uses every host core (torchrun exports OMP_NUM_THREADS=1 for its workers)
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'tem_suite_column_level_steps_per_s'
UNIT = 'col*lev*steps/s'
FP64_PEAK_TFLOPS = 37.1   # measured DMMA m8n8k4 issue-rate peak on this pool (profiles/r01_microbench_fp64.log)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.max_mhz = None
        self.reasons = set()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def workload(name):
    from pytemdiags_b200 import synthetic as syn
    cfg = dict(syn.CONFIGS[name])
    L = cfg['L'] if not isinstance(cfg['L'], tuple) else cfg['L'][2]
    return cfg, L


def cpu_suite_sample(cfg, L, tsample, seed=0, repeat=1):
    """Times the oracle port (factored form) on `tsample` time steps of the workload.  Returns
    (points per second, seconds per suite, setup seconds, threads)."""
    import oracle
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    from pytemdiags_b200 import synthetic as syn
    lat, lon = syn.make_grid(cfg['grid'])
    plev = syn.default_plev(cfg['K'])
    t0 = time.time()
    mats = oracle.sph_matrices(lat, oracle.zm_latitudes(1), L, method='pinv')
    setup = time.time() - t0
    f = syn.synth_fields(lat, lon, plev, tsample, seed=seed)
    tr = lambda x: np.ascontiguousarray(x.transpose(2, 1, 0))
    args = [tr(f[n]) for n in ('ua', 'va', 'ta', 'wap')]
    best = None
    for _ in range(repeat):
        t0 = time.time()
        oracle.tem_suite(*args, plev, lat, L=L, literal=False, matrices=mats)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    pts = lat.shape[0] * cfg['K'] * tsample
    return pts / best, best, setup, os.cpu_count()


def run_reference(args):
    """The reference arm: NumPy oracle port of the reference algorithm on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg, L = workload(args.config)
    tsample = args.cpu_tsample
    import oracle
    from pytemdiags_b200 import synthetic as syn
    lat, lon = syn.make_grid(cfg['grid'])
    plev = syn.default_plev(cfg['K'])
    mats = oracle.sph_matrices(lat, oracle.zm_latitudes(1), L, method='pinv')
    f = syn.synth_fields(lat, lon, plev, tsample, seed=0)
    tr = lambda x: np.ascontiguousarray(x.transpose(2, 1, 0))
    a = [tr(f[n]) for n in ('ua', 'va', 'ta', 'wap')]
    for _ in range(args.warmup):
        oracle.tem_suite(*a, plev, lat, L=L, literal=False, matrices=mats)
    t0 = time.time()
    for _ in range(args.steps):
        oracle.tem_suite(*a, plev, lat, L=L, literal=False, matrices=mats)
    dt = (time.time() - t0) / args.steps
    pts = lat.shape[0] * cfg['K'] * tsample
    val = pts / dt
    sample = '%d of %d time steps per step of %s (factored NumPy oracle; literal N x N operator infeasible)' % (
        tsample, cfg['T'], args.config)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s: %s, K=%d, L=%d, %d-step sample' % (args.config, cfg['grid'], cfg['K'], L, tsample)},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pytemdiags_b200 import synthetic as syn
    from pytemdiags_b200 import constants as const
    from pytemdiags_b200.engine import Engine
    from pytemdiags_b200 import TEMDiagnostics

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = None
    if world > 1:
        # bind this rank to the CPUs / NUMA node next to its GPU so that pinned host buffers are local to the
        # GPU's PCIe root (8 ranks copying at once otherwise share one socket's memory bandwidth)
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            numa = 'nvml cpu affinity (%d cpus)' % len(os.sched_getaffinity(0))
        except Exception as e:      # noqa: BLE001
            numa = 'unbound (%s)' % type(e).__name__
        dist.init_process_group('nccl', device_id=dev)

    cfg, L = workload(args.config)
    K = cfg['K']
    lat, lon = syn.make_grid(cfg['grid'])
    N = lat.shape[0]
    plev = syn.default_plev(K)
    lat_zm = np.arange(-89.5, 90.0, 1.0)
    Ts = args.slab_steps
    t_off = rank * Ts                      # every rank owns its own time slab of the record

    eng = Engine(lat, lat_zm, L, device=dev)
    t0 = time.time()
    eng.build_basis()
    torch.cuda.synchronize()
    basis_s = time.time() - t0

    latr, lonr = eng._dev(np.deg2rad(lat)), eng._dev(np.deg2rad(lon))
    plev_d = eng._dev(plev)
    xs = [eng.synth_fields(fi, 0, t_off, Ts, plev, latr, lonr, plev_d) for fi in range(4)]
    p_pa = plev * 100
    lev_scale = eng._dev((const.P0 / p_pa) ** const.k)
    f_zm = 2 * const.Om * np.sin(lat_zm * np.pi / 180)
    coslat = np.cos(lat_zm * np.pi / 180)
    torch.cuda.synchronize()

    ev = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        ev.setdefault(name, []).append((a, b))
        return r

    gathered = [None]

    def step():
        c4 = timed('project', lambda: eng.project(xs, lev_scale=lev_scale, scale_field=2, nlev=K))
        cf = timed('eddy_flux_project', lambda: eng.eddy_flux_project(xs[0], xs[1], xs[2], xs[3], c4, lev_scale, K))
        coef = torch.cat([c4, cf], 0)
        eng.check_finite(coef, 'fields')
        zm = timed('synth_out', lambda: eng.synth_out(coef)).reshape(7, Ts, K, eng.M)
        res = timed('epilogue', lambda: eng.tem_epilogue(zm, p_pa, f_zm, coslat))
        if world > 1:
            pub = torch.stack([res[n] for n in ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv',
                                                'utendepfd', 'utendvtem', 'utendwtem')]).contiguous()
            out = torch.empty((world,) + tuple(pub.shape), dtype=pub.dtype, device=dev)
            dist.all_gather_into_tensor(out, pub)
            gathered[0] = out
        return res
    LAUNCHES_PER_STEP = 2 + 2 + 1 + 1 + 4   # project+reduce, eddy+reduce, check_finite, synth_out, 4 epilogue passes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ev.clear()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    barrier()
    clocks = sampler.result()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pts_step = N * K * Ts * world
    value = pts_step / (ms * 1e-3)

    kms = {n: float(np.mean([a.elapsed_time(b) for a, b in v])) for n, v in ev.items()}

    # ---- end-to-end through the public API with pinned host arrays (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        Te = args.e2e_steps
        host = []
        for fi in range(4):
            h = torch.empty((Te, K, N), dtype=torch.float64).pin_memory()
            h.copy_(xs[fi][:Te * K].reshape(Te, K, N))
            host.append(h.numpy())
        torch.cuda.synchronize()
        h2d = 4 * Te * K * N * 8
        d2h = 10 * eng.M * K * Te * 8
        names = ('vtem', 'omegatem', 'wtem', 'psitem', 'epfy', 'epfz', 'epdiv', 'utendepfd', 'utendvtem', 'utendwtem')

        def e2e_step():
            tem = TEMDiagnostics(host[0], host[1], host[2], host[3], plev, lat, L=L, dims=('time', 'lev', 'ncol'),
                                 debug_level=0, device=dev)
            return [getattr(tem, n)() for n in names]
        # raw pinned H2D bandwidth of this box, for context (the e2e figure is PCIe-bound)
        tmp = torch.empty_like(xs[0][:Te * K].reshape(Te, K, N))
        hsrc = torch.from_numpy(host[0])
        torch.cuda.synchronize()
        t0 = time.time()
        tmp.copy_(hsrc, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = hsrc.numel() * 8 / (time.time() - t0) / 1e9
        del tmp
        for _ in range(2):
            e2e_step()
        barrier()
        nrep = 5
        calls = []
        t0 = time.time()
        for _ in range(nrep):
            tc = time.time()
            outs = e2e_step()
            torch.cuda.synchronize()
            calls.append(time.time() - tc)
        barrier()
        dt = (time.time() - t0) / nrep
        if os.environ.get('TEMD_BENCH_DEBUG'):
            print('e2e per-call ms:', [round(c * 1e3, 1) for c in calls], file=sys.stderr)
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {'value': N * K * Te * world / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'time_steps_per_call': Te, 'ms_per_call': dt * 1e3, 'pinned_h2d_gbs_measured': h2d_gbs, 'cpu_binding': numa,
               'note': 'public API TEMDiagnostics(ua, va, ta, wap, p, lat) on pinned host arrays; H2D of slab i+1 overlaps '
                       'compute of slab i; basis cached across calls like the reference\'s maps/ cache'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    Lp = L + 1
    rows = Ts * K
    fl_eddy = 14.0 * Lp * N * rows
    fl_proj = 8.0 * Lp * N * rows
    pk = peaks()
    roof = {'bound': 'tensor', 'kernel': 'k_eddy (temd_eddy_flux_project)',
            'achieved': fl_eddy / (kms['eddy_flux_project'] * 1e-3) / 1e12, 'peak': FP64_PEAK_TFLOPS, 'unit': 'TFLOP/s',
            'traffic': None,
            'peak_source': 'FP64 is not in MEASURED_PEAKS.json; 37.1 TFLOP/s = DMMA.8x8x4 issue-rate peak measured on '
                           'this pool (profiles/r01_microbench_fp64.log), nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2',
            'algorithmic_flops_per_launch': fl_eddy}
    roof['frac'] = roof['achieved'] / roof['peak']
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get('%s:%d' % (args.config, Ts))
        if tr:
            roof['traffic'] = tr['k_eddy']['dram_bytes']
            roof['traffic_unit'] = 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json)'
            roof['algorithmic_bytes_per_launch'] = 32.0 * N * rows
    except Exception:
        pass
    roof_proj = {'bound': 'tensor', 'kernel': 'k_project (temd_project)',
                 'achieved': fl_proj / (kms['project'] * 1e-3) / 1e12, 'peak': FP64_PEAK_TFLOPS, 'unit': 'TFLOP/s'}
    roof_proj['frac'] = roof_proj['achieved'] / roof_proj['peak']
    step_tf = 22.0 * Lp * N * rows / (ms * 1e-3) / 1e12 / world * world   # per-GPU == aggregate/world
    hbm = {'bytes_per_point': 64, 'achieved_gbs': 64.0 * N * rows / (ms * 1e-3) / 1e9, 'peak_gbs': pk.get('hbm_gbs')}

    cpu = None
    if not args.no_cpu and world == 1:
        v, secs, setup, cores = cpu_suite_sample(cfg, L, args.cpu_tsample)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '%d of %d time steps of %s, factored NumPy oracle (literal N x N form infeasible: %.0f GB); '
                         'suite %.1f s, matrix setup (pinv) %.1f s not included'
                         % (args.cpu_tsample, cfg['T'], args.config, 8.0 * N * N / 1e9, secs, setup)}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': '%s: grid %s (%d cols) x %d lev, L=%d, time slab of %d steps per GPU per step '
                               '(%d slabs = the %d-step record)' % (args.config, cfg['grid'], N, K, L, Ts,
                                                                   -(-cfg['T'] // Ts), cfg['T']),
                   'slab_steps': Ts, 'l2': 'inputs (%.1f GB per GPU) are far larger than the 126 MB L2'
                                           % (4 * 8.0 * N * rows / 1e9),
                   'parallelism': 'time slabs, one per GPU' if world > 1 else 'single GPU'},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': LAUNCHES_PER_STEP * args.steps,
        'roofline': roof, 'roofline_project': roof_proj,
        'step_fp64_tflops_per_gpu': step_tf, 'hbm_algorithmic': hbm,
        'kernel_ms': kms, 'basis_build_s': basis_s, 'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='config2')
    ap.add_argument('--slab-steps', type=int, default=73)
    ap.add_argument('--e2e-steps', type=int, default=16)
    ap.add_argument('--cpu-tsample', type=int, default=8)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
