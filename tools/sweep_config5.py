"""BASELINE config 5: zonal-mean-only sweep L in {25,...,800} on ne120pg2 x 72 levels x 24 steps.
Times the forward projection (K4), the output-grid synthesis and the native-grid synthesis separately and
reports them against the FP64 (37.1 TFLOP/s measured DMMA) and HBM (MEASURED_PEAKS.json) rooflines."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytemdiags_b200 import synthetic as syn
from pytemdiags_b200.engine import Engine

PEAK_TF = 37.1
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    HBM = 6650.0
lat, lon = syn.pg2_grid(120)
N, K, T = lat.shape[0], 72, 24
rows = K * T
lat_out = np.arange(-89.5, 90, 1.0)
dev = torch.device('cuda:0')


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = []
x = None
for L in (25, 50, 100, 200, 400, 800):
    eng = Engine(lat, lat_out, L, device=dev)
    import time
    t0 = time.time(); eng.build_basis(); torch.cuda.synchronize(); tb = time.time() - t0
    if x is None:
        x = eng.synth_fields(0, 0, 0, T, syn.default_plev(K), eng._dev(np.deg2rad(lat)), eng._dev(np.deg2rad(lon)), eng._dev(syn.default_plev(K)))
    Lp = L + 1
    coef = eng.project([x])
    nat = torch.empty((rows, N), dtype=torch.float64, device=dev)
    t_proj = timeit(lambda: eng.project([x]))
    t_out = timeit(lambda: eng.synth_out(coef))
    t_nat = timeit(lambda: eng.synth_native(coef[0], out=nat))
    pts = N * rows
    r = dict(L=L, basis_build_ms=tb * 1e3,
             project_ms=t_proj, project_tflops=2.0 * Lp * pts / t_proj / 1e9, project_hbm_gbs=8.0 * pts / t_proj / 1e6,
             synth_out_ms=t_out,
             synth_native_ms=t_nat, synth_native_tflops=2.0 * Lp * pts / t_nat / 1e9, synth_native_hbm_gbs=8.0 * pts / t_nat / 1e6)
    r['project_frac_fp64'] = r['project_tflops'] / PEAK_TF
    r['project_frac_hbm'] = r['project_hbm_gbs'] / HBM
    r['synth_native_frac_fp64'] = r['synth_native_tflops'] / PEAK_TF
    r['synth_native_frac_hbm'] = r['synth_native_hbm_gbs'] / HBM
    r['zonal_mean_pts_per_s'] = pts / ((t_proj + t_out) * 1e-3)
    r['zonal_mean_native_pts_per_s'] = pts / ((t_proj + t_nat) * 1e-3)
    out.append(r)
    print('L %3d  basis %6.1f ms | project %7.3f ms %5.2f TF (%4.1f%% fp64, %4.1f%% hbm) | synth_out %6.3f ms | synth_native %7.3f ms %5.2f TF (%4.1f%% fp64) %6.0f GB/s written (%4.1f%% hbm)'
          % (L, r['basis_build_ms'], t_proj, r['project_tflops'], 100 * r['project_frac_fp64'], 100 * r['project_frac_hbm'], t_out, t_nat,
             r['synth_native_tflops'], 100 * r['synth_native_frac_fp64'], r['synth_native_hbm_gbs'], 100 * r['synth_native_frac_hbm']), flush=True)
    del eng
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/config5_sweep.json', 'w'), indent=1)
