// K6 `tem_epilogue`: every stencil / scan / closed-form diagnostic on the (time, lev, lat) zonal means.
//
// Replaces reference PyTEMDiags/tem_diagnostics.py `_compute_derivatives` (:574-599) and the methods
// vtem (:622), omegatem (:639), wtem (:657), psitem (:674), epfy (:691-692), epfz (:709-710),
// epdiv (:730-736), utendepfd (:753), utendvtem (:771-773), utendwtem (:790-791), and the helpers in
// PyTEMDiags/tem_util.py: multiply_lat (:80), multiply_p (:117), lat_gradient (:154),
// p_gradient (:192), p_integral (:230-232).  Arrays are [time][lev][lat] (lat contiguous, leading
// dimension ld) and tiny (2.5 MB each at config 1).  TWO launches: k_epi_1 (first derivatives, psi and the pressure
// integral) and
// k_epi_2 (everything that needs psi neighbours, and the EP-flux divergence, whose F_phi cos(phi) / F_p neighbours
// are re-evaluated from the planes of k_epi_1 instead of being staged through scratch planes and a third pass).
// The tracer epilogue is ONE launch (k_tr) with the same halo re-evaluation.  HBM/L2-bound.
//
// np.gradient semantics (edge_order=1): one-sided first differences at the two ends; interior is
// a*f[i-1] + b*f[i] + c*f[i+1] with the non-uniform second-order coefficients, or
// (f[i+1]-f[i-1])/(2h) when NumPy detects an exactly uniform coordinate.
#include "../../include/temd.h"
#include "temd_common.cuh"
#include "temd_internal.h"

namespace temd {

struct Axis {
    const double* x;      // coordinate [n]
    const double* coef;   // [3][n] interior coefficients a, b, c
    int n;
    int uniform;
    double h;
};

__device__ __forceinline__ double grad3(const Axis& ax, int i, double fm, double f0, double fp) {
    if (i == 0) return ax.uniform ? (fp - f0) / ax.h : (fp - f0) / (ax.x[1] - ax.x[0]);
    if (i == ax.n - 1) return ax.uniform ? (f0 - fm) / ax.h : (f0 - fm) / (ax.x[ax.n - 1] - ax.x[ax.n - 2]);
    if (ax.uniform) return (fp - fm) / (2.0 * ax.h);
    return ax.coef[i] * fm + ax.coef[ax.n + i] * f0 + ax.coef[2 * ax.n + i] * fp;
}

struct EpiDev {
    int nt, nlev, nlat;
    size_t ld, plane;     // plane = nt*nlev*ld
    const double* zm;
    double* out;
    const double* p;
    const double* coslat;
    const double* f;
    Axis ap, al;
    double p0, a, H, g0, pi;
};

#define ZM(q) (e.zm + (size_t)(q) * e.plane)
#define OUT(q) (e.out + (size_t)(q) * e.plane)
enum { Z_UB = 0, Z_VB, Z_THETAB, Z_WAPB, Z_UPVPB, Z_UPWAPPB, Z_VPTPB };

__device__ __forceinline__ bool epi_index(const EpiDev& e, int& t, int& k, int& m, size_t& idx) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)e.nt * e.nlev * e.nlat;
    if (gid >= total) return false;
    m = (int)(gid % e.nlat);
    k = (int)((gid / e.nlat) % e.nlev);
    t = (int)(gid / ((size_t)e.nlat * e.nlev));
    idx = ((size_t)t * e.nlev + k) * e.ld + m;
    return true;
}

// pass A for one point: dub_dp, dthetab_dp, psi, ubcoslat, dubcoslat_dlat, psicoslat   (tem_diagnostics.py:579-592)
__device__ __forceinline__ void epi_point_a(const EpiDev& e, int k, int m, size_t idx) {
    const size_t up = (k > 0) ? idx - e.ld : idx, dn = (k < e.nlev - 1) ? idx + e.ld : idx;
    const double* ub = ZM(Z_UB);
    const double* th = ZM(Z_THETAB);
    const double dub_dp = grad3(e.ap, k, ub[up], ub[idx], ub[dn]);
    const double dth_dp = grad3(e.ap, k, th[up], th[idx], th[dn]);
    const double psi = ZM(Z_VPTPB)[idx] / dth_dp;                                     // :590
    const double c0 = e.coslat[m];
    const double ubc = ub[idx] * c0;                                                  // :584
    const double ubc_m = (m > 0) ? ub[idx - 1] * e.coslat[m - 1] : ubc;
    const double ubc_p = (m < e.nlat - 1) ? ub[idx + 1] * e.coslat[m + 1] : ubc;
    OUT(TEMD_OUT_DUB_DP)[idx] = dub_dp;
    OUT(TEMD_OUT_DTHETAB_DP)[idx] = dth_dp;
    OUT(TEMD_OUT_PSI)[idx] = psi;
    OUT(TEMD_OUT_UBCOSLAT)[idx] = ubc;
    OUT(TEMD_OUT_DUBCOSLAT_DLAT)[idx] = grad3(e.al, m, ubc_m, ubc, ubc_p);            // :586
    OUT(TEMD_OUT_PSICOSLAT)[idx] = psi * c0;                                          // :592
}

// k_epi_1 = pass A + int_vbdp (tem_util.py:230-232): cumulative trapezoid from the model top,
//   out[k] = sum_{j<=k} (p_j - p_{j-1}) (v_j + v_{j-1}) / 2,  out[0] = 0,
// accumulated sequentially along the level axis like the reference's trapz loop (bit-identical to
// a cumsum).  Thread per (t, k, lat) for pass A; the threads of level 0 additionally walk down their column (adjacent
// threads = adjacent latitudes, so every load / store of the walk is coalesced; the loads are independent and
// pipelined, only the adds are serial).  Measured 3x faster than the shared-memory warp-scan it replaces: the planes
// are L2-resident and a few hundred short columns do not need a parallel scan.
__global__ void __launch_bounds__(256) k_epi_1(const EpiDev e) {
    int t, k, m; size_t idx;
    if (!epi_index(e, t, k, m, idx)) return;
    epi_point_a(e, k, m, idx);
    if (k != 0) return;
    const double* __restrict__ vb = ZM(Z_VB) + idx;
    double* __restrict__ o = OUT(TEMD_OUT_INT_VBDP) + idx;
    double acc = 0.0, vprev = vb[0];
    o[0] = 0.0;
    // batches of 8 levels: all loads of a batch are issued before the serial adds (the first version, one load per
    // iteration behind a store the compiler could not prove independent, paid one L2 latency per level: 54 us)
    constexpr int WB = 8;
    for (int k0 = 1; k0 < e.nlev; k0 += WB) {
        double v[WB];
#pragma unroll
        for (int j = 0; j < WB; j++) v[j] = (k0 + j < e.nlev) ? vb[(size_t)(k0 + j) * e.ld] : 0.0;
#pragma unroll
        for (int j = 0; j < WB; j++) {
            const int kk = k0 + j;
            if (kk < e.nlev) {
                acc += (e.p[kk] - e.p[kk - 1]) * (v[j] + vprev) / 2.0;
                o[(size_t)kk * e.ld] = acc;
                vprev = v[j];
            }
        }
    }
}

// F_phi cos(phi) and F_p of tem_diagnostics.py:730-733 at an arbitrary point, from the planes written by k_epi_1
__device__ __forceinline__ double epi_epfy(const EpiDev& e, size_t idx, int m, double pk) {
    return ((OUT(TEMD_OUT_DUB_DP)[idx] * OUT(TEMD_OUT_PSI)[idx] - ZM(Z_UPVPB)[idx]) * (e.a * e.coslat[m])) * (pk / e.p0);   // :691-692
}
__device__ __forceinline__ double epi_epfz(const EpiDev& e, size_t idx, int m) {
    const double acos = e.a * e.coslat[m];
    const double xz = e.f[m] - OUT(TEMD_OUT_DUBCOSLAT_DLAT)[idx] * (1.0 / acos);      // :709
    return -e.H / e.p0 * ((xz * OUT(TEMD_OUT_PSI)[idx] - ZM(Z_UPWAPPB)[idx]) * acos); // :710
}
__device__ __forceinline__ double epi_fphicos(const EpiDev& e, size_t idx, int m, double pk) {
    return (epi_epfy(e, idx, m, pk) * (e.p0 / pk)) * e.coslat[m];                     // :730, :733
}
__device__ __forceinline__ double epi_fp(const EpiDev& e, size_t idx, int m) {
    return epi_epfz(e, idx, m) * -e.p0 / e.H;                                         // :731
}

// k_epi_2: everything that needs psi / psicoslat neighbours (tem_diagnostics.py:594-597, 615-716, 763-797) and the
// EP-flux divergence (:734-736, 753), whose neighbour values of F_phi cos(phi) (lat +-1) and F_p (lev +-1) are
// re-evaluated from the k_epi_1 planes.
__global__ void k_epi_2(const EpiDev e) {
    int t, k, m; size_t idx;
    if (!epi_index(e, t, k, m, idx)) return;
    const size_t up = (k > 0) ? idx - e.ld : idx, dn = (k < e.nlev - 1) ? idx + e.ld : idx;
    const size_t lm = (m > 0) ? idx - 1 : idx, lp = (m < e.nlat - 1) ? idx + 1 : idx;
    const double* psi_a = OUT(TEMD_OUT_PSI);
    const double* psic_a = OUT(TEMD_OUT_PSICOSLAT);
    const double psi = psi_a[idx];
    const double dpsi_dp = grad3(e.ap, k, psi_a[up], psi, psi_a[dn]);                 // :596
    const double dpsic_dlat = grad3(e.al, m, psic_a[lm], psic_a[idx], psic_a[lp]);    // :594
    const double c0 = e.coslat[m], pk = e.p[k];
    const double iacos = 1.0 / (e.a * c0);
    const double dub_dp = OUT(TEMD_OUT_DUB_DP)[idx];
    const double vtem = ZM(Z_VB)[idx] - dpsi_dp;                                      // :622
    const double omegatem = ZM(Z_WAPB)[idx] + dpsic_dlat * iacos;                     // :639
    const double wtem = omegatem * (-e.H / pk);                                       // :657
    const double psitem = 2 * e.pi * e.a / e.g0 * ((OUT(TEMD_OUT_INT_VBDP)[idx] - psi) * c0);   // :674
    const double epfy = epi_epfy(e, idx, m, pk);
    const double xz = e.f[m] - OUT(TEMD_OUT_DUBCOSLAT_DLAT)[idx] * iacos;             // :709
    const double epfz = epi_epfz(e, idx, m);
    const int mm = (m > 0) ? m - 1 : m, mp = (m < e.nlat - 1) ? m + 1 : m;
    const double epdiv = grad3(e.al, m, epi_fphicos(e, lm, mm, pk), (epfy * (e.p0 / pk)) * c0, epi_fphicos(e, lp, mp, pk)) * iacos
                       + grad3(e.ap, k, epi_fp(e, up, m), epfz * -e.p0 / e.H, epi_fp(e, dn, m));               // :734-736
    OUT(TEMD_OUT_DPSI_DP)[idx] = dpsi_dp;
    OUT(TEMD_OUT_DPSICOSLAT_DLAT)[idx] = dpsic_dlat;
    OUT(TEMD_OUT_VTEM)[idx] = vtem;
    OUT(TEMD_OUT_OMEGATEM)[idx] = omegatem;
    OUT(TEMD_OUT_WTEM)[idx] = wtem;
    OUT(TEMD_OUT_PSITEM)[idx] = psitem;
    OUT(TEMD_OUT_EPFY)[idx] = epfy;
    OUT(TEMD_OUT_EPFZ)[idx] = epfz;
    OUT(TEMD_OUT_UTENDVTEM)[idx] = vtem * xz;                                         // :771-773
    OUT(TEMD_OUT_UTENDWTEM)[idx] = -omegatem * dub_dp;                                // :790-791
    OUT(TEMD_OUT_EPDIV)[idx] = epdiv;
    OUT(TEMD_OUT_UTENDEPFD)[idx] = epdiv * iacos;                                     // :753
}

// ---------------------------------------------------------------------------------------------
// Tracer TEM for one tracer (tem_diagnostics.py:602-611, 801-991)
// ---------------------------------------------------------------------------------------------
struct TrDev {
    int nt, nlev, nlat;
    size_t ld, plane;
    const double *zmq, *psi, *vtem, *omegatem, *p, *coslat;
    double* out;
    Axis ap, al;
    double p0, a, H;
};
#define TOUT(q) (e.out + (size_t)(q) * e.plane)

__device__ __forceinline__ bool tr_index(const TrDev& e, int& t, int& k, int& m, size_t& idx) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)e.nt * e.nlev * e.nlat) return false;
    m = (int)(gid % e.nlat);
    k = (int)((gid / e.nlat) % e.nlev);
    t = (int)(gid / ((size_t)e.nlat * e.nlev));
    idx = ((size_t)t * e.nlev + k) * e.ld + m;
    return true;
}

// d qb / dp (:604) and d (qb cos) / dlat (:608) at an arbitrary point
__device__ __forceinline__ double tr_dqb_dp(const TrDev& e, size_t idx, int k) {
    const double* qb = e.zmq;
    const size_t up = (k > 0) ? idx - e.ld : idx, dn = (k < e.nlev - 1) ? idx + e.ld : idx;
    return grad3(e.ap, k, qb[up], qb[idx], qb[dn]);
}
__device__ __forceinline__ double tr_dqbc_dlat(const TrDev& e, size_t idx, int m) {
    const double* qb = e.zmq;
    const double qbc = qb[idx] * e.coslat[m];
    const double qbc_m = (m > 0) ? qb[idx - 1] * e.coslat[m - 1] : qbc;
    const double qbc_p = (m < e.nlat - 1) ? qb[idx + 1] * e.coslat[m + 1] : qbc;
    return grad3(e.al, m, qbc_m, qbc, qbc_p);
}
__device__ __forceinline__ double tr_etfy(const TrDev& e, size_t idx, int k, int m) {
    return ((tr_dqb_dp(e, idx, k) * e.psi[idx] - e.zmq[e.plane + idx]) * (e.a * e.coslat[m])) * (e.p[k] / e.p0);   // :825-826
}
__device__ __forceinline__ double tr_etfz(const TrDev& e, size_t idx, int m) {
    const double acos = e.a * e.coslat[m];
    const double xq = -(tr_dqbc_dlat(e, idx, m) * (1.0 / acos));                      // :859
    return -e.H / e.p0 * ((xq * e.psi[idx] - e.zmq[2 * e.plane + idx]) * acos);      // :860
}

// ONE launch per tracer: the flux-divergence neighbours (M_phi cos(phi) at lat +-1, M_p at lev +-1, :893-899) are
// re-evaluated from qb / psi instead of being staged through scratch planes.
__global__ void k_tr(const TrDev e) {
    int t, k, m; size_t idx;
    if (!tr_index(e, t, k, m, idx)) return;
    const size_t up = (k > 0) ? idx - e.ld : idx, dn = (k < e.nlev - 1) ? idx + e.ld : idx;
    const size_t lm = (m > 0) ? idx - 1 : idx, lp = (m < e.nlat - 1) ? idx + 1 : idx;
    const int mm = (m > 0) ? m - 1 : m, mp = (m < e.nlat - 1) ? m + 1 : m;
    const int km = (k > 0) ? k - 1 : k, kp = (k < e.nlev - 1) ? k + 1 : k;
    const double c0 = e.coslat[m], pk = e.p[k];
    const double iacos = 1.0 / (e.a * c0);
    const double dqb_dp = tr_dqb_dp(e, idx, k);
    const double dqbc_dlat = tr_dqbc_dlat(e, idx, m);
    const double etfy = tr_etfy(e, idx, k, m);
    const double etfz = tr_etfz(e, idx, m);
    auto mphicos = [&](size_t i, int mi) { return (tr_etfy(e, i, k, mi) * (e.p0 / pk)) * e.coslat[mi]; };   // :893, :896
    auto mpp = [&](size_t i) { return tr_etfz(e, i, m) * -e.p0 / e.H; };                                      // :894
    (void)km; (void)kp;
    const double etdiv = grad3(e.al, m, mphicos(lm, mm), (etfy * (e.p0 / pk)) * c0, mphicos(lp, mp)) * iacos
                       + grad3(e.ap, k, mpp(up), etfz * -e.p0 / e.H, mpp(dn));                               // :897-899
    TOUT(TEMD_TROUT_DQB_DP)[idx] = dqb_dp;
    TOUT(TEMD_TROUT_QBCOSLAT)[idx] = e.zmq[idx] * c0;                                 // :606
    TOUT(TEMD_TROUT_DQBCOSLAT_DLAT)[idx] = dqbc_dlat;
    TOUT(TEMD_TROUT_ETFY)[idx] = etfy;
    TOUT(TEMD_TROUT_ETFZ)[idx] = etfz;
    TOUT(TEMD_TROUT_QTENDVTEM)[idx] = -e.vtem[idx] * (dqbc_dlat * iacos);             // :958-959
    TOUT(TEMD_TROUT_QTENDWTEM)[idx] = -e.omegatem[idx] * dqb_dp;                      // :986-987
    TOUT(TEMD_TROUT_ETDIV)[idx] = etdiv;
    TOUT(TEMD_TROUT_QTENDETFD)[idx] = etdiv * iacos;                                  // :928
}

int launch_tracer_epilogue(const temd_tracer_args& a, cudaStream_t stream) {
    TrDev e;
    e.nt = a.nt; e.nlev = a.nlev; e.nlat = a.nlat; e.ld = a.ld;
    e.plane = (size_t)a.nt * a.nlev * a.ld;
    e.zmq = a.zmq; e.psi = a.psi; e.vtem = a.vtem; e.omegatem = a.omegatem; e.p = a.p; e.coslat = a.coslat; e.out = a.out;
    e.ap = Axis{a.p, a.gp, a.nlev, a.p_uniform, a.hp};
    e.al = Axis{a.latr, a.gl, a.nlat, a.lat_uniform, a.hlat};
    e.p0 = a.p0; e.a = a.a; e.H = a.H;
    const size_t total = (size_t)a.nt * a.nlev * a.nlat;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    k_tr<<<blocks, 256, 0, stream>>>(e);
    return (int)cudaGetLastError();
}

struct EpilogueArgs { temd_epilogue_args a; };

int launch_tem_epilogue(const EpilogueArgs& wrap, cudaStream_t stream) {
    const temd_epilogue_args& a = wrap.a;
    EpiDev e;
    e.nt = a.nt; e.nlev = a.nlev; e.nlat = a.nlat; e.ld = a.ld;
    e.plane = (size_t)a.nt * a.nlev * a.ld;
    e.zm = a.zm; e.out = a.out; e.p = a.p; e.coslat = a.coslat; e.f = a.f;
    e.ap = Axis{a.p, a.gp, a.nlev, a.p_uniform, a.hp};
    e.al = Axis{a.latr, a.gl, a.nlat, a.lat_uniform, a.hlat};
    e.p0 = a.p0; e.a = a.a; e.H = a.H; e.g0 = a.g0; e.pi = a.pi;
    const size_t total = (size_t)a.nt * a.nlev * a.nlat;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    k_epi_1<<<blocks, 256, 0, stream>>>(e);
    k_epi_2<<<blocks, 256, 0, stream>>>(e);
    return (int)cudaGetLastError();
}

}  // namespace temd
