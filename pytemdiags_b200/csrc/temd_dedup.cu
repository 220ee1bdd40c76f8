// Structure-exploiting fast path: columns that share a latitude share a row of the basis
// (reference PyTEMDiags/sph_zonal_mean.py:361-363: Y0[i, l] depends on lat_i only; a raveled lat-lon grid,
// tem_util.py:331, has NLAT distinct rows for NLAT*NLON columns, a pg2 cubed sphere ~N/8).  With the columns grouped
// by unique x = sin(lat) (group u, multiplicity n_u):
//
//   Q^T A          = sum_u Qw[u]^T (S_u / sqrt(n_u)),      S_u = sum_{i in u} A_i,   Qw[u] = sqrt(n_u) Q(x_u)
//   (Q c)_i        = Qw[u(i)] c / sqrt(n_u)                 the native zonal mean is constant inside a group
//   sum_{i in u} (a_i - abar_u)(b_i - bbar_u)               only group sums of a, b and a*b are needed
//
// so the four projections, four native syntheses and three eddy-flux projections of the TEM suite
// (tem_diagnostics.py:510-558) collapse to ONE pass over the four input fields (32 B per column.level.step, HBM
// bound) followed by the usual tensor-core kernels on the U unique columns.  Qw is the orthonormalised basis of the
// weighted unique grid (temd_basis_build_dedup): its Gram matrix equals the dense one, so the coefficients are those
// of the dense path up to rounding.
//
// Numerics: the flux sums are accumulated about a per-group shift (the group's first element a0), i.e.
//   s_f = sum (x_f - a0_f),   p_fg = sum (x_f - a0_f)(x_g - a0_g),
// and combined with the spectral means m_f as  p_fg - d_f s_g - d_g s_f + n d_f d_g,  d_f = m_f - a0_f:  every term is
// eddy-sized, so there is no cancellation against the (much larger) squared mean.  Summation order is fixed (lane-strided
// partial sums + a shuffle tree, or one thread per group): results are run-to-run reproducible.
#include "../../include/temd.h"
#include "temd_common.cuh"
#include "temd_internal.h"

namespace temd {

// plane order of temd_group_sums(with_products = 1); every plane is [rows][ld_out]
enum { GS_A0 = 0, GS_S = 4, GS_P = 8, GS_SW = 11, GS_NPLANES = 15 };

struct GsumParams {
    const double* x[4];
    int nfields, rows, nlev, scale_field, ngroups, with_products;
    size_t ld, ld_out, plane;
    const int* perm;      // sorted position -> column (null: identity, groups are contiguous column ranges)
    const int* goff;      // [ngroups + 1] offsets into the sorted positions
    const double* rsq;    // [ngroups] 1 / sqrt(multiplicity)
    const double* lev_scale;
    double* out;
};

template <bool PROD>
__device__ __forceinline__ void gs_store(const GsumParams& p, int row, int g, int cnt, const double (&a0)[4],
                                         const double (&s)[4], const double (&pr)[3]) {
    const size_t o = (size_t)row * p.ld_out + g;
    const double rs = p.rsq[g];
    if (PROD) {
#pragma unroll
        for (int f = 0; f < 4; f++) {
            p.out[(GS_A0 + f) * p.plane + o] = a0[f];
            p.out[(GS_S + f) * p.plane + o] = s[f];
            p.out[(GS_SW + f) * p.plane + o] = ((double)cnt * a0[f] + s[f]) * rs;
        }
#pragma unroll
        for (int q = 0; q < 3; q++) p.out[(GS_P + q) * p.plane + o] = pr[q];
    } else {
#pragma unroll
        for (int f = 0; f < 4; f++)      // static indices: a run-time loop bound would push a0 / s to local memory
            if (f < p.nfields) p.out[f * p.plane + o] = ((double)cnt * a0[f] + s[f]) * rs;
    }
}

// groups with >= 32 members: one warp per (row, group), lane-strided, fixed-order shuffle tree.
//   MODE 0: members listed in perm (scattered columns);  MODE 1: the group is a contiguous column range (raveled
//   lat-lon grids), scalar loads;  MODE 2: contiguous, every group starts at an even column and has an even count:
//   16-byte loads, each lane owns element pairs (2 lane + 64 i, +1).
// The main loop issues the loads of a whole batch before consuming them (ncu r02_prof_gsum: with each element's loads
// issued right before their use the kernel sat on 4 outstanding 8-byte loads per lane and reached 4.56 TB/s).
// Per-lane accumulation is in element order, so the result depends on MODE (which is a function of the grid) only.
template <bool PROD, int MODE>
__global__ void __launch_bounds__(256) k_gsum_warp(const GsumParams p) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int g = (int)(wid % p.ngroups);
    const int row = (int)(wid / p.ngroups);
    if (row >= p.rows) return;
    const int j0 = p.goff[g], cnt = p.goff[g + 1] - j0;
    if (cnt < 32) return;
    const int nf = PROD ? 4 : p.nfields;
    const double lsc = (p.lev_scale != nullptr && p.scale_field >= 0) ? p.lev_scale[row % p.nlev] : 1.0;
    double sc[4];
#pragma unroll
    for (int f = 0; f < 4; f++) sc[f] = (f == p.scale_field) ? lsc : 1.0;
    auto col = [&](int j) { return MODE == 0 ? p.perm[j0 + j] : j0 + j; };
    const int c0 = col(0);
    double a0[4] = {0.0, 0.0, 0.0, 0.0}, s[4] = {0.0, 0.0, 0.0, 0.0}, pr[3] = {0.0, 0.0, 0.0};
    const double* xr[4];
    xr[0] = p.x[0] + (size_t)row * p.ld;
    xr[1] = (nf > 1 ? p.x[1] : p.x[0]) + (size_t)row * p.ld;
    xr[2] = (nf > 2 ? p.x[2] : p.x[0]) + (size_t)row * p.ld;
    xr[3] = (nf > 3 ? p.x[3] : p.x[0]) + (size_t)row * p.ld;
#pragma unroll
    for (int f = 0; f < 4; f++) if (f < nf) a0[f] = sc[f] * xr[f][c0];
    auto consume = [&](const double (&v)[4]) {
        double d[4];
#pragma unroll
        for (int f = 0; f < 4; f++) d[f] = (f < nf) ? sc[f] * v[f] - a0[f] : 0.0;
#pragma unroll
        for (int f = 0; f < 4; f++) s[f] += d[f];
        if (PROD) {
            pr[0] = fma(d[0], d[1], pr[0]);
            pr[1] = fma(d[0], d[3], pr[1]);
            pr[2] = fma(d[1], d[2], pr[2]);
        }
    };
    if constexpr (MODE == 2) {
        constexpr int UNR = 2;
        const int npair = cnt >> 1;
        int jp = lane;
        for (; jp + 32 * (UNR - 1) < npair; jp += 32 * UNR) {
            double2 v[UNR][4];
#pragma unroll
            for (int u = 0; u < UNR; u++)
#pragma unroll
                for (int f = 0; f < 4; f++)
                    v[u][f] = (f < nf) ? *reinterpret_cast<const double2*>(xr[f] + j0 + 2 * (jp + 32 * u)) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const double lo[4] = {v[u][0].x, v[u][1].x, v[u][2].x, v[u][3].x};
                const double hi[4] = {v[u][0].y, v[u][1].y, v[u][2].y, v[u][3].y};
                consume(lo);
                consume(hi);
            }
        }
        for (; jp < npair; jp += 32) {
            double lo[4], hi[4];
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const double2 t2 = (f < nf) ? *reinterpret_cast<const double2*>(xr[f] + j0 + 2 * jp) : make_double2(0.0, 0.0);
                lo[f] = t2.x; hi[f] = t2.y;
            }
            consume(lo);
            consume(hi);
        }
    } else {
        constexpr int UNR = 4;
        int j = lane;
        for (; j + 32 * (UNR - 1) < cnt; j += 32 * UNR) {
            double v[UNR][4];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int c = col(j + 32 * u);
#pragma unroll
                for (int f = 0; f < 4; f++) v[u][f] = (f < nf) ? xr[f][c] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UNR; u++) consume(v[u]);
        }
        for (; j < cnt; j += 32) {
            const int c = col(j);
            double v[4];
#pragma unroll
            for (int f = 0; f < 4; f++) v[f] = (f < nf) ? xr[f][c] : 0.0;
            consume(v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int f = 0; f < 4; f++) s[f] += __shfl_down_sync(0xffffffffu, s[f], o);
        if (PROD) {
#pragma unroll
            for (int q = 0; q < 3; q++) pr[q] += __shfl_down_sync(0xffffffffu, pr[q], o);
        }
    }
    if (lane == 0) gs_store<PROD>(p, row, g, cnt, a0, s, pr);
}

// groups with < 32 members: one thread per (row, group)
template <bool PROD>
__global__ void __launch_bounds__(256) k_gsum_thread(const GsumParams p) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = (int)(tid % p.ngroups);
    const int row = (int)(tid / p.ngroups);
    if (row >= p.rows) return;
    const int j0 = p.goff[g], cnt = p.goff[g + 1] - j0;
    if (cnt >= 32 || cnt < 1) return;
    const int nf = PROD ? 4 : p.nfields;
    const double lsc = (p.lev_scale != nullptr && p.scale_field >= 0) ? p.lev_scale[row % p.nlev] : 1.0;
    double sc[4];
#pragma unroll
    for (int f = 0; f < 4; f++) sc[f] = (f == p.scale_field) ? lsc : 1.0;
    double a0[4] = {0.0, 0.0, 0.0, 0.0}, s[4] = {0.0, 0.0, 0.0, 0.0}, pr[3] = {0.0, 0.0, 0.0};
    const double* xr[4];
    const int c0 = p.perm ? p.perm[j0] : j0;
    xr[0] = p.x[0] + (size_t)row * p.ld;
    xr[1] = (nf > 1 ? p.x[1] : p.x[0]) + (size_t)row * p.ld;
    xr[2] = (nf > 2 ? p.x[2] : p.x[0]) + (size_t)row * p.ld;
    xr[3] = (nf > 3 ? p.x[3] : p.x[0]) + (size_t)row * p.ld;
#pragma unroll
    for (int f = 0; f < 4; f++) if (f < nf) a0[f] = sc[f] * xr[f][c0];
    for (int j = 1; j < cnt; j++) {
        const int c = p.perm ? p.perm[j0 + j] : j0 + j;
        double d[4];
#pragma unroll
        for (int f = 0; f < 4; f++) d[f] = (f < nf) ? sc[f] * xr[f][c] - a0[f] : 0.0;
#pragma unroll
        for (int f = 0; f < 4; f++) s[f] += d[f];
        if (PROD) {
            pr[0] = fma(d[0], d[1], pr[0]);
            pr[1] = fma(d[0], d[3], pr[1]);
            pr[2] = fma(d[1], d[2], pr[2]);
        }
    }
    gs_store<PROD>(p, row, g, cnt, a0, s, pr);
}

// flux sums about the spectral means:  F_q[row][u] = (p_q - d_a s_b - d_b s_a + n d_a d_b) / sqrt(n),
// d_f = rsq * mw_f - a0_f  (mw = Qw c = sqrt(n) * native zonal mean);  products (u,v), (u,w), (v,theta)
__global__ void k_dedup_flux(const double* __restrict__ gs, size_t plane, size_t ld_gs, const double* __restrict__ mw,
                             size_t plane_m, size_t ld_m, const int* __restrict__ goff, const double* __restrict__ rsq,
                             int rows, int ngroups, double* __restrict__ out, size_t plane_o, size_t ld_o) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = (int)(tid % ngroups);
    const int row = (int)(tid / ngroups);
    if (row >= rows) return;
    const double rs = rsq[g];
    const double n = (double)(goff[g + 1] - goff[g]);
    const size_t o = (size_t)row * ld_gs + g, om = (size_t)row * ld_m + g;
    double d[4], s[4];
#pragma unroll
    for (int f = 0; f < 4; f++) {
        d[f] = rs * mw[f * plane_m + om] - gs[(GS_A0 + f) * plane + o];
        s[f] = gs[(GS_S + f) * plane + o];
    }
    const int fa[3] = {0, 0, 1}, fb[3] = {1, 3, 2};
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const int a = fa[q], b = fb[q];
        const double v = gs[(GS_P + q) * plane + o] - d[a] * s[b] - d[b] * s[a] + n * d[a] * d[b];
        out[q * plane_o + (size_t)row * ld_o + g] = v * rs;
    }
}

// out[row][i] = alpha * scale[row % nlev] * x[row][i] + beta * rsq[gid[i]] * mw[row][gid[i]]
// (x = null: the expanded native zonal mean; alpha = 1, beta = -1: the eddy field)
__global__ void k_dedup_expand(const double* __restrict__ x, size_t ld_x, const double* __restrict__ scale, int nlev,
                               const double* __restrict__ mw, size_t ld_m, const int* __restrict__ gid,
                               const double* __restrict__ rsq, double alpha, double beta, double* __restrict__ out,
                               size_t ld_out, int rows, int n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n) return;
    const int g = gid[i];
    const double rs = rsq ? rsq[g] : 1.0;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        double v = beta * rs * mw[(size_t)r * ld_m + g];
        if (x != nullptr) v += alpha * (scale ? scale[r % nlev] : 1.0) * x[(size_t)r * ld_x + i];
        out[(size_t)r * ld_out + i] = v;
    }
}

int launch_group_sums(const double* const* x, int nfields, int rows, size_t ld, const int* perm, const int* goff,
                      int ngroups, int max_count, int min_count, int even_groups, const double* rsq, const double* lev_scale,
                      int scale_field, int nlev, int with_products, double* out, size_t ld_out, cudaStream_t stream) {
    GsumParams p;
    for (int f = 0; f < 4; f++) p.x[f] = x[f < nfields ? f : 0];
    p.nfields = nfields; p.rows = rows; p.nlev = nlev < 1 ? 1 : nlev; p.scale_field = scale_field;
    p.ngroups = ngroups; p.with_products = with_products;
    p.ld = ld; p.ld_out = ld_out; p.plane = (size_t)rows * ld_out;
    p.perm = perm; p.goff = goff; p.rsq = rsq; p.lev_scale = lev_scale; p.out = out;
    const long long units = (long long)rows * ngroups;
    if (max_count >= 32) {
        const unsigned blocks = (unsigned)((units * 32 + 255) / 256);
        // mode: 0 scattered, 1 contiguous, 2 contiguous with even offsets / counts and 16-byte-aligned rows
        const bool al = ((ld & 1) == 0);
        bool al_ok = al;
        for (int f = 0; f < nfields; f++) al_ok = al_ok && ((reinterpret_cast<uintptr_t>(x[f]) & 15) == 0);
        const int mode = (perm != nullptr) ? 0 : (even_groups && al_ok ? 2 : 1);
#define GS_LAUNCH(PRODV, MODEV) k_gsum_warp<PRODV, MODEV><<<blocks, 256, 0, stream>>>(p)
        if (with_products) { if (mode == 0) GS_LAUNCH(true, 0); else if (mode == 1) GS_LAUNCH(true, 1); else GS_LAUNCH(true, 2); }
        else { if (mode == 0) GS_LAUNCH(false, 0); else if (mode == 1) GS_LAUNCH(false, 1); else GS_LAUNCH(false, 2); }
#undef GS_LAUNCH
    }
    if (min_count < 32) {
        const unsigned blocks = (unsigned)((units + 255) / 256);
        if (with_products) k_gsum_thread<true><<<blocks, 256, 0, stream>>>(p);
        else k_gsum_thread<false><<<blocks, 256, 0, stream>>>(p);
    }
    return (int)cudaGetLastError();
}

int launch_dedup_flux(const double* gs, size_t ld_gs, const double* mw, size_t ld_m, const int* goff, const double* rsq,
                      int rows, int ngroups, double* out, size_t ld_o, cudaStream_t stream) {
    const long long units = (long long)rows * ngroups;
    k_dedup_flux<<<(unsigned)((units + 255) / 256), 256, 0, stream>>>(gs, (size_t)rows * ld_gs, ld_gs, mw, (size_t)rows * ld_m,
                                                                     ld_m, goff, rsq, rows, ngroups, out, (size_t)rows * ld_o, ld_o);
    return (int)cudaGetLastError();
}

int launch_dedup_expand(const double* x, size_t ld_x, const double* scale, int nlev, const double* mw, size_t ld_m,
                        const int* gid, const double* rsq, double alpha, double beta, double* out, size_t ld_out, int rows,
                        int n, cudaStream_t stream) {
    dim3 grid((n + 255) / 256, rows < 32768 ? rows : 32768);
    k_dedup_expand<<<grid, 256, 0, stream>>>(x, ld_x, scale, nlev < 1 ? 1 : nlev, mw, ld_m, gid, rsq, alpha, beta, out, ld_out,
                                             rows, n);
    return (int)cudaGetLastError();
}

}  // namespace temd
