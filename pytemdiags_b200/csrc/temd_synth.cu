// Generic synthesis GEMM   S[row][n] = sum_l C[row][l] * B[l][n]   (FP64 DMMA, TMA-fed).
//
// Second GEMM of `np.matmul(np.matmul(Y, self.Y0inv), AA)` (reference sph_zonal_mean.py:251) in
// factored form: zonal mean = Q c.  Used for (a) the native-grid mean of `sph_zonal_mean_native`
// (sph_zonal_mean.py:285-290), B = QT [lpad][ncol]; (b) the output-grid mean of `sph_zonal_mean`
// (:291-296), B = QpT [lpad][M]; (c) whitening the basis, C = L^-1, B = QT0.
//
// Classic 128 x 64 x 16 tiling: both operands arrive as 128-B-swizzled TMA boxes (C rows are
// K-major, B rows are "MN-major"); the contraction index inside a 16-wide block is visited in the
// order k(t, s) = {0,3,12,15}[t] ^ 2s which makes BOTH fragment loads bank-conflict-free.
// The contraction is short (lpad / 16 = 7 .. 64 blocks), so a CTA's prologue (barrier set-up, first TMA round trip)
// and epilogue (64 KB of stores) are not negligible: 4 pipeline stages (96 KB) let TWO CTAs share an SM, one
// computing while the other starts or drains.
//
// Fused eddy epilogue (field = blockIdx.x / row tiles): out_f = lev_scale * X_f - C_f B, i.e. the native-grid eddy field
// X' = X - ZM.sph_zonal_mean_native(X) of tem_diagnostics.py:517-529 without a separate element-wise pass.  The X tile
// arrives by TMA through the pipeline stages that the last k-blocks free.
#include "temd_common.cuh"
#include "temd_internal.h"

namespace temd {

constexpr int SY_BM = 128, SY_BN = 64, SY_BK = 16;
constexpr int SY_WARPS = 8;
constexpr int SY_THREADS = (SY_WARPS + 1) * 32;
constexpr int SY_STAGE_BYTES = SY_BM * TILE_ROW_BYTES + (SY_BN / 16) * SY_BK * TILE_ROW_BYTES;  // 16 KB + 8 KB
constexpr int SY_STAGES = 4;

struct SynthMaps {
    CUtensorMap c;      // dims {lpad, rows},  box {16, 128}
    CUtensorMap b;      // dims {ncol, lpad},  box {16, 16}
    CUtensorMap x[4];   // eddy mode only: the field X_f of this row batch, dims {ncol, rows}, box {16, 128}
};

// eddy epilogue of field f (x[f] == nullptr: plain synthesis into out[f])
struct SynthEpi {
    const double* x[4];
    double* out[4];
    size_t ld_x;
    const double* lev_scale;   // applied to field `scale_field` only (theta = lev_scale * T)
    int scale_field, nlev;
    int c_field_rows;          // row offset between consecutive fields in the coefficient map
    int row_base;              // first row of this batch inside a field (coefficient rows, lev_scale index)
};

__global__ void __launch_bounds__(SY_THREADS, 2)
k_synth(const __grid_constant__ SynthMaps maps, int rows, int ncol, int nkb, const SynthEpi epi, size_t ld_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bars[2 * SY_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // grid.x = row tiles x fields (fastest): the CTAs that share a B (basis) column tile run back to back, so the
    // basis is read from HBM once per launch and served from L2 to every row tile and field
    const int row_tiles = (rows + SY_BM - 1) / SY_BM;
    const int row0 = (blockIdx.x % row_tiles) * SY_BM;
    const int col0 = blockIdx.y * SY_BN;
    const int fld = blockIdx.x / row_tiles;
    const int crow0 = fld * epi.c_field_rows + epi.row_base + row0;   // row coordinate in the coefficient map
    const bool eddy = (fld == 0 ? epi.x[0] : fld == 1 ? epi.x[1] : fld == 2 ? epi.x[2] : epi.x[3]) != nullptr;
    const uint32_t smem_base = smem_u32(smem), bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (SY_STAGES + s); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < SY_STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), SY_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == SY_WARPS) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.c);
            tma_prefetch_desc(&maps.b);
            for (int i = 0; i < nkb; i++) {
                const int s = i % SY_STAGES;
                mbar_wait(empty_bar(s), ((i / SY_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(full_bar(s), SY_STAGE_BYTES);
                const uint32_t dst = smem_base + s * SY_STAGE_BYTES;
                tma_load_2d(dst, &maps.c, i * SY_BK, crow0, full_bar(s));
#pragma unroll
                for (int b = 0; b < SY_BN / 16; b++)
                    tma_load_2d(dst + SY_BM * TILE_ROW_BYTES + b * SY_BK * TILE_ROW_BYTES, &maps.b, col0 + b * 16,
                                i * SY_BK, full_bar(s));
            }
            if (eddy) {
                // Eddy mode: the [128 x 64] X tile follows the operands through the ring as four [128 x 16] boxes, one
                // per stage as the last k-blocks release them: its DRAM latency is hidden under the tail of the main
                // loop instead of being paid by every thread in the epilogue (ncu r02_prof_syntheddy: 34 % of the warp
                // samples waiting on those loads, tensor pipe 83 % active against 92 % for the plain synthesis).
                tma_prefetch_desc(&maps.x[fld]);
                for (int j = 0; j < SY_BN / 16; j++) {
                    const int i = nkb + j, s = i % SY_STAGES;
                    mbar_wait(empty_bar(s), ((i / SY_STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(full_bar(s), SY_BM * TILE_ROW_BYTES);
                    tma_load_2d(smem_base + s * SY_STAGE_BYTES, &maps.x[fld], col0 + j * 16, row0, full_bar(s));
                }
            }
        }
        return;
    }

    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int kbase = (t == 0) ? 0 : (t == 1) ? 3 : (t == 2) ? 12 : 15;
    uint32_t a_off[4], b_off[4];   // per k-step byte offsets (row part for B, column part for A)
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int k = kbase ^ (2 * s);
        a_off[s] = (uint32_t)((wm * 32 + g) * TILE_ROW_BYTES) + (uint32_t)((((k >> 1) ^ g) & 7) << 4) + (uint32_t)((k & 1) << 3);
        // B: column inside box = (jn&1)*8 + g ; fold the jn-independent part here
        b_off[s] = (uint32_t)(SY_BM * TILE_ROW_BYTES + k * TILE_ROW_BYTES);
    }

    for (int i = 0; i < nkb; i++) {
        const int s = i % SY_STAGES;
        mbar_wait(full_bar(s), (i / SY_STAGES) & 1);
        const uint32_t st = smem_base + s * SY_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const int k = kbase ^ (2 * ks);
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = lds64(st + a_off[ks] + mi * 8 * TILE_ROW_BYTES);
#pragma unroll
            for (int jn = 0; jn < 4; jn++) {
                const int box = 2 * wn + (jn >> 1);
                const int col = (jn & 1) * 8 + g;
                b[jn] = lds64(st + b_off[ks] + box * (SY_BK * TILE_ROW_BYTES) + ((((col >> 1) ^ k) & 7) << 4) + ((col & 1) << 3));
            }
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int jn = 0; jn < 4; jn++) dmma(acc[mi][jn][0], acc[mi][jn][1], a[mi], b[jn]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
    }

    // explicit selects: indexing a kernel-parameter array with a run-time index would copy it to local memory
    double* __restrict__ out = fld == 0 ? epi.out[0] : fld == 1 ? epi.out[1] : fld == 2 ? epi.out[2] : epi.out[3];
    if (eddy) {
        // this warp's 32 columns are X boxes 2 wn and 2 wn + 1 (stages of "k-blocks" nkb + 2 wn, nkb + 2 wn + 1)
        for (int j = 2 * wn; j < 2 * wn + 2; j++) {
            const int i = nkb + j;
            mbar_wait(full_bar(i % SY_STAGES), (i / SY_STAGES) & 1);
        }
    }
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int rl = wm * 32 + mi * 8 + g;          // row inside the tile
        const int row = row0 + rl;
        if (row >= rows) continue;
        double sc = 1.0;
        if (eddy && epi.lev_scale != nullptr && fld == epi.scale_field) sc = epi.lev_scale[(epi.row_base + row) % epi.nlev];
#pragma unroll
        for (int jn = 0; jn < 4; jn++) {
            const int col = col0 + wn * 32 + jn * 8 + 2 * t;
            double* dst = out + (size_t)row * ld_out + col;
            double v0 = acc[mi][jn][0], v1 = acc[mi][jn][1];
            if (eddy) {
                const int i = nkb + 2 * wn + (jn >> 1);
                const double2 x2 = lds128(smem_base + (i % SY_STAGES) * SY_STAGE_BYTES + swz_off(rl, (jn & 1) * 8 + 2 * t));
                v0 = sc * x2.x - v0;
                v1 = sc * x2.y - v1;
            }
            if (col + 1 < ncol) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            else if (col < ncol) *dst = v0;
        }
    }
}

static int launch_synth_impl(const double* c, int c_rows_total, int rows, int lpad, size_t ld_c, const double* b, int ncol,
                             size_t ld_b, const SynthEpi& epi, int nfields, size_t ld_out, cudaStream_t stream) {
    SynthMaps maps;
    int rc = make_tma_2d(&maps.c, c, (uint64_t)lpad, (uint64_t)c_rows_total, ld_c * sizeof(double), SY_BK, SY_BM);
    if (rc) return rc;
    rc = make_tma_2d(&maps.b, b, (uint64_t)ncol, (uint64_t)lpad, ld_b * sizeof(double), 16, SY_BK);
    if (rc) return rc;
    for (int f = 0; f < nfields; f++)
        if ((ld_out & 1) || (reinterpret_cast<uintptr_t>(epi.out[f]) & 15))
            return temd_set_error(-1, "synth: output must be 16-byte aligned with an even leading dimension");
    for (int f = 0; f < 4; f++) {
        if (f < nfields && epi.x[f] != nullptr) {
            rc = make_tma_2d(&maps.x[f], epi.x[f], (uint64_t)ncol, (uint64_t)rows, epi.ld_x * sizeof(double), 16, SY_BM);
            if (rc) return rc;
        } else {
            maps.x[f] = maps.b;
        }
    }
    constexpr int smem = SY_STAGES * SY_STAGE_BYTES + 1024;
    // per-device attribute: set on every launch (cheap) so that several devices in one process all work
    cudaError_t e = cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return temd_set_error((int)e, "synth: cudaFuncSetAttribute failed");
    const int col_tiles = (ncol + SY_BN - 1) / SY_BN;
    if (col_tiles > 65535) return temd_set_error(-1, "synth: too many column tiles");
    dim3 grid(((rows + SY_BM - 1) / SY_BM) * nfields, col_tiles);
    const int nkb = (lpad + SY_BK - 1) / SY_BK;
    k_synth<<<grid, SY_THREADS, smem, stream>>>(maps, rows, ncol, nkb, epi, ld_out);
    rc = (int)cudaGetLastError();
    if (rc) return temd_set_error(rc, "synth: kernel launch failed");
    return 0;
}

int launch_synth(const double* c, int rows, int lpad, size_t ld_c, const double* b, int ncol, size_t ld_b, double* out,
                 size_t ld_out, cudaStream_t stream) {
    SynthEpi epi = {};
    epi.out[0] = out;
    epi.scale_field = -1;
    epi.nlev = 1;
    return launch_synth_impl(c, rows, rows, lpad, ld_c, b, ncol, ld_b, epi, 1, ld_out, stream);
}

// Eddy fields of a row batch: out[f][r][n] = lev_scale * x[f][r0 + r][n] - (C_f B)[r0 + r][n], r < rows, f < 4
// (coef4 is [4][rows_total][lpad]; x[f] / out[f] point at the batch's first row).
int launch_synth_eddy4(const double* coef4, int rows_total, int r0, int rows, int lpad, const double* b, int ncol, size_t ld_b,
                       const double* const* x, size_t ld_x, const double* lev_scale, int scale_field, int nlev,
                       double* const* out, size_t ld_out, cudaStream_t stream) {
    SynthEpi epi = {};
    for (int f = 0; f < 4; f++) { epi.x[f] = x[f]; epi.out[f] = out[f]; }
    epi.ld_x = ld_x;
    epi.lev_scale = lev_scale;
    epi.scale_field = lev_scale ? scale_field : -1;
    epi.nlev = nlev < 1 ? 1 : nlev;
    epi.c_field_rows = rows_total;
    epi.row_base = r0;
    return launch_synth_impl(coef4, 4 * rows_total, rows, lpad, (size_t)lpad, b, ncol, ld_b, epi, 4, ld_out, stream);
}

}  // namespace temd
