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
#include "temd_common.cuh"
#include "temd_internal.h"

namespace temd {

constexpr int SY_BM = 128, SY_BN = 64, SY_BK = 16;
constexpr int SY_WARPS = 8;
constexpr int SY_THREADS = (SY_WARPS + 1) * 32;
constexpr int SY_STAGE_BYTES = SY_BM * TILE_ROW_BYTES + (SY_BN / 16) * SY_BK * TILE_ROW_BYTES;  // 16 KB + 8 KB
constexpr int SY_STAGES = 8;

struct SynthMaps {
    CUtensorMap c;   // dims {lpad, rows},  box {16, 128}
    CUtensorMap b;   // dims {ncol, lpad},  box {16, 16}
};

__global__ void __launch_bounds__(SY_THREADS, 1)
k_synth(const __grid_constant__ SynthMaps maps, int rows, int ncol, int nkb, double* __restrict__ out, size_t ld_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bars[2 * SY_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * SY_BM;
    const int col0 = blockIdx.y * SY_BN;
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
                tma_load_2d(dst, &maps.c, i * SY_BK, row0, full_bar(s));
#pragma unroll
                for (int b = 0; b < SY_BN / 16; b++)
                    tma_load_2d(dst + SY_BM * TILE_ROW_BYTES + b * SY_BK * TILE_ROW_BYTES, &maps.b, col0 + b * 16,
                                i * SY_BK, full_bar(s));
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

#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int row = row0 + wm * 32 + mi * 8 + g;
        if (row >= rows) continue;
#pragma unroll
        for (int jn = 0; jn < 4; jn++) {
            const int col = col0 + wn * 32 + jn * 8 + 2 * t;
            double* dst = out + (size_t)row * ld_out + col;
            if (col + 1 < ncol) *reinterpret_cast<double2*>(dst) = make_double2(acc[mi][jn][0], acc[mi][jn][1]);
            else if (col < ncol) *dst = acc[mi][jn][0];
        }
    }
}

int launch_synth(const double* c, int rows, int lpad, size_t ld_c, const double* b, int ncol, size_t ld_b, double* out,
                 size_t ld_out, cudaStream_t stream) {
    SynthMaps maps;
    int rc = make_tma_2d(&maps.c, c, (uint64_t)lpad, (uint64_t)rows, ld_c * sizeof(double), SY_BK, SY_BM);
    if (rc) return rc;
    rc = make_tma_2d(&maps.b, b, (uint64_t)ncol, (uint64_t)lpad, ld_b * sizeof(double), 16, SY_BK);
    if (rc) return rc;
    if ((ld_out & 1) || (reinterpret_cast<uintptr_t>(out) & 15)) return temd_set_error(-1, "synth: output must be 16-byte aligned with an even leading dimension");
    constexpr int smem = SY_STAGES * SY_STAGE_BYTES + 1024;
    // per-device attribute: set on every launch (cheap) so that several devices in one process all work
    cudaError_t e = cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return temd_set_error((int)e, "synth: cudaFuncSetAttribute failed");
    dim3 grid((rows + SY_BM - 1) / SY_BM, (ncol + SY_BN - 1) / SY_BN);
    const int nkb = (lpad + SY_BK - 1) / SY_BK;
    k_synth<<<grid, SY_THREADS, smem, stream>>>(maps, rows, ncol, nkb, out, ld_out);
    rc = (int)cudaGetLastError();
    if (rc) return temd_set_error(rc, "synth: kernel launch failed");
    return 0;
}

}  // namespace temd
