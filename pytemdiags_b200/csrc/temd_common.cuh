// Shared device helpers for the temd kernels (sm_100a only): DMMA, TMA, mbarrier, swizzle.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace temd {

// ---------------------------------------------------------------------------------------------
// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// lane = 4*g + t:  a = A[g][t],  b = B[t][g],  d0,d1 = D[g][2t], D[g][2t+1].
// Measured 37.1 TFLOP/s on B200 (profiles/r01_microbench_fp64.log) = 64 FMA/clk/SM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};\n" :: "r"(addr), "d"(x), "d"(y) : "memory");
}

// ---------------------------------------------------------------------------------------------
// mbarrier (shared::cta) wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA: 2-D tiled bulk tensor load, global -> shared, completion on an mbarrier.  SASS: UTMALDG.
// c0 = innermost coordinate (elements), c1 = row coordinate.  Out-of-bounds elements read as 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" :: "l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Shared-memory tile geometry.  Every TMA box is [rows][16 doubles] = rows x 128 B written with
// CU_TENSOR_MAP_SWIZZLE_128B into a 1024-B-aligned buffer: the 16-B chunk index (0..7) of an
// element is XORed with (row & 7).  Byte offset of element (row, col) inside a tile:
// ---------------------------------------------------------------------------------------------
constexpr int TILE_K = 16;             // doubles per tile row (128 B)
constexpr int TILE_ROW_BYTES = 128;

__device__ __forceinline__ uint32_t swz_off(int row, int col) {
    return static_cast<uint32_t>(row * TILE_ROW_BYTES + ((((col >> 1) ^ row) & 7) << 4) + ((col & 1) << 3));
}

// "K-major" fragment column permutation: both MMA operands are read as tile[row = base+g][k] and
// the contraction index k inside a 16-wide tile is split over 4 DMMA steps kk = 0..3 as
//     k(t, kk) = 8*(t>>1) + 2*kk + (t&1)
// (any permutation of k is legal as long as A and B agree).  With the 128-B swizzle this makes
// every LDS.64 of a half-warp hit 16 distinct 8-byte bank pairs: conflict-free.
// Byte offset inside the 128-B row for lane (g, t) at step kk (row & 7 == g):
__device__ __forceinline__ uint32_t kmajor_col_off(int g, int t, int kk) {
    return static_cast<uint32_t>(((((t >> 1) << 2) + kk) ^ g) << 4) + static_cast<uint32_t>((t & 1) << 3);
}

// "MN-major" B operand (contraction index runs over tile ROWS, the n index over tile columns):
// inside an 8-row group the contraction is split over 2 DMMA steps c = 0, 1 as
//     k(t, c) = {0,5,2,7}[t]  (c = 0),   {1,4,3,6}[t]  (c = 1)
// so that (a) the swizzled B loads are conflict-free and (b) an A operand stored row-major with a
// row stride = 4 or 12 (mod 16) doubles is conflict-free too.
__device__ __forceinline__ int mnmajor_k(int t, int c) {
    return (((t & 2) >> 1) | ((t & 1) << 1)) * 2 + (c ^ (t & 1));
}

struct alignas(64) TmaMap { CUtensorMap m; };

}  // namespace temd
