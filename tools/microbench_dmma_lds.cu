// DMMA fed from shared memory: how close to the 37.1 TFLOP/s issue peak do the three inner-loop shapes
// of the temd kernels get (project: 2x13 tiles/warp, eddy GEMM1: 2x2, eddy GEMM2: 3x7), 8 warps per SM?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(unsigned addr) { double v; asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr)); return v; }

template <int MI, int NJ, int WARPS, int UNROLL>
__global__ void __launch_bounds__(WARPS * 32, 1) k(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + g * 128 + (((((t >> 1) << 2) + 1) ^ g) << 4) + ((t & 1) << 3);   // conflict-free K-major fragment pattern
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0;
#pragma unroll UNROLL
    for (int it = 0; it < iters; it++) {
        const unsigned st = base + (it & 7) * 8192;
        double a[MI], b[NJ];
#pragma unroll
        for (int i = 0; i < MI; i++) a[i] = lds64(st + i * 1024);
#pragma unroll
        for (int j = 0; j < NJ; j++) b[j] = lds64(st + 32768 + j * 1024);
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[0] = s;
}

// explicit software pipelining: fragments of step it+1 are loaded before the DMMAs of step it
template <int MI, int NJ, int WARPS, int UNROLL>
__global__ void __launch_bounds__(WARPS * 32, 1) kp(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + g * 128 + (((((t >> 1) << 2) + 1) ^ g) << 4) + ((t & 1) << 3);   // conflict-free K-major fragment pattern
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0;
    double a[2][MI], b[2][NJ];
#pragma unroll
    for (int i = 0; i < MI; i++) a[0][i] = lds64(base + i * 1024);
#pragma unroll
    for (int j = 0; j < NJ; j++) b[0][j] = lds64(base + 32768 + j * 1024);
#pragma unroll 1
    for (int it = 0; it < iters; it += 2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const unsigned st = base + ((it + h + 1) & 7) * 8192;
#pragma unroll
            for (int i = 0; i < MI; i++) a[h ^ 1][i] = lds64(st + i * 1024);
#pragma unroll
            for (int j = 0; j < NJ; j++) b[h ^ 1][j] = lds64(st + 32768 + j * 1024);
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[h][i], b[h][j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[0] = s;
}

template <int MI, int NJ, int WARPS, int UNROLL>
void runp(const char* name, double* out, int sms) {
    const int iters = (200000 / (MI * NJ)) & ~1;
    CK(cudaFuncSetAttribute(kp<MI, NJ, WARPS, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kp<MI, NJ, WARPS, UNROLL><<<sms, WARPS * 32, 160 * 1024>>>(out, iters); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 3; r++) kp<MI, NJ, WARPS, UNROLL><<<sms, WARPS * 32, 160 * 1024>>>(out, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
    double fl = 2.0 * 256 * MI * NJ * (double)iters * WARPS * sms;
    printf("PIPELINED %-18s MI %d NJ %2d warps %2d : %7.3f ms %6.2f TFLOP/s (%.1f%%)\n", name, MI, NJ, WARPS, ms, fl / ms * 1e-9, fl / ms * 1e-9 / 37.1 * 100);
}

template <int MI, int NJ, int WARPS, int UNROLL>
void run(const char* name, double* out, int sms) {
    const int iters = 200000 / (MI * NJ);
    CK(cudaFuncSetAttribute(k<MI, NJ, WARPS, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MI, NJ, WARPS, UNROLL><<<sms, WARPS * 32, 160 * 1024>>>(out, iters); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 3; r++) k<MI, NJ, WARPS, UNROLL><<<sms, WARPS * 32, 160 * 1024>>>(out, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
    double fl = 2.0 * 256 * MI * NJ * (double)iters * WARPS * sms;
    printf("%-28s MI %d NJ %2d warps %2d unroll %d : %7.3f ms %6.2f TFLOP/s (%.1f%%)\n", name, MI, NJ, WARPS, UNROLL, ms, fl / ms * 1e-9, fl / ms * 1e-9 / 37.1 * 100);
}


// gemm2-like with the A operands formed by NMUL DMULs per step (as in k_eddy's GEMM2)
template <int NJ, int NMUL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) km(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + g * 128 + (((((t >> 1) << 2) + 1) ^ g) << 4) + ((t & 1) << 3);
    double acc[3][NJ][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const unsigned st = base + (it & 7) * 8192;
        double e[4], b[NJ], a[3];
#pragma unroll
        for (int i = 0; i < 4; i++) e[i] = lds64(st + i * 1024);
#pragma unroll
        for (int j = 0; j < NJ; j++) b[j] = lds64(st + 32768 + j * 1024);
        if (NMUL == 3) { a[0] = e[0] * e[1]; a[1] = e[0] * e[3]; a[2] = e[1] * e[2]; }
        else { a[0] = e[0]; a[1] = e[1]; a[2] = e[2]; }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[0] = s;
}
template <int NJ, int NMUL, int WARPS>
void runm(double* out, int sms) {
    const int iters = 200000 / (3 * NJ);
    CK(cudaFuncSetAttribute(km<NJ, NMUL, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    km<NJ, NMUL, WARPS><<<sms, WARPS * 32, 160 * 1024>>>(out, iters); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 3; r++) km<NJ, NMUL, WARPS><<<sms, WARPS * 32, 160 * 1024>>>(out, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
    double fl = 2.0 * 256 * 3 * NJ * (double)iters * WARPS * sms;
    printf("DMUL-test NJ %d nmul %d warps %d : %7.3f ms %6.2f TFLOP/s (%.1f%%)\n", NJ, NMUL, WARPS, ms, fl / ms * 1e-9, fl / ms * 1e-9 / 37.1 * 100);
}
template <int MI>
void sweep_row(double* out, int sms) {
    runp<MI, 1, 8, 1>("sweep", out, sms); runp<MI, 2, 8, 1>("sweep", out, sms); runp<MI, 3, 8, 1>("sweep", out, sms);
    runp<MI, 4, 8, 1>("sweep", out, sms); runp<MI, 5, 8, 1>("sweep", out, sms); runp<MI, 6, 8, 1>("sweep", out, sms);
    runp<MI, 7, 8, 1>("sweep", out, sms); runp<MI, 8, 8, 1>("sweep", out, sms); runp<MI, 10, 8, 1>("sweep", out, sms);
    runp<MI, 13, 8, 1>("sweep", out, sms);
}

int main(int argc, char** argv) {
    double* out; CK(cudaMalloc(&out, 8));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    runm<7, 0, 8>(out, sms); runm<7, 3, 8>(out, sms); runm<6, 3, 8>(out, sms); runm<4, 3, 8>(out, sms); runm<4, 0, 8>(out, sms); runm<4, 3, 16>(out, sms);
    return 0;
}
