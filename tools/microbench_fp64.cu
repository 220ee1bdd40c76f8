// FP64 issue-rate micro-benchmarks for B200 (sm_100a): scalar DFMA vs warp-level DMMA.
// Answers SURVEY.md §7.2: "measure a DMMA issue-rate micro-benchmark on the box".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NT>
__global__ void __launch_bounds__(1024) k_dmma884(double* out, int iters, double a, double b) {
    double c0[NT], c1[NT];
#pragma unroll
    for (int i = 0; i < NT; i++) { c0[i] = i; c1[i] = -i; }
    double fa = a + threadIdx.x * 1e-9, fb = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NT; i++) dmma884(c0[i], c1[i], fa, fb);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; i++) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NT>
__global__ void __launch_bounds__(1024) k_dmma16816(double* out, int iters, double a0, double b0) {
    double c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = a0 + i * 1e-9 + threadIdx.x * 1e-10;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = b0 + i * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NT; i++) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <typename F>
double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    double* out; CK(cudaMalloc(&out, 8));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s SMs %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    int iters = 20000;
    int reps = argc > 1 ? atoi(argv[1]) : 5;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int grid = sms * bps;
            {
                double ms = time_ms([&] { k_dfma<16><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, reps);
                double fl = 2.0 * 16 * iters * (double)threads * grid;
                printf("DFMA        thr %4d bps %d : %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { k_dmma884<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, reps);
                double fl = 2.0 * 256 * 8 * iters * (double)(threads / 32) * grid;
                printf("DMMA m8n8k4 thr %4d bps %d : %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { k_dmma16816<4><<<grid, threads>>>(out, iters / 4, 1.0000001, 1e-9); }, reps);
                double fl = 2.0 * 16 * 8 * 16 * 4 * (iters / 4) * (double)(threads / 32) * grid;
                printf("DMMA m16n8k16 thr %4d bps %d : %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
            }
        }
    }
    // sustained (~2 s each) at the best-looking config to expose power capping
    for (int which = 0; which < 2; which++) {
        int threads = 512, grid = sms * 2;
        double ms = time_ms([&] {
            if (which == 0) k_dfma<16><<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
            else k_dmma884<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 200);
        double fl = which == 0 ? 2.0 * 16 * iters * (double)threads * grid : 2.0 * 256 * 8 * iters * (double)(threads / 32) * grid;
        printf("SUSTAINED %s : %8.3f ms/launch  %7.2f TFLOP/s\n", which == 0 ? "DFMA" : "DMMA m8n8k4", ms, fl / ms * 1e-9);
    }
    return 0;
}
