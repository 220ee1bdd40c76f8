// Device generator of the synthetic benchmark fields; mirrors
// pytemdiags_b200/synthetic.py::synth_fields (same closed forms, same counter-based hash noise).
// Not part of the reference: there are no datasets in the build/bench environment.
#include "temd_common.cuh"
#include "temd_internal.h"

namespace temd {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_synth_fields(double* __restrict__ out, int field, unsigned long long key, int t0, int nlev, int ncol,
                               size_t ld, const double* __restrict__ lat, const double* __restrict__ lon,
                               const double* __restrict__ plev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncol) return;
    const int k = blockIdx.y, it = blockIdx.z;
    const double phi = lat[i], lam = lon[i], s = plev[k] / 1000.0, t = (double)(t0 + it);
    const double c = cos(phi);
    const unsigned long long idx = ((unsigned long long)(t0 + it) * nlev + k) * (unsigned long long)ncol + i;
    const unsigned long long z = splitmix64(idx ^ key);
    const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    const double xi = (2.0 * u - 1.0) * sqrt(3.0);
    double v;
    if (field == 2) v = 210.0 + 75.0 * pow(s, 0.19) * c * c + 3.0 * sin(3 * lam + 0.3 * t) * c * c * c * s + 0.5 * xi;
    else if (field == 0) { const double s2 = sin(2 * phi); v = 30.0 * s2 * s2 * (1.0 - s) + 5.0 * cos(4 * lam - 0.2 * t) * c * c * c * c + 2.0 * xi; }
    else if (field == 1) v = sin(2 * phi) * s + 4.0 * sin(4 * lam - 0.2 * t + 0.5) * c * c * c * c + 2.0 * xi;
    else if (field == 3) v = 0.01 * cos(3 * phi) * s + 0.05 * sin(3 * lam + 0.3 * t + 1.0) * c * c * c * s + 0.02 * xi;
    else v = 1e-3 * s * s * c * c * (1.0 + 0.3 * sin(2 * lam + 0.1 * t)) + 1e-5 * xi;
    out[((size_t)it * nlev + k) * ld + i] = v;
}

static unsigned long long host_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int launch_synth_fields(double* out, int field, int seed, int t0, int nt, int nlev, int ncol, size_t ld,
                        const double* lat_rad, const double* lon_rad, const double* plev_hpa, cudaStream_t stream) {
    const unsigned long long key = host_splitmix64((unsigned long long)seed * 8ull + (unsigned long long)field);
    dim3 grid((ncol + 255) / 256, nlev, nt);
    k_synth_fields<<<grid, 256, 0, stream>>>(out, field, key, t0, nlev, ncol, ld, lat_rad, lon_rad, plev_hpa);
    return (int)cudaGetLastError();
}

}  // namespace temd
