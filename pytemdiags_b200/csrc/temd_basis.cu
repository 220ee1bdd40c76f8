// K1 `basis_build` and K3 Cholesky / triangular inverse.
//
// K1 replaces the Python loops `Y0[:,ll] = sph_harm(0, ll, 0, coalt).real` (reference
// PyTEMDiags/sph_zonal_mean.py:359-370) with the normalised three-term Legendre recurrence, one
// thread per column, written transposed (QT0[l][n], n contiguous) so every degree is one coalesced
// row store.  K3 replaces `lstsq(Y0, identity(N))` (sph_zonal_mean.py:389): the Gram matrix
// G = Y0^T Y0 (computed by K4 `project`) is Cholesky-factored and L^-1 formed explicitly so the basis
// can be orthonormalised once (Q = Y0 L^-T); see DESIGN.md "whitened basis".
#include "temd_common.cuh"
#include "temd_internal.h"

namespace temd {

// QT0[l][n] = sqrt((2l+1)/4pi) P_l(x_n);  rows l in [L+1, lpad) and columns n in [n, ld) are zeroed.
__global__ void k_basis(const double* __restrict__ x, int n, int L, const double* __restrict__ rec_a,
                        const double* __restrict__ rec_b, double* __restrict__ qt, size_t ld, int lpad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    if (i >= (size_t)n) {
        for (int l = 0; l < lpad; l++) qt[(size_t)l * ld + i] = 0.0;
        return;
    }
    const double xi = x[i];
    double pm2 = rec_a[0];            // Y_0 = sqrt(1/4pi)
    qt[i] = pm2;
    if (L >= 1) {
        double pm1 = rec_a[1] * xi;   // Y_1 = sqrt(3/4pi) x
        qt[ld + i] = pm1;
        for (int l = 2; l <= L; l++) {
            const double p = rec_a[l] * xi * pm1 - rec_b[l] * pm2;
            qt[(size_t)l * ld + i] = p;
            pm2 = pm1;
            pm1 = p;
        }
    }
    for (int l = L + 1; l < lpad; l++) qt[(size_t)l * ld + i] = 0.0;
}

int launch_basis(const double* x, int n, int L, const double* rec_a, const double* rec_b, double* qt, size_t ld,
                 int lpad, cudaStream_t stream) {
    const int threads = 128;
    k_basis<<<(unsigned)((ld + threads - 1) / threads), threads, 0, stream>>>(x, n, L, rec_a, rec_b, qt, ld, lpad);
    return (int)cudaGetLastError();
}

// Latitude derivative of the basis, D[l][n] = d/dphi [ sqrt((2l+1)/4pi) P_l(sin phi) ] = cos(phi) Y_l'(x), x = sin(phi),
// from the differentiated recurrence  Y_l' = a_l (Y_{l-1} + x Y_{l-1}') - b_l Y_{l-2}'  (bounded at the poles, where
// the closed form l (P_{l-1} - x P_l) / cos(phi) is 0/0).  Optional Legendre-space derivative output (BASELINE.json
// north_star); the reference itself differentiates by finite differences (tem_util.py:154), so nothing on the
// default path uses it.
__global__ void k_basis_dlat(const double* __restrict__ x, int n, int L, const double* __restrict__ rec_a,
                             const double* __restrict__ rec_b, double* __restrict__ dt, size_t ld, int lpad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    if (i >= (size_t)n) {
        for (int l = 0; l < lpad; l++) dt[(size_t)l * ld + i] = 0.0;
        return;
    }
    const double xi = x[i];
    const double c = sqrt(fmax(0.0, (1.0 - xi) * (1.0 + xi)));      // cos(phi)
    double pm2 = rec_a[0], dm2 = 0.0;                               // Y_0, Y_0'
    dt[i] = 0.0;
    if (L >= 1) {
        double pm1 = rec_a[1] * xi, dm1 = rec_a[1];                 // Y_1, Y_1'
        dt[ld + i] = c * dm1;
        for (int l = 2; l <= L; l++) {
            const double p = rec_a[l] * xi * pm1 - rec_b[l] * pm2;
            const double d = rec_a[l] * (pm1 + xi * dm1) - rec_b[l] * dm2;
            dt[(size_t)l * ld + i] = c * d;
            pm2 = pm1; pm1 = p;
            dm2 = dm1; dm1 = d;
        }
    }
    for (int l = L + 1; l < lpad; l++) dt[(size_t)l * ld + i] = 0.0;
}

int launch_basis_dlat(const double* x, int n, int L, const double* rec_a, const double* rec_b, double* dt, size_t ld,
                      int lpad, cudaStream_t stream) {
    const int threads = 128;
    k_basis_dlat<<<(unsigned)((ld + threads - 1) / threads), threads, 0, stream>>>(x, n, L, rec_a, rec_b, dt, ld, lpad);
    return (int)cudaGetLastError();
}

// Single-CTA left-looking Cholesky G = L L^T (G symmetric n x n, leading dimension ldg) followed by the
// explicit lower-triangular inverse.  LT is scratch [n][n] holding L transposed (LT[k][i] = L[i][k]) so
// the row-parallel inner products are coalesced.  Linv is [lpad][ldi], zero outside the n x n triangle.
// status: 0 ok, j+1 if pivot j is not positive / numerically zero (rank-deficient basis).
__global__ void __launch_bounds__(1024, 1)
k_chol_inv(const double* __restrict__ G, int ldg, int n, double* __restrict__ LT, double* __restrict__ Linv, int ldi,
           int lpad, int* __restrict__ status) {
    const int tid = threadIdx.x;
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (int j = 0; j < n; j++) {
        for (int i = j + tid; i < n; i += blockDim.x) {
            double s = G[(size_t)j * ldg + i];
            for (int k = 0; k < j; k++) s -= LT[(size_t)k * n + i] * LT[(size_t)k * n + j];
            LT[(size_t)j * n + i] = s;
        }
        __syncthreads();
        const double d = LT[(size_t)j * n + j];
        const double gjj = G[(size_t)j * ldg + j];
        if (!(d > 1e-13 * gjj) || !(gjj > 0.0)) {
            if (tid == 0) { s_fail = j + 1; }
        }
        __syncthreads();
        if (s_fail) break;
        const double inv = 1.0 / sqrt(d);
        for (int i = j + tid; i < n; i += blockDim.x)
            LT[(size_t)j * n + i] = (i == j) ? sqrt(d) : LT[(size_t)j * n + i] * inv;
        __syncthreads();
    }
    if (tid == 0) *status = s_fail;
    // zero-fill Linv
    for (size_t e = tid; e < (size_t)lpad * ldi; e += blockDim.x) Linv[e] = 0.0;
    __syncthreads();
    if (s_fail) return;
    // column c of L^-1 by forward substitution, one thread per column
    for (int c = tid; c < n; c += blockDim.x) {
        for (int i = c; i < n; i++) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; k++) s -= LT[(size_t)k * n + i] * Linv[(size_t)k * ldi + c];
            Linv[(size_t)i * ldi + c] = s / LT[(size_t)i * n + i];
        }
    }
}

int launch_chol_inv(const double* G, int ldg, int n, double* LT, double* Linv, int ldi, int lpad, int* status,
                    cudaStream_t stream) {
    k_chol_inv<<<1, 1024, 0, stream>>>(G, ldg, n, LT, Linv, ldi, lpad, status);
    return (int)cudaGetLastError();
}

// C = A * B for small n x n (leading dimension ld) matrices: combines the two CholeskyQR passes.
__global__ void k_matmul_small(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int n,
                               int ld) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n || i >= n) return;
    double s = 0.0;
    for (int k = 0; k < n; k++) s += A[(size_t)i * ld + k] * B[(size_t)k * ld + j];
    C[(size_t)i * ld + j] = s;
}

int launch_matmul_small(const double* A, const double* B, double* C, int n, int ld, cudaStream_t stream) {
    dim3 grid((n + 127) / 128, n);
    k_matmul_small<<<grid, 128, 0, stream>>>(A, B, C, n, ld);
    return (int)cudaGetLastError();
}

}  // namespace temd
