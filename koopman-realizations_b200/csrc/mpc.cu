// First consumer-side hot spot of the fitted bilinear model: Kmpc.get_costB_bilinear (Kmpc.m:569-596),
// the state-dependent condensed input matrix of the bilinear MPC problem, rebuilt at EVERY control step
// (Kmpc.m:get_costH_bilinear / get_costG_bilinear / get_costD_bilinear call it).
//
//   Bcol block i (i = 1..h) = A^(i-1) Beta(z_i),   Beta(z) = B kron(I_m, z)  (Ksysid.m:1288-1289), z_i = z(i,:) or z(1,:)
//   B(:, block c) = Lshift^(c-1) Bcol  (block Toeplitz: block (r, c) = Bcol block r - c + 1... as written: a lower shift by N
//   rows per block column), B is N (h+1) x m h with a zero first block row.
//
// One CTA per (horizon step i, problem): Beta(z_i) is formed in shared memory (N x m) and multiplied i-1 times by A
// (column-major: thread = output row, coalesced reads of A's columns from L2); the block is then scattered to every
// block column it appears in.  Work is tiny (N ~ 30-100, h ~ 10-50): the point is ONE launch per control step for a
// whole batch of problems instead of h interpreted matrix powers.
#include "kf_internal.h"

namespace {

__global__ void __launch_bounds__(256) kf_mpc_costB_kernel(int N, int m, int h, int nz, const double* __restrict__ A,
                                                           const double* __restrict__ B, const double* __restrict__ Z, double* __restrict__ out,
                                                           long long z_stride, long long out_stride) {
    extern __shared__ __align__(16) double sm[];
    double* cur = sm;                 // N x m
    double* nxt = cur + (size_t)N * m;
    double* zs = nxt + (size_t)N * m; // N
    const int i = blockIdx.x + 1;     // horizon step 1..h
    const int prob = blockIdx.y;
    const double* z = Z + (size_t)prob * z_stride;          // nz x N column-major
    double* Bo = out + (size_t)prob * out_stride;           // N (h+1) x m h column-major
    const int zi = nz > 1 ? i - 1 : 0;
    for (int k = threadIdx.x; k < N; k += blockDim.x) zs[k] = z[(size_t)k * nz + zi];
    __syncthreads();
    // Beta(z) = B kron(I_m, z): column q = B(:, q N : (q+1) N) z
    for (int e = threadIdx.x; e < N * m; e += blockDim.x) {
        const int r = e % N, q = e / N;
        const double* Bq = B + (size_t)q * N * N;
        double acc = 0.0;
        for (int k = 0; k < N; ++k) acc = fma(Bq[(size_t)k * N + r], zs[k], acc);
        cur[q * N + r] = acc;
    }
    __syncthreads();
    for (int p = 1; p < i; ++p) {     // A^(i-1) Beta: i-1 products with A
        for (int e = threadIdx.x; e < N * m; e += blockDim.x) {
            const int r = e % N, q = e / N;
            double acc = 0.0;
            for (int k = 0; k < N; ++k) acc = fma(A[(size_t)k * N + r], cur[q * N + k], acc);
            nxt[q * N + r] = acc;
        }
        __syncthreads();
        double* t = cur; cur = nxt; nxt = t;
    }
    // scatter: block row i of Bcol appears at block (row i + c, column c), c = 0 .. h - i   (lower shift per block column)
    const long long ldo = (long long)N * (h + 1);
    for (int c = 0; c + i <= h; ++c)
        for (int e = threadIdx.x; e < N * m; e += blockDim.x) {
            const int r = e % N, q = e / N;
            Bo[(size_t)(c * m + q) * ldo + (size_t)(i + c) * N + r] = cur[q * N + r];
        }
}

}  // namespace

extern "C" int kf_mpc_costB_bilinear(kf_ctx* ctx, int N, int m, int horizon, int nbatch, const double* A, const double* B, int nz,
                                     const double* z, double* Bout) {
    if (!ctx) return KF_EINVAL;
    if (N < 1 || m < 1 || horizon < 1 || nbatch < 1 || !A || !B || !z || !Bout || (nz != 1 && nz != horizon)) {
        ctx->err = "kf_mpc_costB_bilinear: N, m, horizon, nbatch >= 1, A, B, z, Bout required; z has 1 row or `horizon` rows";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nA = (size_t)N * N, nB = nA * m, nZ = (size_t)nz * N, nO = (size_t)N * (horizon + 1) * m * horizon;
    KF_CUDA(ctx, ctx->d_qr.ensure((nA + nB + nZ * nbatch + nO * nbatch) * sizeof(double)));
    double* dA = ctx->d_qr.as<double>();
    double* dB = dA + nA;
    double* dZ = dB + nB;
    double* dO = dZ + nZ * nbatch;
    KF_CUDA(ctx, cudaMemcpyAsync(dA, A, nA * sizeof(double), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(dB, B, nB * sizeof(double), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(dZ, z, nZ * nbatch * sizeof(double), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemsetAsync(dO, 0, nO * nbatch * sizeof(double), st));
    const size_t smem = ((size_t)2 * N * m + N) * sizeof(double);
    KF_CUDA(ctx, kf_ensure_smem(ctx, kf_mpc_costB_kernel, smem));
    kf_mpc_costB_kernel<<<dim3(horizon, nbatch), 256, smem, st>>>(N, m, horizon, nz, dA, dB, dZ, dO, (long long)nZ, (long long)nO);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    KF_CUDA(ctx, cudaMemcpyAsync(Bout, dO, nO * nbatch * sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    return KF_OK;
}
