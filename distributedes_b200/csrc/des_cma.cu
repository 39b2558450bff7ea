// CMA-ES rank-mu covariance update (the arithmetic inside es.tell, cma_es.py:90; Hansen tutorial
// arXiv:1604.00772 eq. 47):   dC = sum_i w_i y_i y_i^T = Y^T diag(w) Y,    C <- decay*C + c1 pc pc^T + cmu dC
//
// SYRK-shaped: only tiles on or above the diagonal are computed and mirrored on store.  fp32 FFMA
// (the 1e-5 parity bar rules out single-pass TF32/BF16 tensor cores here), register
// micro-tiles fed from shared memory, k-panels of 16 members.
// Bound: CUDA-core FMA, lambda*n*(n+tile) flop with symmetry; traffic 4*lambda*n (Y) + 4*n*n (dC).
#include "des_common.cuh"

namespace des {

constexpr int kCmaThreads = 256;
constexpr int kCmaKP = 16;   // members per k-panel

// TILE x TILE outputs per CTA, 256 threads as 16x16, each (TILE/16)^2 outputs strided by 16 so that
// shared-memory reads are conflict-free broadcasts/rows.
template <int TILE>
__global__ void __launch_bounds__(kCmaThreads) cma_rank_mu_kernel(float *__restrict__ dC, const float *__restrict__ Y,
                                                                  const float *__restrict__ w, int64_t lambda, int64_t n,
                                                                  int tiles_per_side) {
    constexpr int MT = TILE / 16;
    __shared__ float As[kCmaKP][TILE + 4];   // w_k * Y[k][i0 + i]
    __shared__ float Bs[kCmaKP][TILE + 4];   //       Y[k][j0 + j]
    // linear block id -> (bi <= bj) upper-triangular tile
    int bi = 0, rem = blockIdx.x;
    while (rem >= tiles_per_side - bi) { rem -= tiles_per_side - bi; ++bi; }
    const int bj = bi + rem;
    const int64_t i0 = (int64_t)bi * TILE, j0 = (int64_t)bj * TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    float acc[MT][MT];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) acc[a][b] = 0.f;

    for (int64_t k0 = 0; k0 < lambda; k0 += kCmaKP) {
        for (int idx = threadIdx.x; idx < kCmaKP * TILE; idx += kCmaThreads) {
            const int kk = idx / TILE, c = idx - kk * TILE;
            const int64_t k = k0 + kk;
            float a = 0.f, b = 0.f;
            if (k < lambda) {
                const float wk = __ldg(w + k);
                if (i0 + c < n) a = wk * __ldg(Y + k * n + i0 + c);
                if (j0 + c < n) b = __ldg(Y + k * n + j0 + c);
            }
            As[kk][c] = a;
            Bs[kk][c] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kCmaKP; ++kk) {
            float av[MT], bv[MT];
#pragma unroll
            for (int a = 0; a < MT; ++a) av[a] = As[kk][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < MT; ++b) bv[b] = Bs[kk][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < MT; ++b) acc[a][b] = __fmaf_rn(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < MT; ++a) {
        const int64_t i = i0 + ty + 16 * a;
#pragma unroll
        for (int b = 0; b < MT; ++b) {
            const int64_t j = j0 + tx + 16 * b;
            if (i < n && j < n) {
                if (bi != bj) {
                    dC[i * n + j] = acc[a][b];
                    dC[j * n + i] = acc[a][b];
                } else if (j >= i) {
                    // diagonal tile: (i,j) and (j,i) are both computed; keep the j >= i one for exact symmetry
                    dC[i * n + j] = acc[a][b];
                    dC[j * n + i] = acc[a][b];
                }
            }
        }
    }
}

__global__ void cma_cov_apply_kernel(float *__restrict__ C, const float *__restrict__ dC, const float *__restrict__ pc,
                                     int64_t n, float decay, float c1, float cmu) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int64_t i = idx / n, j = idx - i * n;
    float v = decay * C[idx];
    if (pc) v = __fmaf_rn(c1 * __ldg(pc + i), __ldg(pc + j), v);
    C[idx] = __fmaf_rn(cmu, dC[idx], v);
}

}  // namespace des

extern "C" DES_API int des_cma_rank_mu(float *dC_out_dev, const float *Y_dev, const float *w_dev, int64_t lambda_local, int64_t n,
                               void *stream) {
    using namespace des;
    DES_REQUIRE(n > 0 && lambda_local >= 0, "des_cma_rank_mu: bad sizes lambda=%lld n=%lld", (long long)lambda_local,
                (long long)n);
    DES_REQUIRE(n <= 46340 * 16, "des_cma_rank_mu: n too large");
    DES_REQUIRE(dC_out_dev && (lambda_local == 0 || (Y_dev && w_dev)), "des_cma_rank_mu: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (lambda_local == 0) {
        DES_CUDA(cudaMemsetAsync(dC_out_dev, 0, (size_t)n * n * sizeof(float), st));
        return DES_OK;
    }
    if (n <= 2048) {
        const int t = (int)((n + 63) / 64);
        cma_rank_mu_kernel<64><<<(unsigned)(t * (t + 1) / 2), kCmaThreads, 0, st>>>(dC_out_dev, Y_dev, w_dev, lambda_local, n, t);
    } else {
        const int t = (int)((n + 127) / 128);
        cma_rank_mu_kernel<128><<<(unsigned)(t * (t + 1) / 2), kCmaThreads, 0, st>>>(dC_out_dev, Y_dev, w_dev, lambda_local, n, t);
    }
    DES_LAUNCH_CHECK("cma_rank_mu_kernel");
    return DES_OK;
}

extern "C" DES_API int des_cma_cov_apply(float *C_dev, const float *dC_dev, const float *pc_dev, int64_t n, double decay, double c1,
                                 double cmu, void *stream) {
    using namespace des;
    DES_REQUIRE(n > 0, "des_cma_cov_apply: n=%lld", (long long)n);
    DES_REQUIRE(C_dev && dC_dev, "des_cma_cov_apply: NULL pointer");
    const int64_t total = n * n;
    cma_cov_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(C_dev, dC_dev, pc_dev, n,
                                                                                          (float)decay, (float)c1, (float)cmu);
    DES_LAUNCH_CHECK("cma_cov_apply_kernel");
    return DES_OK;
}
