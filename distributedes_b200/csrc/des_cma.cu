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

// TILE x TILE outputs per CTA, 256 threads as 16 x 16.  Each thread owns (TILE/16)^2 outputs arranged as blocks of
// 4 consecutive rows/columns spaced 64 apart (rows ty*4 + {0..3} + 64*g), so every shared-memory operand read is one
// conflict-free LDS.128 (16 FMA per LDS for the 128 tile).  k-panels of 16 members are double buffered: the next
// panel's global loads are in flight while the current one is multiplied.
template <int TILE>
__global__ void __launch_bounds__(kCmaThreads) cma_rank_mu_kernel(float *__restrict__ dC, const float *__restrict__ Y,
                                                                  const float *__restrict__ w, int64_t lambda, int64_t n,
                                                                  int tiles_per_side, int packed) {
    constexpr int G = TILE / 64;                 // 4-wide groups per thread and dimension (1 or 2)
    constexpr int MT = 4 * G;
    constexpr int LD = kCmaKP * TILE / kCmaThreads;   // elements each thread stages per operand and panel
    __shared__ __align__(16) float As[2][kCmaKP][TILE];   // w_k * Y[k][i0 + i]
    __shared__ __align__(16) float Bs[2][kCmaKP][TILE];   //       Y[k][j0 + j]
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

    float ra[LD], rb[LD];
    auto fetch = [&](int64_t k0) {
#pragma unroll
        for (int e = 0; e < LD; ++e) {
            const int idx = threadIdx.x + e * kCmaThreads;
            const int kk = idx / TILE, c = idx - kk * TILE;
            const int64_t k = k0 + kk;
            float a = 0.f, b = 0.f;
            if (k < lambda) {
                const float wk = __ldg(w + k);
                if (i0 + c < n) a = wk * __ldg(Y + k * n + i0 + c);
                if (j0 + c < n) b = __ldg(Y + k * n + j0 + c);
            }
            ra[e] = a;
            rb[e] = b;
        }
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int e = 0; e < LD; ++e) {
            const int idx = threadIdx.x + e * kCmaThreads;
            const int kk = idx / TILE, c = idx - kk * TILE;
            As[buf][kk][c] = ra[e];
            Bs[buf][kk][c] = rb[e];
        }
    };
    fetch(0);
    stage(0);
    __syncthreads();
    int buf = 0;
    for (int64_t k0 = 0; k0 < lambda; k0 += kCmaKP) {
        const bool more = k0 + kCmaKP < lambda;
        if (more) fetch(k0 + kCmaKP);            // global loads overlap the multiply below
#pragma unroll
        for (int kk = 0; kk < kCmaKP; ++kk) {
            float av[MT], bv[MT];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 a4 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4 + 64 * g]);
                const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4 + 64 * g]);
                av[4 * g] = a4.x; av[4 * g + 1] = a4.y; av[4 * g + 2] = a4.z; av[4 * g + 3] = a4.w;
                bv[4 * g] = b4.x; bv[4 * g + 1] = b4.y; bv[4 * g + 2] = b4.z; bv[4 * g + 3] = b4.w;
            }
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < MT; ++b) acc[a][b] = __fmaf_rn(av[a], bv[b], acc[a][b]);
        }
        if (more) {
            stage(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    if (packed) {
        // upper-triangular tiles only, tile after tile ([tile][TILE][TILE], the collective's payload: half the bytes of
        // the full matrix); entries beyond n are zero so that partial sums of different ranks can be added blindly
        float *tile = dC + (int64_t)blockIdx.x * TILE * TILE;
#pragma unroll
        for (int a = 0; a < MT; ++a) {
            const int li = ty * 4 + (a & 3) + 64 * (a >> 2);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int lj = tx * 4 + 64 * g;
                float4 v = make_float4(acc[a][4 * g], acc[a][4 * g + 1], acc[a][4 * g + 2], acc[a][4 * g + 3]);
                if (i0 + li >= n) v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + lj + 0 >= n) v.x = 0.f;
                if (j0 + lj + 1 >= n) v.y = 0.f;
                if (j0 + lj + 2 >= n) v.z = 0.f;
                if (j0 + lj + 3 >= n) v.w = 0.f;
                *reinterpret_cast<float4 *>(tile + li * TILE + lj) = v;
            }
        }
        return;
    }
#pragma unroll
    for (int a = 0; a < MT; ++a) {
        const int64_t i = i0 + ty * 4 + (a & 3) + 64 * (a >> 2);
#pragma unroll
        for (int b = 0; b < MT; ++b) {
            const int64_t j = j0 + tx * 4 + (b & 3) + 64 * (b >> 2);
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

// C <- decay*C + c1 pc pc^T + cmu*dC with dC given as packed upper-triangular tiles (cma_rank_mu_kernel, packed = 1).
// One CTA per 64 x 64 block of an upper tile: the block is staged in shared memory, applied to C[i-block][j-block] and,
// transposed, to C[j-block][i-block] — both coalesced.  Diagonal blocks take the j >= i entry for both sides, so C stays
// exactly symmetric.
template <int TILE>
__global__ void __launch_bounds__(256) cma_cov_apply_packed_kernel(float *__restrict__ C, const float *__restrict__ tiles,
                                                                   const float *__restrict__ pc, int64_t n, float decay,
                                                                   float c1, float cmu, int tiles_per_side) {
    constexpr int SB = TILE / 64;                     // 64 x 64 sub-blocks per tile side
    __shared__ float blk[64][65];
    int bi = 0, rem = blockIdx.x / (SB * SB);
    while (rem >= tiles_per_side - bi) { rem -= tiles_per_side - bi; ++bi; }
    const int bj = bi + rem;
    const int sub = blockIdx.x % (SB * SB), si = sub / SB, sj = sub % SB;
    const float *tile = tiles + (int64_t)(blockIdx.x / (SB * SB)) * TILE * TILE;
    const int64_t i0 = (int64_t)bi * TILE + si * 64, j0 = (int64_t)bj * TILE + sj * 64;
    if (bi == bj && sj < si) return;                  // lower sub-block of a diagonal tile: written by its mirror
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;       // 64 x 4
    for (int r = ty; r < 64; r += 4) blk[r][tx] = tile[(si * 64 + r) * TILE + sj * 64 + tx];
    __syncthreads();
    const bool diag = (i0 == j0);
    for (int r = ty; r < 64; r += 4) {                // C[i0 + r][j0 + tx]
        const int64_t i = i0 + r, j = j0 + tx;
        if (i < n && j < n) {
            const float d = diag ? blk[min(r, tx)][max(r, tx)] : blk[r][tx];
            float v = decay * C[i * n + j];
            if (pc) v = __fmaf_rn(c1 * __ldg(pc + i), __ldg(pc + j), v);
            C[i * n + j] = __fmaf_rn(cmu, d, v);
        }
    }
    if (!diag) {
        for (int r = ty; r < 64; r += 4) {            // mirror: C[j0 + r][i0 + tx] = f(blk[tx][r])
            const int64_t i = j0 + r, j = i0 + tx;
            if (i < n && j < n) {
                float v = decay * C[i * n + j];
                if (pc) v = __fmaf_rn(c1 * __ldg(pc + i), __ldg(pc + j), v);
                C[i * n + j] = __fmaf_rn(cmu, blk[tx][r], v);
            }
        }
    }
}

}  // namespace des

static int cma_tile(int64_t n) { return n <= 2048 ? 64 : 128; }

extern "C" DES_API int64_t des_cma_packed_elems(int64_t n) {
    if (n <= 0) return 0;
    const int64_t tile = cma_tile(n), t = (n + tile - 1) / tile;
    return t * (t + 1) / 2 * tile * tile;
}

extern "C" DES_API int des_cma_rank_mu_packed(float *tiles_out_dev, const float *Y_dev, const float *w_dev, int64_t lambda_local,
                                              int64_t n, void *stream) {
    using namespace des;
    DES_REQUIRE(n > 0 && lambda_local >= 0, "des_cma_rank_mu_packed: bad sizes lambda=%lld n=%lld", (long long)lambda_local,
                (long long)n);
    DES_REQUIRE(n <= 46340 * 16, "des_cma_rank_mu_packed: n too large");
    DES_REQUIRE(tiles_out_dev && (lambda_local == 0 || (Y_dev && w_dev)), "des_cma_rank_mu_packed: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (lambda_local == 0) {
        DES_CUDA(cudaMemsetAsync(tiles_out_dev, 0, (size_t)des_cma_packed_elems(n) * sizeof(float), st));
        return DES_OK;
    }
    const int tile = cma_tile(n), t = (int)((n + tile - 1) / tile);
    if (tile == 64) cma_rank_mu_kernel<64><<<(unsigned)(t * (t + 1) / 2), kCmaThreads, 0, st>>>(tiles_out_dev, Y_dev, w_dev, lambda_local, n, t, 1);
    else cma_rank_mu_kernel<128><<<(unsigned)(t * (t + 1) / 2), kCmaThreads, 0, st>>>(tiles_out_dev, Y_dev, w_dev, lambda_local, n, t, 1);
    DES_LAUNCH_CHECK("cma_rank_mu_kernel(packed)");
    return DES_OK;
}

extern "C" DES_API int des_cma_cov_apply_packed(float *C_dev, const float *tiles_dev, const float *pc_dev, int64_t n, double decay,
                                                double c1, double cmu, void *stream) {
    using namespace des;
    DES_REQUIRE(n > 0, "des_cma_cov_apply_packed: n=%lld", (long long)n);
    DES_REQUIRE(C_dev && tiles_dev, "des_cma_cov_apply_packed: NULL pointer");
    const int tile = cma_tile(n), t = (int)((n + tile - 1) / tile);
    const unsigned upper = (unsigned)(t * (t + 1) / 2);
    if (tile == 64)
        cma_cov_apply_packed_kernel<64><<<upper, 256, 0, (cudaStream_t)stream>>>(C_dev, tiles_dev, pc_dev, n, (float)decay, (float)c1, (float)cmu, t);
    else
        cma_cov_apply_packed_kernel<128><<<upper * 4, 256, 0, (cudaStream_t)stream>>>(C_dev, tiles_dev, pc_dev, n, (float)decay, (float)c1, (float)cmu, t);
    DES_LAUNCH_CHECK("cma_cov_apply_packed_kernel");
    return DES_OK;
}

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
        cma_rank_mu_kernel<64><<<(unsigned)(t * (t + 1) / 2), kCmaThreads, 0, st>>>(dC_out_dev, Y_dev, w_dev, lambda_local, n, t, 0);
    } else {
        const int t = (int)((n + 127) / 128);
        cma_rank_mu_kernel<128><<<(unsigned)(t * (t + 1) / 2), kCmaThreads, 0, st>>>(dC_out_dev, Y_dev, w_dev, lambda_local, n, t, 0);
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
