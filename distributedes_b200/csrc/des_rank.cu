// des_centered_rank: centered-rank fitness shaping, fitness_shift utils.py:142-148.
//
//   rank_i = #{j : f_j < f_i} + #{j < i : f_j == f_i}       (ascending; ties by index; NaN last)
//   s_i    = rank_i/(N-1) - 0.5
//
// Counting rank instead of a sort: a shard needs ranks only for ITS members but against ALL N
// fitnesses (ranks are global), so the work is n_local x N comparisons, embarrassingly parallel,
// integer-exact and deterministic.  Keys are order-preserving uint32 images of the floats; each
// (i, j) pair costs one 64-bit compare.  The j range is split over blockIdx.y, partial counts are
// combined with integer atomics (exact, order independent).
#include "des_common.cuh"

namespace des {

constexpr int kRankThreads = 256;
constexpr int kRankTile = 2048;   // keys staged in shared memory per step

// float -> uint32 whose unsigned order is the numpy sort order: -inf < ... < -0 == +0 < ... < +inf < NaN
__device__ __forceinline__ uint32_t order_key(float f) {
    uint32_t b = __float_as_uint(f);
    if ((b & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;   // any NaN -> last
    if ((b << 1) == 0u) b = 0u;                                 // -0 -> +0
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(kRankThreads) rank_count_kernel(int32_t *__restrict__ counts,
                                                                   const float *__restrict__ fitness, int64_t N,
                                                                   int64_t member_offset, int64_t n_local,
                                                                   int64_t j_per_block) {
    __shared__ uint32_t keys[kRankTile];
    const int64_t il = (int64_t)blockIdx.x * kRankThreads + threadIdx.x;
    const bool live = il < n_local;
    const int64_t ig = member_offset + il;
    const uint64_t mine = live ? (((uint64_t)order_key(__ldg(fitness + ig)) << 32) | (uint64_t)(uint32_t)ig) : 0;
    const int64_t j_begin = (int64_t)blockIdx.y * j_per_block;
    const int64_t j_end = min(N, j_begin + j_per_block);
    int32_t cnt = 0;
    for (int64_t j0 = j_begin; j0 < j_end; j0 += kRankTile) {
        const int n = (int)min((int64_t)kRankTile, j_end - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += kRankThreads) keys[t] = order_key(__ldg(fitness + j0 + t));
        __syncthreads();
        if (live) {
            const uint32_t jb = (uint32_t)j0;
#pragma unroll 8
            for (int t = 0; t < n; ++t) {
                const uint64_t other = ((uint64_t)keys[t] << 32) | (uint64_t)(jb + (uint32_t)t);
                cnt += (other < mine) ? 1 : 0;
            }
        }
    }
    if (live && cnt) atomicAdd(counts + il, cnt);
}

__global__ void rank_finish_kernel(float *__restrict__ shaped, int32_t *__restrict__ rank_out,
                                   const int32_t *__restrict__ counts, int64_t N, int64_t n_local) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const int32_t r = counts[i];
    if (rank_out) rank_out[i] = r;
    // utils.py:146-147 in fp64, then one rounding to fp32
    shaped[i] = (float)((double)r / (double)(N - 1) - 0.5);
}

}  // namespace des

extern "C" DES_API size_t des_rank_workspace_bytes(int64_t n_local) {
    return n_local > 0 ? (size_t)n_local * sizeof(int32_t) : 0;
}

extern "C" DES_API int des_centered_rank(float *shaped_out_dev, int32_t *rank_out_dev, const float *fitness_all_dev, int64_t N,
                                 int64_t member_offset, int64_t n_local, void *workspace_dev, size_t workspace_bytes,
                                 void *stream) {
    using namespace des;
    DES_REQUIRE(N >= 2, "des_centered_rank: N=%lld, need N >= 2 (utils.py:146 divides by N-1)", (long long)N);
    DES_REQUIRE(N <= ((int64_t)1 << 31) - 1, "des_centered_rank: N too large");
    DES_REQUIRE(member_offset >= 0 && n_local >= 0 && member_offset + n_local <= N,
                "des_centered_rank: shard [%lld, %lld) outside population of %lld", (long long)member_offset,
                (long long)(member_offset + n_local), (long long)N);
    if (n_local == 0) return DES_OK;
    DES_REQUIRE(shaped_out_dev && fitness_all_dev, "des_centered_rank: NULL pointer");
    if (!workspace_dev || workspace_bytes < des_rank_workspace_bytes(n_local)) {
        set_error("des_centered_rank: workspace %zu B < required %zu B", workspace_bytes,
                  des_rank_workspace_bytes(n_local));
        return DES_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *counts = (int32_t *)workspace_dev;
    DES_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_local * sizeof(int32_t), st));
    const unsigned bx = (unsigned)((n_local + kRankThreads - 1) / kRankThreads);
    // enough j-slices to fill the machine a few times over, each a multiple of the smem tile
    int64_t want_blocks = 148 * 8;
    int64_t by = (want_blocks + bx - 1) / bx;
    int64_t max_by = (N + kRankTile - 1) / kRankTile;
    if (by > max_by) by = max_by;
    if (by < 1) by = 1;
    if (by > 65535) by = 65535;
    int64_t j_per_block = (N + by - 1) / by;
    j_per_block = ((j_per_block + kRankTile - 1) / kRankTile) * kRankTile;
    by = (N + j_per_block - 1) / j_per_block;
    rank_count_kernel<<<dim3(bx, (unsigned)by), kRankThreads, 0, st>>>(counts, fitness_all_dev, N, member_offset,
                                                                       n_local, j_per_block);
    DES_LAUNCH_CHECK("rank_count_kernel");
    rank_finish_kernel<<<(unsigned)((n_local + 255) / 256), 256, 0, st>>>(shaped_out_dev, rank_out_dev, counts, N,
                                                                          n_local);
    DES_LAUNCH_CHECK("rank_finish_kernel");
    return DES_OK;
}
