// Hardware probe (development tool): tcgen05 cta_group::2 mechanics for a paired-CTA eval kernel.
//   * cluster of 2 CTAs, tcgen05.alloc.cta_group::2 by the same warp of both CTAs
//   * M = 256 = 2 x 128 rows: each CTA holds its 128 rows of A (TMEM, fp16 packed) and of D (TMEM, fp32)
//   * B [N=64 x K] split by N: CTA r holds rows [32r, 32r+32) in ITS shared memory at the same offset
//   * leader CTA waits on an mbarrier that both CTAs arrive on (remote arrive through mapa), issues
//     tcgen05.mma.cta_group::2, commits with .multicast::cluster to the barrier of both CTAs
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int N, int KATOMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe2_kernel(const __half *A, const __half *B, float *D) {
    constexpr int K = 64 * KATOMS, NH = N / 2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_ready, bar_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();

    if (tid == 0) {
        mbar_init(&bar_ready, 2);        // one arrive per CTA (meaningful in the leader)
        mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    cluster_sync();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    // this CTA's half of B: rows [NH*rank, NH*rank + NH)
    for (int idx = tid; idx < NH * KATOMS * 8; idx += 128) {
        const int c = idx & 7, r = (idx >> 3) % NH, ka = idx / (8 * NH);
        const uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)(rank * NH + r) * K + ka * 64 + c * 8);
        *reinterpret_cast<uint4 *>(smem + (size_t)ka * NH * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    constexpr int ACOL = 128;
    for (int ka = 0; ka < KATOMS; ++ka) {     // this CTA's 128 rows of A -> its TMEM
        uint32_t w[32];
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A + (size_t)(rank * 128 + tid) * K + ka * 64);
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = src[i];
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
            "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(tmem + lane_base + ACOL + ka * 32),
            "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
            "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]),
            "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]),
            "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid == 0) mbar_arrive_remote(&bar_ready, 0);          // tell the leader this CTA's operands are in place
    if (rank == 0 && tid == 0) {
        mbar_wait(&bar_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        for (int ka = 0; ka < KATOMS; ++ka) {
            const uint64_t bdesc0 = make_desc_sw128(smem_u32(smem + (size_t)ka * NH * 128));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t bdesc = bdesc0 + (uint64_t)(ks * 2);
                const uint32_t a_addr = tmem + ACOL + ka * 32 + ks * 8;
                const uint32_t accum = (ka | ks) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(a_addr),
                    "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
    }
    mbar_wait(&bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
            "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + lane_base + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) D[(size_t)(rank * 128 + tid) * N + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

template <int N, int KATOMS>
int run() {
    constexpr int K = 64 * KATOMS, M = 256;
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
    srand(7);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
            ref[m * N + n] = (float)s;
        }
    __half *dA, *dB; float *dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, out.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, out.size() * 4));
    const int smem = (N / 2) * K * 2 + 1024;
    CK(cudaFuncSetAttribute(probe2_kernel<N, KATOMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe2_kernel<N, KATOMS><<<2, 128, smem>>>(dA, dB, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double e = fabs((double)out[i] - ref[i]); if (!(e <= 1e-3)) ++bad; if (e > maxerr || e != e) maxerr = e; }
    printf("probe2 f16 TS cta_group::2 M=256 N=%d K=%d: max abs err %.3e, bad %d / %d  -> %s\n", N, K, maxerr, bad, M * N, bad ? "FAIL" : "OK");
    if (bad) {
        for (int m : {0, 128}) { printf(" row %d:", m); for (int n : {0, 1, 31, 32, 33, 63}) printf(" [%d] %.3f/%.3f", n, out[m * N + n], ref[m * N + n]); printf("\n"); }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad != 0;
}

int main() {
    int fail = 0;
    fail |= run<64, 1>();
    fail |= run<64, 4>();
    fail |= run<128, 2>();
    return fail;
}
