// Hardware probe (development tool, not product code): checks the tcgen05 layout assumptions the
// tensor-core eval kernel relies on, against a host reference:
//   * kind::f16 MMA, M=128, A from TMEM (fp16 packed 2 per 32-bit column, lane = row), B from shared
//     memory K-major SWIZZLE_128B, fp32 accumulator in TMEM (lane = row, column = n)
//   * descriptor K-advance by +32 B inside the 128 B swizzle atom, multiple atoms along K
//   * tcgen05.st / tcgen05.ld 32x32b addressing, tcgen05.commit -> mbarrier
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ; run on a B200.
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
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B: start>>4 | LBO=1 (unused) | SBO=1024B>>4 | version=1 | layout=2
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int N, int KATOMS>
__global__ void __launch_bounds__(128) probe_kernel(const __half *A, const __half *B, float *D) {
    // A [128][64*KATOMS], B [N][64*KATOMS] row-major fp16; D [128][N] fp32
    constexpr int K = 64 * KATOMS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // B -> smem: atom ka = [N rows][128 B], 16-byte chunk c of row r stored at chunk (c ^ (r & 7))
    for (int idx = tid; idx < N * KATOMS * 8; idx += 128) {
        const int c = idx & 7, r = (idx >> 3) % N, ka = idx / (8 * N);
        const uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)r * K + ka * 64 + c * 8);
        *reinterpret_cast<uint4 *>(smem + (size_t)ka * N * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    constexpr int ACOL = 256;   // A at columns [256, 256 + K/2), D at [0, N)
    // A row (this thread = lane tid) -> TMEM, 32 packed words per atom
    for (int ka = 0; ka < KATOMS; ++ka) {
        uint32_t w[32];
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A + (size_t)tid * K + ka * 64);
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
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        // idesc: c=f32 (1<<4), a=b=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ka = 0; ka < KATOMS; ++ka) {
            const uint64_t bdesc0 = make_desc_sw128(smem_u32(smem + (size_t)ka * N * 128));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t bdesc = bdesc0 + (uint64_t)(ks * 2);        // +32 B in 16-byte units
                const uint32_t a_addr = tmem + ACOL + ka * 32 + ks * 8;     // 16 fp16 = 8 columns
                const uint32_t accum = (ka | ks) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(a_addr),
                    "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
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
        for (int i = 0; i < 32; ++i) D[(size_t)tid * N + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, int KATOMS>
int run() {
    constexpr int K = 64 * KATOMS;
    std::vector<__half> hA(128 * K), hB(N * K);
    std::vector<float> fA(128 * K), fB(N * K), ref(128 * N), out(128 * N);
    srand(1);
    for (int i = 0; i < 128 * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    for (int m = 0; m < 128; ++m)
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
    const int smem = N * K * 2 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel<N, KATOMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<N, KATOMS><<<1, 128, smem>>>(dA, dB, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < 128 * N; ++i) { double e = fabs((double)out[i] - ref[i]); if (!(e <= 1e-3)) ++bad; if (e > maxerr || e != e) maxerr = e; }
    printf("probe f16 TS M=128 N=%d K=%d: max abs err %.3e, bad %d / %d  -> %s\n", N, K, maxerr, bad, 128 * N, bad ? "FAIL" : "OK");
    if (bad) {
        for (int m = 0; m < 2; ++m) { printf(" row %d:", m); for (int n = 0; n < 8; ++n) printf(" %.3f/%.3f", out[m * N + n], ref[m * N + n]); printf("\n"); }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad != 0;
}

int main() {
    int fail = 0;
    fail |= run<128, 1>();
    fail |= run<128, 2>();
    fail |= run<256, 4>();
    fail |= run<64, 1>();
    return fail;
}
