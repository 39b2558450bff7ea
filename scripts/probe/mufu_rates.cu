// Throughput of the MUFU (XU pipe) operations the ES kernels use, per SM: independent chains, 32 warps per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rates mufu_rates.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define OP(name, asmstr)                                                                     \
    __global__ void k_##name(float *out, float seed, int iters) {                            \
        float a = seed + threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;      \
        for (int i = 0; i < iters; ++i) {                                                    \
            asm volatile(asmstr : "+f"(a)); asm volatile(asmstr : "+f"(b));                  \
            asm volatile(asmstr : "+f"(c)); asm volatile(asmstr : "+f"(d));                  \
        }                                                                                    \
        if (a + b + c + d == 123.456f) out[0] = a;                                           \
    }
OP(ex2, "ex2.approx.ftz.f32 %0, %0;")
OP(rcp, "rcp.approx.ftz.f32 %0, %0;")
OP(lg2, "lg2.approx.ftz.f32 %0, %0;")
OP(sqrt, "sqrt.approx.ftz.f32 %0, %0;")
OP(rsqrt, "rsqrt.approx.ftz.f32 %0, %0;")
OP(sin, "sin.approx.ftz.f32 %0, %0;")
OP(cos, "cos.approx.ftz.f32 %0, %0;")
OP(tanh, "tanh.approx.f32 %0, %0;")
OP(fma, "fma.rn.f32 %0, %0, %0, %0;")
template <typename K>
static void run(const char *name, K k, int per_iter_extra) {
    float *out; cudaMalloc(&out, 4);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096;
    k<<<sms * 4, 256>>>(out, 1.5f, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<sms * 4, 256>>>(out, 1.5f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double warp_insts_per_sm = 4.0 * 8 * 4 * iters;     // 4 CTAs x 8 warps x 4 ops x iters
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-6s %8.3f ms  %6.2f cycles per warp-instruction per SM  (%5.2f lanes/clk/SM)%s\n", name, ms, cycles / warp_insts_per_sm,
           32.0 * warp_insts_per_sm / cycles, per_iter_extra ? "  [includes the range-reduction FMUL]" : "");
    cudaFree(out);
}
int main() {
    run("ex2", k_ex2, 0); run("rcp", k_rcp, 0); run("lg2", k_lg2, 0); run("sqrt", k_sqrt, 0); run("rsqrt", k_rsqrt, 0);
    run("sin", k_sin, 1); run("cos", k_cos, 1); run("tanh", k_tanh, 0); run("fma", k_fma, 0);
    return 0;
}
