// Shared device helpers: Philox4x32-10 counter RNG, Box-Muller, error plumbing.
// Target: sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/des_b200.h"

namespace des {

// ---- error plumbing (host) ---------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define DES_REQUIRE(cond, ...)                    \
    do {                                          \
        if (!(cond)) {                            \
            ::des::set_error(__VA_ARGS__);        \
            return DES_ERR_INVALID_ARGUMENT;      \
        }                                         \
    } while (0)

#define DES_CUDA(call)                                            \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return ::des::cuda_fail(e__, #call); \
    } while (0)

#define DES_LAUNCH_CHECK(name)                                    \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return ::des::cuda_fail(e__, name); \
    } while (0)

// ---- layout of the flat parameter vector (model.py:8-25, 30-32) -----------------------------------
struct Layout {
    int d0, H, A;
    int off_w1, off_b1, off_w2, off_b2, off_w3, off_b3, P;
    __host__ __device__ Layout() {}
    __host__ __device__ Layout(int d0_, int H_, int A_) : d0(d0_), H(H_), A(A_) {
        off_w1 = 0;
        off_b1 = off_w1 + H * d0;
        off_w2 = off_b1 + H;
        off_b2 = off_w2 + H * H;
        off_w3 = off_b2 + H;
        off_b3 = off_w3 + A * H;
        P = off_b3 + A;
    }
};

// ---- Philox4x32-10 -------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kStreamNesEps = 0u;
constexpr uint32_t kStreamCmaZ = 1u;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ (k0 + (uint32_t)r * kPhiloxW0);
        const uint32_t n2 = hi0 ^ c3 ^ (k1 + (uint32_t)r * kPhiloxW1);
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// uint32 -> fp32 uniform in (0, 1]: one fused rounding of float(x)*2^-32 + 2^-33.
__device__ __forceinline__ float u32_to_unit(uint32_t x) {
    return __fmaf_rn(__uint2float_rn(x), 0x1p-32f, 0x1p-33f);
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cos_approx(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Box-Muller: (z0, z1) = sqrt(-2 ln u1) * (cos 2*pi*u2, sin 2*pi*u2).
// The angle is shifted into (-pi, pi] where the MUFU sin/cos error bound (2^-21.4 abs) holds:
// cos(2*pi*u) = -cos(2*pi*u - pi), sin likewise, so the sign is folded into r.
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float &z0, float &z1) {
    const float u1 = u32_to_unit(xa);
    const float u2 = u32_to_unit(xb);
    // -2 ln u1 = (-2 ln 2) * lg2(u1)
    const float nr = -sqrt_approx(-1.3862943611198906f * lg2_approx(u1));
    const float ang = __fmaf_rn(u2, 6.283185307179586f, -3.141592653589793f);
    z0 = nr * cos_approx(ang);
    z1 = nr * sin_approx(ang);
}

// The four normals of quad q of `member` at `gen`.
__device__ __forceinline__ float4 noise_quad(uint32_t q, uint32_t member, uint32_t gen, uint32_t tag,
                                             uint32_t k0, uint32_t k1) {
    const uint4 x = philox4x32_10(q, member, gen, tag, k0, k1);
    float4 z;
    box_muller(x.x, x.y, z.x, z.y);
    box_muller(x.z, x.w, z.z, z.w);
    return z;
}

__device__ __forceinline__ uint64_t load_generation(const des_state *st, uint64_t fallback) {
    return st ? st->generation : fallback;
}

}  // namespace des
