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

// ---- Philox4x32-R, R = kPhiloxRounds ---------------------------------------------------------------
// Round 2 moved the noise contract from 10 to 7 rounds: Philox4x32-7 is the smallest round count Salmon et al.
// (SC'11, table 2) report as passing BigCrush; the generator is paid for twice per generation (forward and
// fitness x noise reduction), where the ten-round multiplies were the busiest pipe (profiles/README.md §3).
// oracle/nes_oracle.py, the goldens and include/des_b200.h moved in the same commit.
constexpr int kPhiloxRounds = 7;
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kStreamNesEps = 0u;
constexpr uint32_t kStreamCmaZ = 1u;

// Round keys k + r*W precomputed on the host (kernel-parameter constant bank): the xor takes them as
// constant operands, so the key schedule costs no instructions.
struct PhiloxKey {
    uint32_t k0[kPhiloxRounds], k1[kPhiloxRounds];
    uint32_t one_bits;          // 0x3F800000, as a kernel parameter: see u32_to_one_two(x, one)
};
__host__ inline PhiloxKey make_philox_key(uint64_t seed) {
    PhiloxKey k;
    for (int r = 0; r < kPhiloxRounds; ++r) {
        k.k0[r] = (uint32_t)seed + (uint32_t)r * kPhiloxW0;
        k.k1[r] = (uint32_t)(seed >> 32) + (uint32_t)r * kPhiloxW1;
    }
    k.one_bits = 0x3F800000u;
    return k;
}

__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKey &key) {
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ key.k0[r];
        const uint32_t n2 = hi0 ^ c3 ^ key.k1[r];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// uint32 -> f = 1 + k*2^-23 in [1,2) from the LOW 23 bits k of the word: one LOP3, no int->float conversion
// (I2F issues on the 16-lane XU pipe that the Box-Muller MUFUs already load).  The uniform is
// u = f - (1 - 2^-24) = (2k+1)*2^-24, on the open interval (0,1).
__device__ __forceinline__ float u32_to_one_two(uint32_t x) {
    return __uint_as_float(0x3F800000u | (x & 0x007FFFFFu));
}

// The same in ONE instruction: with both constants immediate ptxas emits two LOP3 (and, or) — an instruction takes one
// immediate.  With 0x3F800000 in a register (`one`, read from the kernel parameters so that it stays a register operand)
// it is LOP3 d = (x & 0x7FFFFF) | one.  8 issue slots less per octet of deviates in the generator loops.
__device__ __forceinline__ float u32_to_one_two(uint32_t x, uint32_t one) {
    uint32_t r;
    asm("lop3.b32 %0, %1, 0x007FFFFF, %2, 0xEA;" : "=r"(r) : "r"(x), "r"(one));
    return __uint_as_float(r);
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

constexpr float kTwoPiF = 6.283185307179586f;             // fl32(2*pi)
constexpr float kAngOffF = 9.424777586262351f;            // fl32(3*pi - pi*2^-23)

// Box-Muller.  With f1, f2 in [1,2) from the two words:
//   u1  = f1 - (1 - 2^-24)                     (exact)
//   ang = fl32(f2*fl32(2*pi) - fl32(3*pi - pi*2^-23))   one FFMA;  ang ~= 2*pi*u2 - pi in (-pi, pi),
//         where the MUFU sin/cos error bound (2^-21.4 abs) holds
//   z0  = -sqrt(-2 ln u1) * cos(ang),  z1 = -sqrt(-2 ln u1) * sin(ang)       (cos(t+pi) = -cos t)
// oracle/nes_oracle.py restates u1 and ang bit-exactly and evaluates ln/sqrt/sin/cos in fp64.
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float &z0, float &z1) {
    const float u1 = u32_to_one_two(xa) - 0.99999994039535522f;
    const float nr = -sqrt_approx(-1.3862943611198906f * lg2_approx(u1));     // -2 ln u = (-2 ln 2) lg2 u
    const float ang = __fmaf_rn(u32_to_one_two(xb), kTwoPiF, -kAngOffF);
    z0 = nr * cos_approx(ang);
    z1 = nr * sin_approx(ang);
}

// Box-Muller in parts, for consumers that fold a scale into the radius: the pair is (nr*c, nr*s) with
// nr = -sqrt(scale2 * -2 ln u1) = -scale*sqrt(-2 ln u1) for scale2 = scale^2 (pass kNeg2Ln2 * scale^2).
constexpr float kNeg2Ln2 = -1.3862943611198906f;
struct BmParts {
    float nr, c, s;
};
__device__ __forceinline__ BmParts box_muller_parts(uint32_t xa, uint32_t xb, float neg2ln2_scale2) {
    BmParts p;
    const float u1 = u32_to_one_two(xa) - 0.99999994039535522f;
    p.nr = -sqrt_approx(neg2ln2_scale2 * lg2_approx(u1));
    const float ang = __fmaf_rn(u32_to_one_two(xb), kTwoPiF, -kAngOffF);
    p.c = cos_approx(ang);
    p.s = sin_approx(ang);
    return p;
}
// hot-loop form: `one` = PhiloxKey::one_bits (identical values)
__device__ __forceinline__ BmParts box_muller_parts(uint32_t xa, uint32_t xb, float neg2ln2_scale2, uint32_t one) {
    BmParts p;
    const float u1 = u32_to_one_two(xa, one) - 0.99999994039535522f;
    p.nr = -sqrt_approx(neg2ln2_scale2 * lg2_approx(u1));
    const float ang = __fmaf_rn(u32_to_one_two(xb, one), kTwoPiF, -kAngOffF);
    p.c = cos_approx(ang);
    p.s = sin_approx(ang);
    return p;
}
// out[e] = base[e] + scale * eps[e] for the four parameters of quad q (scale folded into the radius:
// differs from fma(scale, eps, base) by <= 1 ulp of scale*eps).
__device__ __forceinline__ float4 perturbed_quad(uint32_t q, uint32_t member, uint32_t gen, uint32_t tag,
                                                 const PhiloxKey &key, float neg2ln2_scale2, float4 base) {
    const uint4 x = philox4x32(q, member, gen, tag, key);
    const BmParts a = box_muller_parts(x.x, x.y, neg2ln2_scale2, key.one_bits);
    const BmParts b = box_muller_parts(x.z, x.w, neg2ln2_scale2, key.one_bits);
    return make_float4(__fmaf_rn(a.nr, a.c, base.x), __fmaf_rn(a.nr, a.s, base.y), __fmaf_rn(b.nr, b.c, base.z),
                       __fmaf_rn(b.nr, b.s, base.w));
}

// The four normals of quad q of `member` at `gen`.
__device__ __forceinline__ float4 noise_quad(uint32_t q, uint32_t member, uint32_t gen, uint32_t tag,
                                             const PhiloxKey &key) {
    const uint4 x = philox4x32(q, member, gen, tag, key);
    float4 z;
    box_muller(x.x, x.y, z.x, z.y);
    box_muller(x.z, x.w, z.z, z.w);
    return z;
}

__device__ __forceinline__ uint64_t load_generation(const des_state *st, uint64_t fallback) {
    return st ? st->generation : fallback;
}

}  // namespace des
