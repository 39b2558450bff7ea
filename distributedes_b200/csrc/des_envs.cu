// des_rollout_eval: closed-loop rollouts with per-member observations — the reference's actual
// workload: Evaluator.eval utils.py:116-124 -> single_run utils.py:126-139, once per member per repetition, with the
// environment stepped on the device.  Environment 0 = Pendulum-v0 (the reference's PendulumConfig, config.py:26-31:
// obs 3, action 1, clip +-2, 200-step episodes).
//
// One warp per member steps all of the member's episodes in lockstep (see rollout_pendulum_kernel); the member's
// perturbed weights theta + sigma*eps are generated once into shared memory with the same counter noise as every other
// kernel.  Policy arithmetic fp32 (FFMA, MUFU-based tanh), dynamics fp64 (gym keeps its state in float64).
// Algorithmic work per environment step: 2*(3H + H*H + H) flop; no HBM traffic beyond theta (L2 resident).
#include "des_common.cuh"

namespace des {

constexpr int kEnvPendulum = 0;
constexpr int kMaxRollH = 128;         // hidden units (4 per lane)
constexpr uint32_t kStreamEnvReset = 2u;
constexpr uint32_t kStreamActNoise = 3u;

struct RollArgs {
    float *fitness;                    // [n_local] mean return over the repetitions (higher is better)
    float *ep_ret;                     // optional [n_local][reps] per-episode returns
    double *stat_part;                 // optional [n_local][2*d0+1]: per-member sum, sum of squares, count of RAW observations
    const float *theta;
    const float *obs_stats;            // optional [m | v | n] (StaticNormalizer offline stats), NULL = identity
    const des_state *state;
    Layout L;
    int reps, horizon;
    float sigma, clip, act_noise;
    PhiloxKey key;
    uint32_t gen;
    uint64_t member_offset;
    uint32_t reset_member_base;        // counter word for the reset stream: member index (or the test-episode index)
    int noiseless;                     // 1: evaluate theta itself (test(), natural_es.py:101-110)
};

__device__ __forceinline__ double unit_open(uint32_t x) { return ((double)(x & 0x7FFFFFu) + 0.5) * (1.0 / 8388608.0); }

// gym Pendulum-v0 (gym/envs/classic_control/pendulum.py): g = 10, m = l = 1, dt = 0.05, max_speed 8, max_torque 2
struct Pendulum {
    double th, thdot, sn, cs;
    __device__ void reset(uint32_t rep, uint32_t member, uint32_t gen, const PhiloxKey &key) {
        const uint4 x = philox4x32(rep, member, gen, kStreamEnvReset, key);
        th = (2.0 * unit_open(x.x) - 1.0) * 3.141592653589793;        // uniform(-pi, pi)
        thdot = (2.0 * unit_open(x.y) - 1.0) * 1.0;                     // uniform(-1, 1)
    }
    // One fp64 sincos per step serves the observation and the torque term: sin(th + pi) = -sin(th).
    __device__ void observe(float *o) {
        sincos(th, &sn, &cs);
        o[0] = (float)cs;
        o[1] = (float)sn;
        o[2] = (float)thdot;
    }
    __device__ double step(double u) {                                  // returns the reward; observe() came first
        u = fmin(fmax(u, -2.0), 2.0);
        const double two_pi = 6.283185307179586, pi = 3.141592653589793;
        const double x = th + pi;
        const double an = (x - two_pi * floor(x * (1.0 / two_pi))) - pi;  // ((th + pi) % 2 pi) - pi, python sign rule
        const double cost = an * an + 0.1 * thdot * thdot + 0.001 * u * u;
        const double nthdot = thdot + (15.0 * sn + 3.0 * u) * 0.05;      // -3g/(2l) sin(th + pi) + 3/(m l^2) u
        th = th + nthdot * 0.05;
        thdot = fmin(fmax(nthdot, -8.0), 8.0);
        return -cost;
    }
};

// tanh as 1 - 2/(1 + 2^(2x log2 e)): two MUFU + three FP32 ops, abs error ~2e-7 (fp32 rounding level of the
// reference's torch.tanh).  40*H/32 of these per member-step make the libm tanhf a third of the instruction count.
__device__ __forceinline__ float tanh_mufu(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return __fmaf_rn(-2.0f, r, 1.0f);
}

constexpr int kEpPerLane = 5;          // episodes per lane half: 2 halves x 5 = up to 10 repetitions (config.py:8)
constexpr int kHS = 8;                 // row stride of an h1 panel (one panel per episode half): 5 episodes + pad

// One warp (= one CTA) per member; the warp steps all `reps` episodes of its member in lockstep, so the hidden layer is
// a [H x H] x [H x reps] product per step instead of `reps` mat-vecs: lane (rg = lane>>1, eg = lane&1) owns the
// R = H/16 hidden units rg*R.. and the episodes 5*eg..5*eg+4 — an R x 5 register tile fed per k by one LDS of R
// transposed weights and one 5-float broadcast of h1.  The fp64 dynamics of episode 5*eg + rg%5 run in every lane
// (the copies in rg >= 5 are redundant), i.e. once per step for the whole member.
template <int HPL>   // H = 32*HPL
__global__ void __launch_bounds__(32) rollout_pendulum_kernel(RollArgs a) {
    constexpr int H = 32 * HPL, R = 2 * HPL, C = kEpPerLane;
    extern __shared__ __align__(16) float sm[];
    float *W2T = sm;                     // [k][j] = W2[j][k]
    // h1 panels, one per episode half, rows in the permuted order p(k) = (k % R)*16 + k/R so that the 16 unit groups
    // storing their r-th unit hit consecutive 32-byte rows; the second panel is shifted by 16 bytes: conflict-free STS.128
    float *hT = W2T + H * H;             // [2][p(k)][kHS] (+4 floats)
    float *xs = hT + 2 * H * kHS + 8;    // [10][4] normalised observations
    float *W1s = xs + 40;                // [H][4]
    float *b1s = W1s + H * 4;            // [H]
    float *b2s = b1s + H;
    float *W3s = b2s + H;
    float *b3s = W3s + H;                // [4]
    double *red = reinterpret_cast<double *>(b3s + 4);      // [10][8] per-episode results
    const Layout L = a.L;
    const int lane = threadIdx.x, rg = lane >> 1, eg = lane & 1;
    const uint32_t gen = a.state ? (uint32_t)a.state->generation : a.gen;
    const uint32_t member = (uint32_t)(a.member_offset + blockIdx.x);

    // ---- theta' = theta + sigma*eps for this member -> shared memory (natural_es.py:28-30)
    for (int q = lane; q < (L.P + 3) / 4; q += 32) {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!a.noiseless) z = noise_quad((uint32_t)q, member, gen, kStreamNesEps, a.key);
        const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            if (j >= L.P) break;
            const float w = __fmaf_rn(a.sigma, zz[e], __ldg(a.theta + j));
            if (j < L.off_b1) W1s[(j / 3) * 4 + (j % 3)] = w;
            else if (j < L.off_w2) b1s[j - L.off_b1] = w;
            else if (j < L.off_b2) { const int r = (j - L.off_w2) / H; const int k = j - L.off_w2 - r * H; W2T[((k % R) * 16 + k / R) * H + r] = w; }
            else if (j < L.off_w3) b2s[j - L.off_b2] = w;
            else if (j < L.off_b3) W3s[j - L.off_w3] = w;
            else b3s[0] = w;
        }
    }
    __syncwarp();
    float w1[R][3], b1[R], b2[R], w3[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = rg * R + r;
        w1[r][0] = W1s[j * 4]; w1[r][1] = W1s[j * 4 + 1]; w1[r][2] = W1s[j * 4 + 2];
        b1[r] = b1s[j]; b2[r] = b2s[j]; w3[r] = W3s[j];
    }
    const float b3 = b3s[0];

    // StaticNormalizer (utils.py:48-51): identity while n == 0
    float nm[3] = {0.f, 0.f, 0.f}, ns[3] = {1.f, 1.f, 1.f};
    if (a.obs_stats && a.obs_stats[6] != 0.f) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { nm[k] = a.obs_stats[k]; ns[k] = sqrtf(a.obs_stats[3 + k] + 1e-6f); }
    }
    float *hp = hT + eg * (H * kHS + 4);                  // this lane's h1 panel
    const int csel = rg % C, ep = C * eg + csel;         // the episode whose dynamics this lane carries
    const bool writer = rg < C;                           // one lane per episode publishes
    Pendulum env;
    env.reset((uint32_t)ep, a.reset_member_base + (a.noiseless ? 0u : (uint32_t)blockIdx.x), gen, a.key);
    double total = 0.0, osum[3] = {0, 0, 0}, osq[3] = {0, 0, 0};
    for (int t = 0; t < a.horizon; ++t) {
        {
            float o[3];
            env.observe(o);
#pragma unroll
            for (int k = 0; k < 3; ++k) { osum[k] += (double)o[k]; osq[k] += (double)o[k] * (double)o[k]; }
            if (writer)
                *reinterpret_cast<float4 *>(xs + ep * 4) =
                    make_float4((o[0] - nm[0]) / ns[0], (o[1] - nm[1]) / ns[1], (o[2] - nm[2]) / ns[2], 0.f);
        }
        __syncwarp();
        // layer 1: the lane's R units x 5 episodes -> h1 panel
        {
            float4 x[C];
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = *reinterpret_cast<const float4 *>(xs + (C * eg + c) * 4);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float v[C];
#pragma unroll
                for (int c = 0; c < C; ++c)
                    v[c] = tanh_mufu(__fmaf_rn(w1[r][2], x[c].z, __fmaf_rn(w1[r][1], x[c].y, __fmaf_rn(w1[r][0], x[c].x, b1[r]))));
                float *dst = hp + (r * 16 + rg) * kHS;
                *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                dst[4] = v[4];
            }
        }
        __syncwarp();
        // layer 2: R x 5 register tile
        float acc[R][C];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[r][c] = b2[r];
#pragma unroll 16
        for (int k = 0; k < H; ++k) {
            float w[R];
            if constexpr (R % 4 == 0) {
#pragma unroll
                for (int r4 = 0; r4 < R / 4; ++r4) {
                    const float4 ww = *reinterpret_cast<const float4 *>(W2T + k * H + rg * R + 4 * r4);
                    w[4 * r4] = ww.x; w[4 * r4 + 1] = ww.y; w[4 * r4 + 2] = ww.z; w[4 * r4 + 3] = ww.w;
                }
            } else {
#pragma unroll
                for (int r2 = 0; r2 < R / 2; ++r2) {
                    const float2 ww = *reinterpret_cast<const float2 *>(W2T + k * H + rg * R + 2 * r2);
                    w[2 * r2] = ww.x; w[2 * r2 + 1] = ww.y;
                }
            }
            const float4 h4 = *reinterpret_cast<const float4 *>(hp + k * kHS);      // k runs in the permuted order
            const float h5 = hp[k * kHS + 4];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc[r][0] = __fmaf_rn(w[r], h4.x, acc[r][0]);
                acc[r][1] = __fmaf_rn(w[r], h4.y, acc[r][1]);
                acc[r][2] = __fmaf_rn(w[r], h4.z, acc[r][2]);
                acc[r][3] = __fmaf_rn(w[r], h4.w, acc[r][3]);
                acc[r][4] = __fmaf_rn(w[r], h5, acc[r][4]);
            }
        }
        // layer 3 (one action per episode): partial over the lane's units, butterfly over the 16 unit groups
        float p[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) s = __fmaf_rn(w3[r], tanh_mufu(acc[r][c]), s);
            p[c] = s;
        }
#pragma unroll
        for (int off = 2; off < 32; off <<= 1)
#pragma unroll
            for (int c = 0; c < C; ++c) p[c] += __shfl_xor_sync(0xffffffffu, p[c], off);
        float act = p[0];
#pragma unroll
        for (int c = 1; c < C; ++c) act = (csel == c) ? p[c] : act;
        act += b3;
        if (a.act_noise != 0.f) {                                        // utils.py:133
            const uint4 xr = philox4x32((uint32_t)t, member * 16u + (uint32_t)ep, gen, kStreamActNoise, a.key);
            float z0, z1;
            box_muller(xr.x, xr.y, z0, z1);
            act = __fmaf_rn(z0, a.act_noise, act);
        }
        act = fminf(fmaxf(act, -a.clip), a.clip);                        // config.action_clip, utils.py:134
        total += env.step((double)act);                                  // utils.py:135-137
    }
    if (writer) {
        red[ep * 8] = total;
#pragma unroll
        for (int k = 0; k < 3; ++k) { red[ep * 8 + 1 + k] = osum[k]; red[ep * 8 + 4 + k] = osq[k]; }
    }
    __syncwarp();
    if (lane == 0) {
        double s = 0.0;
        for (int r = 0; r < a.reps; ++r) s += red[r * 8];
        a.fitness[blockIdx.x] = (float)(s / a.reps);                     // -cost of utils.py:124
    }
    if (a.ep_ret && lane < a.reps) a.ep_ret[(int64_t)blockIdx.x * a.reps + lane] = (float)red[lane * 8];
    if (a.stat_part) {      // raw observations fed to the normaliser by this member, episodes summed in a fixed order
        if (lane < 6) {
            double s = 0.0;
            for (int r = 0; r < a.reps; ++r) s += red[r * 8 + 1 + lane];
            a.stat_part[(int64_t)blockIdx.x * 7 + lane] = s;
        }
        if (lane == 6) a.stat_part[(int64_t)blockIdx.x * 7 + 6] = (double)a.reps * a.horizon;
    }
}

// Chan-merge the per-member observation partials (all members of all ranks, after an all-reduce of the three sums) into
// the shared statistics (utils.py:85-96).  totals = [sum(3) | sumsq(3) | count] in fp64.
__global__ void obs_stats_merge_totals_kernel(float *__restrict__ stats, const double *__restrict__ totals, int d0) {
    const int k = threadIdx.x;
    if (k >= d0) return;
    const double nB = totals[2 * d0];
    if (nB <= 0) return;
    const double mb = totals[k] / nB, vb = fmax(totals[d0 + k] / nB - mb * mb, 0.0);
    const double nA = (double)stats[2 * d0], n = nA + nB;
    const double mA = (double)stats[k], vA = (double)stats[d0 + k];
    const double delta = mb - mA;
    const double m = mA + delta * nB / n;
    const double v = (vA * nA + vb * nB + delta * delta * nA * nB / n) / n;
    __syncthreads();
    stats[k] = (float)m;
    stats[d0 + k] = (float)v;
    if (k == 0) stats[2 * d0] = (float)n;
}

__global__ void stat_part_reduce_kernel(double *__restrict__ totals, const double *__restrict__ part, int64_t n_local, int width) {
    // one thread per column, fixed order over members: deterministic
    const int c = threadIdx.x;
    if (c >= width) return;
    double s = 0.0;
    for (int64_t i = 0; i < n_local; ++i) s += part[i * width + c];
    totals[c] = s;
}

}  // namespace des

extern "C" DES_API int des_rollout_eval(float *fitness_out_dev, float *episode_returns_out_dev,
                                        double *obs_totals_out_dev, const float *theta_dev,
                                        const float *obs_stats_dev, int env, des_dims dims, int32_t repetitions,
                                        double sigma, double clip, double action_noise_std, uint64_t seed,
                                        uint64_t generation, const des_state *state_dev, int64_t member_offset,
                                        int64_t n_local, int noiseless, void *workspace_dev, size_t workspace_bytes,
                                        void *stream) {
    using namespace des;
    DES_REQUIRE(env == kEnvPendulum, "des_rollout_eval: unknown environment %d (0 = Pendulum-v0)", env);
    DES_REQUIRE(dims.state_dim == 3 && dims.action_dim == 1, "des_rollout_eval: Pendulum-v0 has state_dim 3, action_dim 1");
    DES_REQUIRE(dims.hidden > 0 && dims.hidden % 32 == 0 && dims.hidden <= kMaxRollH,
                "des_rollout_eval: hidden must be a multiple of 32, <= %d (got %d)", kMaxRollH, dims.hidden);
    DES_REQUIRE(repetitions >= 1 && repetitions <= 10, "des_rollout_eval: repetitions must be in [1, 10] (one warp each)");
    DES_REQUIRE(dims.tape_len >= 1, "des_rollout_eval: episode length (dims.tape_len) must be >= 1");
    DES_REQUIRE(n_local >= 0 && member_offset >= 0 && member_offset + n_local <= (int64_t)1 << 28,
                "des_rollout_eval: bad member range");
    if (n_local == 0) return DES_OK;
    DES_REQUIRE(fitness_out_dev && theta_dev, "des_rollout_eval: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    RollArgs a;
    a.fitness = fitness_out_dev; a.ep_ret = episode_returns_out_dev; a.theta = theta_dev; a.obs_stats = obs_stats_dev; a.state = state_dev;
    a.L = Layout(3, dims.hidden, 1);
    a.reps = repetitions; a.horizon = dims.tape_len;
    a.sigma = noiseless ? 0.f : (float)sigma; a.clip = (float)clip; a.act_noise = (float)action_noise_std;
    a.key = make_philox_key(seed); a.gen = (uint32_t)generation;
    a.member_offset = (uint64_t)member_offset;
    a.reset_member_base = noiseless ? 0x40000000u : (uint32_t)member_offset;     // test episodes use their own reset stream
    a.noiseless = noiseless ? 1 : 0;
    a.stat_part = nullptr;
    if (obs_totals_out_dev) {
        const size_t need = (size_t)n_local * 7 * sizeof(double);
        if (!workspace_dev || workspace_bytes < need) {
            set_error("des_rollout_eval: workspace %zu B < required %zu B", workspace_bytes, need);
            return DES_ERR_WORKSPACE;
        }
        a.stat_part = (double *)workspace_dev;
    }
    const int H = dims.hidden;
    const size_t smem = sizeof(float) * ((size_t)H * H + 2 * (size_t)H * kHS + 8 + 40 + (size_t)H * 4 + 3 * (size_t)H + 4) +
                        sizeof(double) * 80;
#define DES_ROLL_LAUNCH(HPL)                                                                                        \
    do {                                                                                                            \
        DES_CUDA(cudaFuncSetAttribute(rollout_pendulum_kernel<HPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        rollout_pendulum_kernel<HPL><<<(unsigned)n_local, 32, smem, st>>>(a);                                  \
    } while (0)
    switch (H / 32) {
        case 1: DES_ROLL_LAUNCH(1); break;
        case 2: DES_ROLL_LAUNCH(2); break;
        case 3: DES_ROLL_LAUNCH(3); break;
        default: DES_ROLL_LAUNCH(4); break;
    }
#undef DES_ROLL_LAUNCH
    DES_LAUNCH_CHECK("rollout_pendulum_kernel");
    if (obs_totals_out_dev) {
        stat_part_reduce_kernel<<<1, 32, 0, st>>>(obs_totals_out_dev, a.stat_part, n_local, 7);
        DES_LAUNCH_CHECK("stat_part_reduce_kernel");
    }
    return DES_OK;
}

extern "C" DES_API int des_obs_stats_merge_totals(float *stats_dev, const double *obs_totals_dev, int32_t state_dim,
                                                  void *stream) {
    DES_REQUIRE(stats_dev && obs_totals_dev && state_dim > 0 && state_dim <= 1024, "des_obs_stats_merge_totals: bad arguments");
    des::obs_stats_merge_totals_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(stats_dev, obs_totals_dev, state_dim);
    DES_LAUNCH_CHECK("obs_stats_merge_totals_kernel");
    return DES_OK;
}
