// des_rollout_eval: closed-loop rollouts with per-member observations — the reference's actual
// workload: Evaluator.eval utils.py:116-124 -> single_run utils.py:126-139, once per member per repetition, with the
// environment stepped on the device.  Environment 0 = Pendulum-v0 (the reference's PendulumConfig, config.py:26-31:
// obs 3, action 1, clip +-2, 200-step episodes).
//
// One CTA per member, one warp per episode (repetition).  The member's perturbed weights theta + sigma*eps are generated
// once into shared memory (same counter noise as every other kernel); each warp then runs its episode sequentially:
// lane l owns hidden units l, l+32, ..., the 64x64 layer is a shuffle-broadcast mat-vec, the scalar dynamics are
// computed redundantly by all lanes in fp64 (gym keeps the state in float64).  Per-step work is ~2*(3H + H*H + H) flop:
// the kernel is latency/issue bound over N*R independent episodes, not a GEMM.
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
    int reps, horizon, S2;             // S2: padded row stride of W2 in shared memory
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
    double th, thdot;
    __device__ void reset(uint32_t rep, uint32_t member, uint32_t gen, const PhiloxKey &key) {
        const uint4 x = philox4x32_10(rep, member, gen, kStreamEnvReset, key);
        th = (2.0 * unit_open(x.x) - 1.0) * 3.141592653589793;        // uniform(-pi, pi)
        thdot = (2.0 * unit_open(x.y) - 1.0) * 1.0;                     // uniform(-1, 1)
    }
    __device__ void observe(float *o) const {
        o[0] = (float)cos(th);
        o[1] = (float)sin(th);
        o[2] = (float)thdot;
    }
    __device__ double step(double u) {                                  // returns the reward
        u = fmin(fmax(u, -2.0), 2.0);
        const double two_pi = 6.283185307179586;
        double an = fmod(th + 3.141592653589793, two_pi);               // python %: result has the sign of the divisor
        if (an < 0) an += two_pi;
        an -= 3.141592653589793;
        const double cost = an * an + 0.1 * thdot * thdot + 0.001 * u * u;
        double nthdot = thdot + (-3.0 * 10.0 / 2.0 * sin(th + 3.141592653589793) + 3.0 * u) * 0.05;
        th = th + nthdot * 0.05;
        thdot = fmin(fmax(nthdot, -8.0), 8.0);
        return -cost;
    }
};

template <int HPL>   // hidden units per lane (H = 32*HPL)
__global__ void __launch_bounds__(320) rollout_pendulum_kernel(RollArgs a) {
    extern __shared__ __align__(16) float sm[];
    const Layout L = a.L;
    const int H = L.H, S2 = a.S2;
    float *W1 = sm;                      // [H][4]  (3 inputs + pad)
    float *b1 = W1 + H * 4;
    float *W2 = b1 + H;                  // [H][S2]
    float *b2 = W2 + H * S2;
    float *W3 = b2 + H;                  // [H]   (action_dim == 1)
    float *b3 = W3 + H;                  // [1]
    __shared__ double ret[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gen = a.state ? (uint32_t)a.state->generation : a.gen;
    const uint32_t member = (uint32_t)(a.member_offset + blockIdx.x);

    // ---- theta' = theta + sigma*eps for this member -> shared memory (natural_es.py:28-30)
    for (int q = threadIdx.x; q < (L.P + 3) / 4; q += blockDim.x) {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!a.noiseless) z = noise_quad((uint32_t)q, member, gen, kStreamNesEps, a.key);
        const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            if (j >= L.P) break;
            const float w = __fmaf_rn(a.sigma, zz[e], __ldg(a.theta + j));
            if (j < L.off_b1) W1[(j / 3) * 4 + (j % 3)] = w;
            else if (j < L.off_w2) b1[j - L.off_b1] = w;
            else if (j < L.off_b2) { const int r = (j - L.off_w2) / H; W2[r * S2 + (j - L.off_w2 - r * H)] = w; }
            else if (j < L.off_w3) b2[j - L.off_b2] = w;
            else if (j < L.off_b3) W3[j - L.off_w3] = w;
            else b3[0] = w;
        }
    }
    __syncthreads();

    // StaticNormalizer (utils.py:48-51): identity while n == 0
    float nm[3] = {0.f, 0.f, 0.f}, ns[3] = {1.f, 1.f, 1.f};
    if (a.obs_stats && a.obs_stats[6] != 0.f) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { nm[k] = a.obs_stats[k]; ns[k] = sqrtf(a.obs_stats[3 + k] + 1e-6f); }
    }
    double total = 0.0, osum[3] = {0, 0, 0}, osq[3] = {0, 0, 0};
    if (warp < a.reps) {
        Pendulum env;
        env.reset((uint32_t)warp, a.reset_member_base + (a.noiseless ? 0u : (uint32_t)blockIdx.x), gen, a.key);
        for (int t = 0; t < a.horizon; ++t) {
            float o[3];
            env.observe(o);
#pragma unroll
            for (int k = 0; k < 3; ++k) { osum[k] += (double)o[k]; osq[k] += (double)o[k] * (double)o[k]; }
            float x[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) x[k] = (o[k] - nm[k]) / ns[k];
            // layer 1: lane owns hidden units lane + 32*i
            float h1[HPL];
#pragma unroll
            for (int i = 0; i < HPL; ++i) {
                const int j = lane + 32 * i;
                const float4 w = *reinterpret_cast<const float4 *>(W1 + j * 4);
                h1[i] = tanhf(__fmaf_rn(w.z, x[2], __fmaf_rn(w.y, x[1], __fmaf_rn(w.x, x[0], b1[j]))));
            }
            // layer 2: mat-vec, h1 broadcast by shuffles, W2 rows from shared memory (padded stride: conflict free)
            float acc[HPL];
#pragma unroll
            for (int i = 0; i < HPL; ++i) acc[i] = b2[lane + 32 * i];
#pragma unroll
            for (int ki = 0; ki < HPL; ++ki) {
#pragma unroll 8
                for (int kl = 0; kl < 32; ++kl) {
                    const float hk = __shfl_sync(0xffffffffu, h1[ki], kl);
                    const int k = kl + 32 * ki;
#pragma unroll
                    for (int i = 0; i < HPL; ++i) acc[i] = __fmaf_rn(W2[(lane + 32 * i) * S2 + k], hk, acc[i]);
                }
            }
            // layer 3 (one action): warp reduction in a fixed order
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < HPL; ++i) part = __fmaf_rn(W3[lane + 32 * i], tanhf(acc[i]), part);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
            float act = part + b3[0];
            if (a.act_noise != 0.f) {                                    // utils.py:133
                const uint4 xr = philox4x32_10((uint32_t)t, member * 16u + (uint32_t)warp, gen, kStreamActNoise, a.key);
                float z0, z1;
                box_muller(xr.x, xr.y, z0, z1);
                act = __fmaf_rn(z0, a.act_noise, act);
            }
            act = fminf(fmaxf(act, -a.clip), a.clip);                    // config.action_clip, utils.py:134
            total += env.step((double)act);                              // utils.py:135-137
        }
    }
    if (lane == 0 && warp < 16) ret[warp] = (warp < a.reps) ? total : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int r = 0; r < a.reps; ++r) s += ret[r];
        a.fitness[blockIdx.x] = (float)(s / a.reps);                     // -cost of utils.py:124
    }
    if (a.ep_ret && threadIdx.x < a.reps) a.ep_ret[(int64_t)blockIdx.x * a.reps + threadIdx.x] = (float)ret[threadIdx.x];
    if (a.stat_part) {      // raw observations seen by this member: per-episode partials combined in a fixed order
        __shared__ double sp[16][6];
        if (lane == 0 && warp < 16)
            for (int k = 0; k < 3; ++k) { sp[warp][k] = (warp < a.reps) ? osum[k] : 0.0; sp[warp][3 + k] = (warp < a.reps) ? osq[k] : 0.0; }
        __syncthreads();
        if (threadIdx.x < 6) {
            double s = 0.0;
            for (int r = 0; r < a.reps; ++r) s += sp[r][threadIdx.x];
            a.stat_part[(int64_t)blockIdx.x * 7 + threadIdx.x] = s;
        }
        if (threadIdx.x == 6) a.stat_part[(int64_t)blockIdx.x * 7 + 6] = (double)a.reps * a.horizon;
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
    a.S2 = dims.hidden + 1;
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
    const size_t smem = sizeof(float) * ((size_t)H * 4 + H + (size_t)H * a.S2 + H + H + 4);
    const int threads = 32 * repetitions;
#define DES_ROLL_LAUNCH(HPL)                                                                                        \
    do {                                                                                                            \
        DES_CUDA(cudaFuncSetAttribute(rollout_pendulum_kernel<HPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        rollout_pendulum_kernel<HPL><<<(unsigned)n_local, threads, smem, st>>>(a);                                  \
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
