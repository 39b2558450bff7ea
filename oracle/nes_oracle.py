"""CPU oracle for the NES population-evaluate-update hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``distributedes_b200``) never does; it fails
loudly when the CUDA library is missing.

It is a numpy restatement of the reference's algorithm (ShangtongZhang/DistributedES @ c4de970).
Every function cites the reference ``file:line`` it follows.  Arithmetic that the reference does in
fp64 (noise, rank shaping, gradient estimate, Adam) is fp64 here; the policy forward is fp64 on
fp32 weights (the reference runs it in fp32 torch; fp64 is the tighter check).

Parity pinning: the reference ships no golden vectors, so the oracle is pinned
against the reference's *own functions* imported from ``/root/reference`` by
``oracle/make_golden.py`` (fixtures committed under ``tests/golden/``) and checked by
``tests/test_oracle_golden.py``.  The counter-based RNG is ours (the reference seeds MT19937 from
OS entropy, ``natural_es.py:23``, so it has no reproducible stream); it is pinned to the published
Random123 Philox4x32 known-answer vectors (the product uses 7 rounds).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------
# Counter-based noise:  Philox4x32-7  +  Box-Muller
# --------------------------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_SH32 = np.uint64(32)

STREAM_NES_EPS = 0  # counter word c3 for NES perturbations
STREAM_CMA_Z = 1    # counter word c3 for CMA-ES z samples


PHILOX_ROUNDS = 7   # the noise contract (include/des_b200.h): Philox4x32-7, the smallest BigCrush-passing round count


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=PHILOX_ROUNDS):
    """Philox4x32 with `rounds` rounds (Salmon et al., SC'11; Random123 reference constants).  The same loop is
    pinned to the Random123 known-answer vectors at 7 and 10 rounds (tests/test_oracle_golden.py).

    All arguments broadcastable integer arrays holding uint32 values; returns four uint32 arrays.
    """
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK32
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for _ in range(rounds):
        p0 = PHILOX_M0 * c0            # < 2^64, exact in uint64
        p1 = PHILOX_M1 * c2
        n0 = (p1 >> _SH32) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK32
        n2 = (p0 >> _SH32) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


_TWO_PI_F = float(np.float32(6.283185307179586))              # fl32(2*pi)
_ANG_OFF_F = float(np.float32(9.424777586262351))             # fl32(3*pi - pi*2^-23)
_ONE_M = float(np.float32(1.0 - 2.0 ** -24))                  # fl32(1 - 2^-24)


def u32_to_one_two(x):
    """f = 1 + k*2^-23 in [1,2) from the LOW 23 bits k of the word (exact; the CUDA side is one LOP3:
    as_float(0x3F800000 | (x & 0x7FFFFF))).  Returned as float64."""
    k = (np.asarray(x, dtype=np.uint32) & np.uint32(0x7FFFFF)).astype(np.float64)
    return 1.0 + k * 2.0 ** -23


def box_muller(xa, xb):
    """The noise contract of include/des_b200.h, fp32 intermediates restated bit-exactly:
         u1  = f1 - (1 - 2^-24)                      = (2k1+1)*2^-24 in (0,1), exact in fp32
         ang = fl32(f2*fl32(2 pi) - fl32(3 pi - pi 2^-23))   one fused rounding; ~ 2 pi u2 - pi in (-pi, pi)
         z0  = -sqrt(-2 ln u1) cos(ang),  z1 = -sqrt(-2 ln u1) sin(ang)      (ln/sqrt/sin/cos in fp64 here)"""
    u1 = u32_to_one_two(xa) - _ONE_M
    ang = (u32_to_one_two(xb) * _TWO_PI_F - _ANG_OFF_F).astype(np.float32).astype(np.float64)
    r = -np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(ang), r * np.sin(ang)


def noise_uint32(seed, gen, member, n_quads, stream=STREAM_NES_EPS):
    """Raw Philox words for one member: counter = (quad, member, gen, stream), key = seed."""
    q = np.arange(n_quads, dtype=np.uint64)
    return philox4x32(q, np.uint64(member), np.uint64(gen & 0xFFFFFFFF), np.uint64(stream),
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def noise(seed, gen, member_offset, n_members, P, stream=STREAM_NES_EPS):
    """eps[n_members, P] fp64.  eps[i, 4q:4q+4] = BM(x0,x1) ++ BM(x2,x3) of Philox(q, member, gen).

    Stands in for ``epsilon = np.random.randn(len(disturbed_param))`` natural_es.py:29 — a pure
    function of (seed, generation, global member index, parameter index), so shards regenerate it
    instead of shipping it (natural_es.py:32 pickles it through a pipe).
    """
    nq = (P + 3) // 4
    q = np.arange(nq, dtype=np.uint64)[None, :]
    m = (np.arange(n_members, dtype=np.uint64) + np.uint64(member_offset))[:, None]
    x0, x1, x2, x3 = philox4x32(q, m, np.uint64(gen & 0xFFFFFFFF), np.uint64(stream),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, z1 = box_muller(x0, x1)
    z2, z3 = box_muller(x2, x3)
    eps = np.stack([z0, z1, z2, z3], axis=-1).reshape(n_members, nq * 4)
    return np.ascontiguousarray(eps[:, :P])


# --------------------------------------------------------------------------------------------
# Flat-weight codec and policy forward
# --------------------------------------------------------------------------------------------
def param_count(d0, H, A):
    return d0 * H + H + H * H + H + H * A + A


def unflatten(flat, d0, H, A):
    """Flat layout of BaseModel.get_weight/set_weight, model.py:8-25 with StandardFCNet model.py:30-32:
    [fc1.weight (H x d0, row-major) | fc1.bias (H) | fc2.weight (H x H) | fc2.bias | fc3.weight (A x H) | fc3.bias]
    Leading batch dims are preserved."""
    flat = np.asarray(flat)
    lead = flat.shape[:-1]
    assert flat.shape[-1] == param_count(d0, H, A)   # model.py:25 asserts full consumption
    o = 0
    W1 = flat[..., o:o + H * d0].reshape(*lead, H, d0); o += H * d0
    b1 = flat[..., o:o + H]; o += H
    W2 = flat[..., o:o + H * H].reshape(*lead, H, H); o += H * H
    b2 = flat[..., o:o + H]; o += H
    W3 = flat[..., o:o + A * H].reshape(*lead, A, H); o += A * H
    b3 = flat[..., o:o + A]; o += A
    return W1, b1, W2, b2, W3, b3


def forward(flat, obs, d0, H, A, dtype=np.float64):
    """StandardFCNet.forward model.py:34-39:  a = W3 tanh(W2 tanh(W1 x + b1) + b2) + b3.

    flat: [..., P]; obs: [T, d0]; returns actions [..., T, A] in ``dtype``."""
    W1, b1, W2, b2, W3, b3 = (w.astype(dtype) for w in unflatten(flat, d0, H, A))
    x = np.asarray(obs, dtype=dtype)
    h = np.tanh(np.einsum('td,...hd->...th', x, W1) + b1[..., None, :])
    h = np.tanh(np.einsum('...tk,...hk->...th', h, W2) + b2[..., None, :])
    return np.einsum('...tk,...ak->...ta', h, W3) + b3[..., None, :]


def tape_fitness(actions, target, clip):
    """Synthetic-tape restatement of Evaluator.single_run utils.py:126-139 with the tape env of
    oracle/gym_stub: action clip (config.py:29,37, applied outside the net utils.py:134), reward
    r_t = -||clip(a_t) - a*_t||^2, return = sum_t r_t (utils.py:137).  Higher is better; the worker
    reports exactly this number (natural_es.py:31-32 negates Evaluator.eval's cost back)."""
    a = np.clip(np.asarray(actions, dtype=np.float64), -clip, clip)
    d = a - np.asarray(target, dtype=np.float64)
    return -(d * d).sum(axis=(-1, -2))


def perturb(theta, sigma, eps):
    """natural_es.py:28-30: fp32 copy of theta, ``+= sigma * epsilon`` (fp64 sum cast to fp32)."""
    theta32 = np.asarray(theta, dtype=np.float32)
    return (theta32.astype(np.float64) + float(sigma) * np.asarray(eps, dtype=np.float64)).astype(np.float32)


def evaluate_population(theta, obs, target, sigma, clip, seed, gen, member_offset, n_members, d0, H, A,
                        chunk=256):
    """Fitness of members [member_offset, member_offset+n_members): sample, perturb, forward, fitness.
    (natural_es.py:27-32 per member.)  Returns fitness fp64 [n_members]."""
    P = param_count(d0, H, A)
    out = np.empty(n_members, dtype=np.float64)
    for s in range(0, n_members, chunk):
        n = min(chunk, n_members - s)
        eps = noise(seed, gen, member_offset + s, n, P)
        thetas = perturb(np.asarray(theta)[None, :], sigma, eps)
        out[s:s + n] = tape_fitness(forward(thetas, obs, d0, H, A), target, clip)
    return out


# --------------------------------------------------------------------------------------------
# Rank shaping, gradient estimate, Adam, update
# --------------------------------------------------------------------------------------------
def ranks_stable(x):
    """Integer ranks 0..N-1, ascending, ties broken by index (the pinned tie rule; the reference's
    default argsort utils.py:145 is unstable, identical on tie-free input).  -0.0 == +0.0; NaN ranks
    last (numpy sort order)."""
    x = np.asarray(x, dtype=np.float64).flatten()
    order = np.argsort(x, kind='stable')
    ranks = np.empty(len(x), dtype=np.int64)
    ranks[order] = np.arange(len(x))
    return ranks


def fitness_shift(x):
    """utils.py:142-148: ranks/(N-1) - 0.5 (fp64)."""
    r = ranks_stable(x).astype(np.float64)
    r /= (len(r) - 1)
    r -= .5
    return r


def nes_gradient(eps, shaped, sigma):
    """natural_es.py:91-92:  mean_i(eps_i * s_i) / sigma  (fp64)."""
    eps = np.asarray(eps, dtype=np.float64)
    s = np.asarray(shaped, dtype=np.float64)
    return (s @ eps) / len(s) / sigma


def nes_gradient_streamed(shaped, sigma, seed, gen, P, chunk=256):
    """Same as nes_gradient with eps regenerated chunk-wise (never materialises [N, P])."""
    s = np.asarray(shaped, dtype=np.float64)
    N = len(s)
    g = np.zeros(P, dtype=np.float64)
    for o in range(0, N, chunk):
        n = min(chunk, N - o)
        g += s[o:o + n] @ noise(seed, gen, o, n, P)
    return g / N / sigma


class Adam:
    """utils.py:150-166 restated (fp64 state, returns the bias-corrected step direction)."""

    def __init__(self, beta1=0.9, beta2=0.999, epsilon=1e-08):
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.beta1_t = self.beta2_t = 1.0
        self.m = 0.0
        self.v = 0.0

    def update(self, g):
        self.beta1_t *= self.beta1
        self.beta2_t *= self.beta2
        self.m = self.beta1 * self.m + (1 - self.beta1) * g
        self.v = self.beta2 * self.v + (1 - self.beta2) * np.power(g, 2)
        m_ = self.m / (1 - self.beta1_t)
        v_ = self.v / (1 - self.beta2_t)
        return m_ / (np.sqrt(v_) + self.epsilon)


def nes_update(theta32, gradient, opt, weight_decay, learning_rate):
    """natural_es.py:93-96.  Returns (theta_new fp32, update fp32) where
    update = lr * fp32(Adam((1 - wd) * g)) — the 'parameter-update vector' of BASELINE.json.

    :93 ``gradient -= wd * gradient`` scales the gradient (it does not decay theta);
    :95 casts the fp64 Adam step to fp32; :96 ``param.add_(lr * gradient)`` is fp32 arithmetic."""
    g = np.array(gradient, dtype=np.float64)
    g -= weight_decay * g
    step = opt.update(g)
    step32 = np.asarray(step, dtype=np.float32)
    update = (np.float32(learning_rate) * step32).astype(np.float32)
    theta_new = (np.asarray(theta32, dtype=np.float32) + update).astype(np.float32)
    return theta_new, update


def nes_generation(theta32, opt, obs, target, *, sigma, clip, seed, gen, N, d0, H, A,
                   weight_decay, learning_rate, fitness=None):
    """One full generation, natural_es.py:62-96 (obs normaliser = identity).  If ``fitness`` is
    given it is used instead of evaluating (layered parity: update given identical fitness)."""
    P = param_count(d0, H, A)
    if fitness is None:
        fitness = evaluate_population(theta32, obs, target, sigma, clip, seed, gen, 0, N, d0, H, A)
    shaped = fitness_shift(fitness)
    g = nes_gradient_streamed(shaped, sigma, seed, gen, P)
    theta_new, update = nes_update(theta32, g, opt, weight_decay, learning_rate)
    return dict(fitness=np.asarray(fitness, dtype=np.float64), shaped=shaped, gradient=g,
                update=update, theta=theta_new)


# --------------------------------------------------------------------------------------------
# Observation normaliser: SharedStats / StaticNormalizer, utils.py:37-106
# --------------------------------------------------------------------------------------------
class ObsStats:
    """utils.py:59-96 restated with the reference's fp32 arithmetic (torch FloatTensors there)."""

    def __init__(self, o_size):
        self.m = np.zeros(o_size, dtype=np.float32)
        self.v = np.zeros(o_size, dtype=np.float32)
        self.n = np.float32(0)

    def feed(self, o):
        """utils.py:68-73 — one observation into the online (Welford) statistics."""
        o = np.asarray(o, dtype=np.float32)
        n = self.n
        new_m = self.m * (n / (n + np.float32(1))) + o / (n + np.float32(1))
        self.v = (self.v * (n / (n + np.float32(1))) + (o - self.m) * (o - new_m) / (n + np.float32(1))).astype(np.float32)
        self.m = new_m.astype(np.float32)
        self.n = np.float32(n + np.float32(1))

    def merge(self, B):
        """utils.py:85-96 — Chan merge of B into self."""
        nA, nB = self.n, B.n
        n = np.float32(nA + nB)
        delta = B.m - self.m
        m = self.m + delta * nB / n
        v = self.v * nA + B.v * nB + delta * delta * nA * nB / n
        v = v / n
        self.m, self.v, self.n = m.astype(np.float32), v.astype(np.float32), n

    def merge_tape(self, obs, n_feed):
        """Closed form of "feed the tape n_feed/T times, then merge": the online statistics of any number of whole
        passes over the same tape are its mean and population variance (what des_obs_stats_merge computes)."""
        obs = np.asarray(obs, dtype=np.float64)
        B = ObsStats(obs.shape[1])
        B.m = obs.mean(0).astype(np.float32)
        B.v = obs.var(0).astype(np.float32)
        B.n = np.float32(n_feed)
        mA, vA, nA = self.m.astype(np.float64), self.v.astype(np.float64), float(self.n)
        mb, vb, nB = obs.mean(0), obs.var(0), float(n_feed)
        n = nA + nB
        delta = mb - mA
        self.m = (mA + delta * nB / n).astype(np.float32)
        self.v = ((vA * nA + vb * nB + delta * delta * nA * nB / n) / n).astype(np.float32)
        self.n = np.float32(n)

    def normalize(self, o):
        """StaticNormalizer.__call__ utils.py:48-51: raw while n == 0, else (o - m)/sqrt(v + 1e-6) in fp32."""
        o = np.asarray(o, dtype=np.float32)
        if self.n == 0:
            return o
        std = (self.v + np.float32(1e-6)) ** np.float32(.5)
        return ((o - self.m) / std).astype(np.float32)


# --------------------------------------------------------------------------------------------
# Synthetic inputs of SURVEY.md §8d (identical on every machine)
# --------------------------------------------------------------------------------------------
def synthetic_tape(T, d0, A, seed=1234):
    """obs X[T,d0] ~ N(0,1) fp32, targets a*[T,A] = tanh(N(0,1)) fp32, RandomState(seed)."""
    rs = np.random.RandomState(seed)
    obs = rs.randn(T, d0).astype(np.float32)
    target = np.tanh(rs.randn(T, A)).astype(np.float32)
    return obs, target


def synthetic_theta(d0, H, A, seed=0):
    """nn.Linear-style init U(-1/sqrt(fan_in), 1/sqrt(fan_in)) from RandomState(seed), in the flat
    layout.  (config.py:14-16 takes torch's default init; torch is not an oracle dependency, so the
    benchmark uses this numpy equivalent — same distribution, documented seed.)"""
    rs = np.random.RandomState(seed)
    parts = []
    for fan_out, fan_in in ((H, d0), (H, H), (A, H)):
        b = 1.0 / np.sqrt(fan_in)
        parts.append(rs.uniform(-b, b, size=fan_out * fan_in))
        parts.append(rs.uniform(-b, b, size=fan_out))
    return np.concatenate(parts).astype(np.float32)
