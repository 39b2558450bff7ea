// des_comm_*: the two exchange steps of a sharded NES generation as kernels over peer memory (NVLink / NVSwitch), one
// process per GPU on one node.  They replace the reference's pipe traffic of natural_es.py:62-75 (every worker ships
// (epsilon, fitness, steps) to the master) and the two NCCL all-reduces round 1 issued from Python:
//
//   des_comm_allgather_fitness   every rank stores its shard of fitness[N] straight into every peer's copy, then a
//                                flag barrier: afterwards all ranks hold the identical fitness_all (centered ranks are
//                                global, utils.py:142-148)
//   des_comm_push_partial        every rank stores its partial[P] = sum_{i in shard} s_i eps_i into slot [rank] of every
//                                peer's slot table and raises its flag there
//   des_comm_reduce_partial      waits for all flags, then sums the G slots in RANK ORDER (fixed order: bit-identical on
//                                every rank, so the update needs no broadcast) — natural_es.py:91 across shards
//
// Memory: each rank cudaMalloc's one block [header | fitness_all[N] | slots[G][Ppad]] and exports it with cudaIpc; peers
// map it (cudaIpcOpenMemHandle enables P2P).  Flags carry a monotonically increasing epoch kept in the owner's header,
// so the kernels are CUDA-graph capturable (no host-side counter).  Ordering: data stores, __threadfence_system(),
// then the flag store; the waiter spins on its OWN memory with volatile loads and fences before reading the data.
// Buffer reuse is safe without further handshakes: a rank can only overwrite a peer's fitness shard of generation g+1
// after the partial exchange of generation g, which that peer enters after it has consumed fitness_all of generation g;
// and partial slots of g+1 are written after the fitness barrier of g+1, which a peer enters after its apply of g.
#include <new>
#include <string.h>
#include "des_common.cuh"

namespace des {

constexpr int kMaxWorld = 16;
constexpr size_t kCommHeader = 1024;      // flags_fit[16] | flags_part[16] | epoch_fit | epoch_part | fit_done | seg_done[16], all_done (u32)

struct CommDev {
    uint8_t *base[kMaxWorld];             // base[r] = rank r's block as mapped in THIS process
    int rank, world;
    int64_t N, Ppad;
    size_t off_fit, off_slots;
};

__device__ __forceinline__ unsigned long long *flags_fit(uint8_t *base) { return reinterpret_cast<unsigned long long *>(base); }
__device__ __forceinline__ unsigned long long *flags_part(uint8_t *base) { return reinterpret_cast<unsigned long long *>(base) + kMaxWorld; }
__device__ __forceinline__ unsigned long long *epoch_fit(uint8_t *base) { return reinterpret_cast<unsigned long long *>(base) + 2 * kMaxWorld; }
__device__ __forceinline__ unsigned long long *epoch_part(uint8_t *base) { return reinterpret_cast<unsigned long long *>(base) + 2 * kMaxWorld + 1; }

__device__ __forceinline__ void wait_flag(volatile unsigned long long *f, unsigned long long epoch) {
    while (*f < epoch) __nanosleep(64);
}

// One CTA per peer: CTA b stores the shard into peer b's fitness_all, raises this rank's flag there, and waits for peer
// b's flag here — the kernel completes when every peer's shard has landed.  (Round 2, first version: a single CTA pushed to
// all peers in turn; at 8 GPUs the two exchange kernels were ~60 us of a 1.75 ms generation.)
__global__ void __launch_bounds__(1024) comm_allgather_fitness_kernel(CommDev c, int64_t offset, int64_t n_local) {
    uint8_t *mine = c.base[c.rank];
    const int b = blockIdx.x;
    const unsigned long long epoch = *epoch_fit(mine) + 1;      // every CTA reads it before the last one bumps it (below)
    if (b != c.rank) {
        const float *src = reinterpret_cast<const float *>(mine + c.off_fit) + offset;
        float *dst = reinterpret_cast<float *>(c.base[b] + c.off_fit) + offset;
        for (int64_t i = threadIdx.x; i < n_local; i += blockDim.x) dst[i] = src[i];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            *reinterpret_cast<volatile unsigned long long *>(flags_fit(c.base[b]) + c.rank) = epoch;
            wait_flag(flags_fit(mine) + b, epoch);
            __threadfence_system();
        }
    }
    __syncthreads();
    // the last CTA to finish advances the epoch for the next generation (a device-side counter: graph replay safe)
    if (threadIdx.x == 0) {
        unsigned int *done = reinterpret_cast<unsigned int *>(epoch_part(mine) + 1);
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            *done = 0;
            *epoch_fit(mine) = epoch;
        }
    }
}

// kPushSeg CTAs per peer: CTA (b, seg) stores segment seg of partial[P] (zero-padded to Ppad) into slot [rank] of rank b's
// table (its own included); the last segment to finish raises this rank's flag at rank b.
constexpr int kPushSeg = 4;
__global__ void __launch_bounds__(512) comm_push_partial_kernel(CommDev c, const float *__restrict__ partial, int64_t P) {
    uint8_t *mine = c.base[c.rank];
    const int b = blockIdx.x / kPushSeg, seg = blockIdx.x % kPushSeg;
    const unsigned long long epoch = *epoch_part(mine) + 1;
    const int64_t nq = c.Ppad / 4;
    const int64_t q0 = nq * seg / kPushSeg, q1 = nq * (seg + 1) / kPushSeg;
    float4 *dst = reinterpret_cast<float4 *>(c.base[b] + c.off_slots) + (int64_t)c.rank * nq;
    for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        float4 v;
        const int64_t j = 4 * q;
        v.x = j < P ? partial[j] : 0.f;
        v.y = j + 1 < P ? partial[j + 1] : 0.f;
        v.z = j + 2 < P ? partial[j + 2] : 0.f;
        v.w = j + 3 < P ? partial[j + 3] : 0.f;
        dst[q] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int *seg_done = reinterpret_cast<unsigned int *>(epoch_part(mine) + 2) + b;       // per-peer segment counters
        unsigned int *all_done = reinterpret_cast<unsigned int *>(epoch_part(mine) + 2) + kMaxWorld;
        if (atomicAdd(seg_done, 1u) == kPushSeg - 1) {
            *seg_done = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(flags_part(c.base[b]) + c.rank) = epoch;
        }
        if (atomicAdd(all_done, 1u) == gridDim.x - 1) {
            *all_done = 0;
            __threadfence();
            *epoch_part(mine) = epoch;                     // read by the reduce kernel that follows on the stream
        }
    }
}

// partial_sum[j] = sum over ranks r = 0..G-1 (in that order) of slot[r][j], after every rank's flag has arrived
__global__ void __launch_bounds__(256) comm_reduce_partial_kernel(CommDev c, float *__restrict__ out, int64_t P) {
    uint8_t *mine = c.base[c.rank];
    if ((int)threadIdx.x < c.world) {
        wait_flag(flags_part(mine) + threadIdx.x, *epoch_part(mine));
        __threadfence_system();
    }
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P) return;
    const volatile float *slots = reinterpret_cast<const volatile float *>(mine + c.off_slots);
    float s = slots[j];
    for (int r = 1; r < c.world; ++r) s += slots[(int64_t)r * c.Ppad + j];
    out[j] = s;
}

}  // namespace des

struct des_comm {
    des::CommDev dev;
    int device;
    size_t bytes;
    bool connected;
    void *peer_mapped[des::kMaxWorld];
};

extern "C" DES_API int des_comm_create(des_comm **out, int rank, int world, int64_t N, int64_t P, void *ipc_handle_out) {
    using namespace des;
    DES_REQUIRE(out && ipc_handle_out, "des_comm_create: NULL pointer");
    DES_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "des_comm_create: bad rank %d / world %d (max %d)",
                rank, world, kMaxWorld);
    DES_REQUIRE(N >= 2 && P >= 1, "des_comm_create: bad sizes N=%lld P=%lld", (long long)N, (long long)P);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle is 64 bytes in the C ABI");
    des_comm *c = new (std::nothrow) des_comm();
    DES_REQUIRE(c, "des_comm_create: out of host memory");
    memset(c, 0, sizeof(*c));
    c->dev.rank = rank; c->dev.world = world; c->dev.N = N;
    c->dev.Ppad = (P + 3) / 4 * 4;
    c->dev.off_fit = kCommHeader;
    c->dev.off_slots = (kCommHeader + (size_t)N * sizeof(float) + 255) & ~(size_t)255;
    c->bytes = c->dev.off_slots + (size_t)world * (size_t)c->dev.Ppad * sizeof(float);
    cudaError_t e = cudaGetDevice(&c->device);
    void *base = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&base, c->bytes);
    if (e == cudaSuccess) e = cudaMemset(base, 0, c->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(ipc_handle_out), base);
    if (e != cudaSuccess) {
        const int rc = cuda_fail(e, "des_comm_create (cudaMalloc / cudaIpcGetMemHandle)");
        if (base) cudaFree(base);
        delete c;
        return rc;
    }
    c->dev.base[rank] = (uint8_t *)base;
    *out = c;
    return DES_OK;
}

extern "C" DES_API int des_comm_connect(des_comm *c, const void *all_handles) {
    using namespace des;
    DES_REQUIRE(c && all_handles, "des_comm_connect: NULL pointer");
    const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(all_handles);
    for (int r = 0; r < c->dev.world; ++r) {
        if (r == c->dev.rank) continue;
        void *p = nullptr;
        DES_CUDA(cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess));
        c->peer_mapped[r] = p;
        c->dev.base[r] = (uint8_t *)p;
    }
    c->connected = true;
    return DES_OK;
}

extern "C" DES_API void des_comm_destroy(des_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->dev.world; ++r)
        if (c->peer_mapped[r]) cudaIpcCloseMemHandle(c->peer_mapped[r]);
    if (c->dev.base[c->dev.rank]) cudaFree(c->dev.base[c->dev.rank]);
    delete c;
}

extern "C" DES_API float *des_comm_fitness_all_dev(des_comm *c) {
    return c ? reinterpret_cast<float *>(c->dev.base[c->dev.rank] + c->dev.off_fit) : nullptr;
}

extern "C" DES_API int des_comm_allgather_fitness(des_comm *c, int64_t member_offset, int64_t n_local, void *stream) {
    using namespace des;
    DES_REQUIRE(c && c->connected, "des_comm_allgather_fitness: communicator not connected");
    DES_REQUIRE(member_offset >= 0 && n_local >= 0 && member_offset + n_local <= c->dev.N, "des_comm_allgather_fitness: bad shard");
    comm_allgather_fitness_kernel<<<c->dev.world, 1024, 0, (cudaStream_t)stream>>>(c->dev, member_offset, n_local);
    DES_LAUNCH_CHECK("comm_allgather_fitness_kernel");
    return DES_OK;
}

extern "C" DES_API int des_comm_allreduce_partial(des_comm *c, float *partial_sum_out_dev, const float *partial_dev, int64_t P,
                                                  void *stream) {
    using namespace des;
    DES_REQUIRE(c && c->connected, "des_comm_allreduce_partial: communicator not connected");
    DES_REQUIRE(partial_sum_out_dev && partial_dev && P >= 1 && (P + 3) / 4 * 4 == c->dev.Ppad,
                "des_comm_allreduce_partial: bad arguments (P=%lld)", (long long)P);
    cudaStream_t st = (cudaStream_t)stream;
    comm_push_partial_kernel<<<c->dev.world * kPushSeg, 512, 0, st>>>(c->dev, partial_dev, P);
    DES_LAUNCH_CHECK("comm_push_partial_kernel");
    comm_reduce_partial_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(c->dev, partial_sum_out_dev, P);
    DES_LAUNCH_CHECK("comm_reduce_partial_kernel");
    return DES_OK;
}
