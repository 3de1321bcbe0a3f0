// p2p.cu — one-shot all-reduce of the small replicated operands over NVLink peer memory.
//
// Why: the power iteration all-reduces {64 x 64 Gram + column sums (f64), n_eff x 64 partial panel (f32)} once per half step —
// 0.5 MB at config 3.  Through NCCL that measured ~60 us per call on 8 B200s (7 GB/s algorithmic bandwidth), 0.64 ms of a
// 5.95 ms fit.  This file tests whether the collective's own latency is the cost (it is not, see p2p_init): every rank owns a
// buffer that all peers map (CUDA IPC); an all-reduce is
// ONE kernel per rank: publish the local operand in the own buffer, raise the own flag (release, system scope), wait for the
// peers' flags (acquire), then every rank reads all operands over NVLink and sums them IN RANK ORDER — all ranks get
// bit-identical sums, which the replicated small-side factorisations rely on.  Two alternating halves make reuse safe: a rank
// can only start all-reduce s + 2 after it saw every peer's flag of s + 1, i.e. after every peer finished reading s.
//
// Falls back to NCCL when peer mapping is unavailable (decided collectively at context creation) or the operand is larger
// than a half (column statistics of very wide matrices).
#include <algorithm>

#include "common.cuh"

namespace salg {

constexpr size_t P2P_HALF = (size_t)4 << 20;       // bytes per half
constexpr int P2P_MAX_RANKS = 8;
constexpr int P2P_CTAS = 128, P2P_THREADS = 256;

struct P2P {
    int nranks = 0, rank = 0;
    uint8_t* local = nullptr;                       // [2][P2P_HALF] data, then flag (u32), then CTA counter (u32)
    uint8_t* peer[P2P_MAX_RANKS] = {nullptr};       // peer[rank] == local
    uint8_t** d_peer = nullptr;
    uint32_t seq = 0;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double2 ld_sys_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_f32x4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// n64 doubles (padded to even) followed by n32 floats (padded to a multiple of 4) per operand
__global__ void __launch_bounds__(P2P_THREADS)
p2p_allreduce_kernel(uint8_t* const* __restrict__ peers, int rank, int nranks, uint32_t seq, double* buf64, size_t n64, float* buf32,
                     size_t n32) {
    const size_t half_off = (size_t)(seq & 1u) * P2P_HALF;
    uint8_t* mine = peers[rank];
    uint8_t* peer_r[P2P_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < P2P_MAX_RANKS; r++) peer_r[r] = peers[r < nranks ? r : rank];
    const size_t n64p = (n64 + 1) & ~(size_t)1, n32p = (n32 + 3) & ~(size_t)3;
    double* my64 = reinterpret_cast<double*>(mine + half_off);
    float* my32 = reinterpret_cast<float*>(mine + half_off + n64p * 8);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    // 1. publish the local operand
    for (size_t i = tid; i < n64p; i += nth) my64[i] = i < n64 ? buf64[i] : 0.0;
    for (size_t i = tid; i < n32p; i += nth) my32[i] = i < n32 ? buf32[i] : 0.f;
    __threadfence_system();
    __syncthreads();
    uint32_t* flag = reinterpret_cast<uint32_t*>(mine + 2 * P2P_HALF);
    uint32_t* counter = flag + 16;
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(counter, 1u);
        if (t == gridDim.x - 1) {                   // every CTA of this rank has published its slice
            *counter = 0;
            __threadfence_system();
            st_release_sys(flag, seq);
        }
    }
    // 2. wait for every rank's flag of this all-reduce
    if (threadIdx.x < nranks) {
        const uint32_t* f = reinterpret_cast<const uint32_t*>(peers[threadIdx.x] + 2 * P2P_HALF);
        unsigned long long spins = 0;
        while ((int32_t)(ld_acquire_sys(f) - seq) < 0) {
            if (++spins > (1ull << 25)) __trap();   // a peer never arrived (about a minute of polling): do not hang the GPU
        }
    }
    __syncthreads();
    // 3. sum in rank order (identical on every rank); all ranks' loads of an element are in flight together
    for (size_t i = tid * 2; i < n64p; i += nth * 2) {
        double2 v[P2P_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; r++)
            if (r < nranks) v[r] = ld_sys_f64x2(reinterpret_cast<const double*>(peer_r[r] + half_off) + i);
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; r++)
            if (r < nranks) {
                acc.x += v[r].x;
                acc.y += v[r].y;
            }
        if (i < n64) buf64[i] = acc.x;
        if (i + 1 < n64) buf64[i + 1] = acc.y;
    }
    for (size_t i = tid * 4; i < n32p; i += nth * 4) {
        float4 v[P2P_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; r++)
            if (r < nranks) v[r] = ld_sys_f32x4(reinterpret_cast<const float*>(peer_r[r] + half_off + n64p * 8) + i);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; r++)
            if (r < nranks) {
                acc.x += v[r].x;
                acc.y += v[r].y;
                acc.z += v[r].z;
                acc.w += v[r].w;
            }
        if (i + 3 < n32) {
            *reinterpret_cast<float4*>(buf32 + i) = acc;
        } else {
            if (i < n32) buf32[i] = acc.x;
            if (i + 1 < n32) buf32[i + 1] = acc.y;
            if (i + 2 < n32) buf32[i + 2] = acc.z;
        }
    }
}

void p2p_destroy(salg_ctx* ctx) {
    P2P* p = (P2P*)ctx->p2p;
    if (!p) return;
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < p->nranks; r++)
        if (r != p->rank && p->peer[r]) cudaIpcCloseMemHandle(p->peer[r]);
    if (p->d_peer) cudaFree(p->d_peer);
    if (p->local) cudaFree(p->local);
    delete p;
    ctx->p2p = nullptr;
}

// Collective (every rank of ctx->comm calls it right after ncclCommInitRank).  Leaves ctx->p2p == nullptr on every rank unless
// every rank could map every peer.
void p2p_init(salg_ctx* ctx) {
    // Opt-in (SALG_P2P=1).  Measured at config 3: 32 us per half-step all-reduce against NCCL's 28 us on 2 GPUs, 63 us against
    // 57 us on 8 — the time of these 0.5 MB all-reduces is the skew between the ranks arriving, not the collective's own
    // latency, so the library's kernel buys nothing over NCCL here and NCCL stays the default.
    const char* e = getenv("SALG_P2P");
    const bool want = e && atoi(e) != 0;
    if (ctx->nranks <= 1 || ctx->nranks > P2P_MAX_RANKS || !want) return;
    cudaStream_t st = ctx->stream;
    P2P* p = new P2P();
    p->nranks = ctx->nranks;
    p->rank = ctx->rank;
    int ok = 1;
    const size_t bytes = 2 * P2P_HALF + 4096;
    if (cudaMalloc(&p->local, bytes) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        p->local = nullptr;
    }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok) {
        cudaMemsetAsync(p->local, 0, bytes, st);
        if (cudaIpcGetMemHandle(&mine, p->local) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        }
    }
    // exchange the handles + the physical device of every rank (and, below, whether everybody could open them) through NCCL
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    struct Card { cudaIpcMemHandle_t h; int pci[4]; } card;
    card.h = mine;
    cudaDeviceProp prop;
    SALG_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    card.pci[0] = prop.pciDomainID; card.pci[1] = prop.pciBusID; card.pci[2] = prop.pciDeviceID; card.pci[3] = 0;
    uint8_t *d_send = nullptr, *d_recv = nullptr;
    SALG_CUDA(cudaMalloc(&d_send, sizeof(Card)));
    SALG_CUDA(cudaMalloc(&d_recv, sizeof(Card) * p->nranks));
    SALG_CUDA(cudaMemcpyAsync(d_send, &card, sizeof(Card), cudaMemcpyHostToDevice, st));
    SALG_NCCL(ncclAllGather(d_send, d_recv, sizeof(Card), ncclChar, ctx->comm, st));
    std::vector<Card> cards(p->nranks);
    SALG_CUDA(cudaMemcpyAsync(cards.data(), d_recv, sizeof(Card) * p->nranks, cudaMemcpyDeviceToHost, st));
    SALG_CUDA(cudaStreamSynchronize(st));
    std::vector<cudaIpcMemHandle_t> all(p->nranks);
    for (int r = 0; r < p->nranks; r++) {
        all[r] = cards[r].h;
        // two ranks on one GPU: their kernels wait on one another and nothing guarantees that they run at the same time
        for (int q = 0; q < r; q++)
            if (memcmp(cards[q].pci, cards[r].pci, sizeof(card.pci)) == 0) ok = 0;
    }
    if (ok) {
        for (int r = 0; r < p->nranks && ok; r++) {
            if (r == p->rank) {
                p->peer[r] = p->local;
                continue;
            }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            } else {
                p->peer[r] = (uint8_t*)ptr;
            }
        }
    }
    int* d_ok = (int*)d_send;
    SALG_CUDA(cudaMemcpyAsync(d_ok, &ok, 4, cudaMemcpyHostToDevice, st));
    SALG_NCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, ctx->comm, st));
    int all_ok = 0;
    SALG_CUDA(cudaMemcpyAsync(&all_ok, d_ok, 4, cudaMemcpyDeviceToHost, st));
    SALG_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_send);
    cudaFree(d_recv);
    ctx->p2p = p;
    if (!all_ok) {
        p2p_destroy(ctx);
        return;
    }
    SALG_CUDA(cudaMalloc(&p->d_peer, sizeof(uint8_t*) * P2P_MAX_RANKS));
    SALG_CUDA(cudaMemcpyAsync(p->d_peer, p->peer, sizeof(uint8_t*) * P2P_MAX_RANKS, cudaMemcpyHostToDevice, st));
    SALG_CUDA(cudaStreamSynchronize(st));
}

// true if the operands were reduced here (otherwise the caller goes through NCCL)
bool p2p_allreduce(salg_ctx* ctx, double* buf64, size_t n64, float* buf32, size_t n32) {
    P2P* p = (P2P*)ctx->p2p;
    if (!p) return false;
    const size_t need = ((n64 + 1) & ~(size_t)1) * 8 + ((n32 + 3) & ~(size_t)3) * 4;
    if (need > P2P_HALF || need == 0) return false;
    p->seq++;
    const size_t work = std::max(n64 / 2, n32 / 4);
    const int ctas = (int)std::max<size_t>(1, std::min<size_t>(P2P_CTAS, (work + P2P_THREADS - 1) / P2P_THREADS));
    p2p_allreduce_kernel<<<ctas, P2P_THREADS, 0, ctx->stream>>>(p->d_peer, p->rank, p->nranks, p->seq, buf64, n64, buf32, n32);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    return true;
}

}  // namespace salg
