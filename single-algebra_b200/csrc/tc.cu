// tc.cu — tile-densified sparse x panel products on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Why: per stored entry a CUDA-core CSR kernel needs one distinct 256 B panel row from L1/L2;
// measured on B200 that gather caps the products of the randomized-SVD power iteration at 2-6 % of the HBM
// roofline (profiles/r01_v1_summary.md).  Here the operator is re-tiled once into 128-row x 64-column tiles
// (entry stream sorted by tile, 8 B per entry).  A CTA scatters tiles into a zeroed dense fp16 tile in shared
// memory (UMMA canonical K-major layout, no swizzle) and contracts it with the matching slice of the panel by
// tcgen05.mma, accumulating in TMEM; the panel slice is fetched once per tile pair instead of once per entry.
//
// Operand roles: the DENSE panel slice is the M = 128 operand (64 panel columns x 2 split terms, interleaved
// m = 2*column + term), the SPARSE tile pair is the N operand (N = 256 rows for A X, N = 128 operator columns for
// A^T Y).  A wide N keeps the shared-memory operand traffic per MMA below the 128 B/clk/SM the tensor core can
// be fed with (an earlier layout with the panel as the N = 64 operand measured 80 clk per MMA instead of 32).
//
// Precision: fp16 x fp16 products are exact in f32 and accumulate in f32.  The panel is scaled by a power of two
// (largest magnitude near 2^13) and split into two fp16 terms (22 significant bits); the operator likewise unless
// every stored value is exactly representable in fp16 (raw counts <= 2048), in which case one term suffices.
// The two term rows of a panel column are summed in the epilogue (adjacent TMEM lanes -> one shuffle).
//
// Roles inside a CTA (warp-specialised, all hand-offs through mbarriers):
//   warps 0-15   scatter: clear the tile buffer, scatter the unit's entries, fence.proxy.async, signal
//   warp  16     panel-slice loader: cp.async.bulk of the pre-split slice (canonical layout) into a ring
//   warp  17     one thread issues tcgen05.mma (M=128, K=16 per instruction) and tcgen05.commit
//   warps 18-21  epilogue: tcgen05.ld the f32 accumulator, sum the term rows, rescale, store / atomically add
//   warp  22     entry loader: lane j owns ring slot j, one cp.async.bulk per tile
#include <cub/cub.cuh>
#include <cuda_fp16.h>

#include "common.cuh"

namespace salg {

constexpr int TC_RB = 128;          // tile rows
constexpr int TC_CB = 64;           // tile columns
constexpr int TC_SCATTER_WARPS = 16;
constexpr int TC_SCATTER_THREADS = TC_SCATTER_WARPS * 32;
constexpr int TC_THREADS = (TC_SCATTER_WARPS + 10) * 32;   // + panel loader, mma, 4 epilogue, 4 entry loaders
constexpr int TC_W_BLOAD = TC_SCATTER_WARPS, TC_W_MMA = TC_SCATTER_WARPS + 1, TC_W_EPI = TC_SCATTER_WARPS + 2,
              TC_W_ELOAD = TC_SCATTER_WARPS + 6;
constexpr int TC_RPAD = 4;            // row blocks are padded to a multiple of this (the A X kernel walks pairs)
constexpr int TC_SLOT_ENTRIES = 768;  // entries per ring slot (one tile); denser tiles read their tail from global memory
constexpr int TC_SLOT_BYTES = (TC_SLOT_ENTRIES + 2) * 8;
constexpr int TC_S_BYTES = 32768;     // sparse operand buffer: 256 x 64 (A X) or 128 x 128 (A^T Y) fp16
constexpr int TC_NSB = 2;             // sparse operand buffers in flight (scatter runs up to 2 passes ahead of the MMA)
constexpr uint32_t TC_SPIN_LIMIT = 1u << 24;

struct TcTiles {
    uint2* entries = nullptr;       // [nnz] .x = half byte-offsets in the A X (bits 0-13) and A^T Y (bits 14-27) operand
                                    //       buffers, .y = f32 bits of the value
    int64_t* tile_ptr = nullptr;    // [n_rb * n_cb + 1]
    int n_rb = 0, n_cb = 0;
    int a_terms = 2;                // 1 when every value is exact in fp16
    float a_scale = 1.f;            // power of two applied to the operator values before the split
    int64_t nnz = 0;
};

void tc_free(salg_ctx* owner, void* p) {
    TcTiles* t = (TcTiles*)p;
    if (!t) return;
    dev_free(owner, t->entries);
    dev_free(owner, t->tile_ptr);
    delete t;
}

// ---- PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ unsigned long long g_tc_dbg[32];   // timing experiment counters of CTA 0 (SALG_TC_DBG=32)
#define TC_T(acc) do { long long _t = clock64(); acc += _t - t_prev; t_prev = _t; } while (0)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && ++spins > TC_SPIN_LIMIT) __trap();   // never hang the GPU on a protocol bug
    } while (!ok);
}
// one lane polls, the warp follows
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Four consecutive K-steps (K = 16 fp16 each) of one product in ONE asm block: the issuing thread is the serial
// resource of the CTA, so per-MMA overhead is two 64-bit adds.  Descriptor start addresses advance by a_step /
// b_step (16 B units) per K-step; only the first MMA may overwrite the accumulator.
__device__ __forceinline__ void umma_f16_run4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate_first, uint64_t a_step, uint64_t b_step) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "mov.b64 da, %1;\n\tmov.b64 db, %2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "add.u64 da, da, %5;\n\tadd.u64 db, db, %6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
        "add.u64 da, da, %5;\n\tadd.u64 db, db, %6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
        "add.u64 da, da, %5;\n\tadd.u64 db, db, %6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
        "}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "l"(a_step), "l"(b_step)
        : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B (128 B contiguous);
// LBO = byte distance between the two 16 B K-chunks of one instruction, SBO = distance between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D f32 (bit 4), A/B fp16 (format 0), both K-major, M = 128, N as given
__host__ __device__ constexpr uint32_t tc_idesc(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp16 term `t` (0, 1) of a scaled float: x ~ h0 + h1 with 11 significant bits each
__device__ __forceinline__ unsigned short f16_term(float x, int t) {
    __half h = __float2half_rn(x);
    if (t > 0) h = __float2half_rn(x - __half2float(h));
    return __half_as_ushort(h);
}

// byte offset of element (mn, k) inside a canonical K-major no-swizzle operand whose 16 B K-chunks are
// `chunk_stride` bytes apart (8-row groups are 128 B apart)
__device__ __forceinline__ uint32_t canon_off(uint32_t mn, uint32_t k, uint32_t chunk_stride) {
    return (k >> 3) * chunk_stride + (mn >> 3) * 128u + (mn & 7u) * 16u + (k & 7u) * 2u;
}

// power-of-two scale that puts `amax` just below 2^14 (fp16 max is 65504; the second term is 2^-11 smaller)
__host__ __device__ inline float tc_pow2_scale(float amax) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int e;
    frexpf(amax, &e);                 // amax = f * 2^e, f in [0.5, 1)
    return ldexpf(1.f, 14 - e);
}

// ---- tile format builder -----------------------------------------------------------------------------------------
template <typename T>
__global__ void tc_keys_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col, const T* __restrict__ val,
                               int64_t nrows, int n_cb, uint32_t* __restrict__ keys, uint2* __restrict__ payload,
                               unsigned* __restrict__ info /* [0] inexact flag, [1] bits of max |v| */) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    float amax = 0.f;
    for (int64_t r = w; r < nrows; r += nw) {
        int64_t s = ptr[r], e = ptr[r + 1];
        uint32_t rb = (uint32_t)(r / TC_RB), lr = (uint32_t)(r % TC_RB);
        for (int64_t p = s + lane; p < e; p += 32) {
            uint32_t c = col[p];
            float v = (float)val[p];
            keys[p] = rb * (uint32_t)n_cb + c / TC_CB;
            const uint32_t lc = c % TC_CB, cb = c / TC_CB;
            // positions inside the 32 KB sparse-operand buffers of the two kernels (see tc_scatter_role)
            const uint32_t off_ax = canon_off(lr + 128u * (rb & 1u), lc, 4096u) >> 1;
            const uint32_t off_aty = canon_off(lc + 64u * (cb & 1u), lr, 2048u) >> 1;
            payload[p] = make_uint2(off_ax | (off_aty << 14), __float_as_uint(v));
            bad |= (__half2float(__float2half_rn(v)) != v);
            amax = fmaxf(amax, fabsf(v));
        }
    }
    if (bad) atomicOr(&info[0], 1u);
    atomicMax(&info[1], __float_as_uint(amax));
}

__global__ void tc_tile_ptr_kernel(const uint32_t* __restrict__ keys_sorted, int64_t nnz, int64_t n_tiles,
                                   int64_t* __restrict__ tile_ptr) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    int64_t lo = 0, hi = nnz;   // first position whose key >= t
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)keys_sorted[mid] < t) lo = mid + 1; else hi = mid;
    }
    tile_ptr[t] = lo;
}

// One CTA per 128-row block: the block's entries are contiguous in the CSR, so the tile order (row block, column block)
// is a counting sort INSIDE the block's own range [ptr[128 rb], ptr[128 (rb+1)]) — histogram of column blocks in shared
// memory, exclusive scan -> tile_ptr, second sweep (L2-resident) places the entries.  The order of the entries inside a
// tile does not matter (every entry owns its slot of the dense tile), so placement uses warp-aggregated shared-memory
// cursors instead of a stable global radix sort (measured 4.2 ms -> see profiles/ for the cfg3 operator).
constexpr int TC_BIN_THREADS = 512;
constexpr int TC_BIN_MAX_CB = 4096;
template <typename T>
__global__ void __launch_bounds__(TC_BIN_THREADS)
tc_bin_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col, const T* __restrict__ val, int64_t nrows,
              int64_t nnz, int n_rb, int n_cb, uint2* __restrict__ entries, int64_t* __restrict__ tile_ptr,
              unsigned* __restrict__ info /* [0] inexact flag, [1] bits of max |v| */) {
    extern __shared__ unsigned bin_sm[];
    unsigned* hist = bin_sm;             // [n_cb] counts, then running cursors
    __shared__ unsigned s_warp[TC_BIN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = TC_BIN_THREADS / 32;
    for (int rb = blockIdx.x; rb < n_rb; rb += gridDim.x) {
        const int64_t r0 = (int64_t)rb * TC_RB;
        const int64_t r1 = r0 + TC_RB < nrows ? r0 + TC_RB : nrows;
        int64_t* tp = tile_ptr + (int64_t)rb * n_cb;
        if (r0 >= nrows) {                                    // padding row block: empty tiles at the end of the stream
            for (int i = tid; i < n_cb; i += TC_BIN_THREADS) tp[i] = nnz;
            continue;
        }
        const int64_t base = ptr[r0];
        __syncthreads();
        for (int i = tid; i < n_cb; i += TC_BIN_THREADS) hist[i] = 0;
        __syncthreads();
        // sweep 1: column-block histogram (columns ascend inside a row: equal blocks sit in adjacent lanes)
        for (int64_t r = r0 + warp; r < r1; r += NW) {
            const int64_t s = ptr[r], e = ptr[r + 1];
            for (int64_t p0 = s; p0 < e; p0 += 32) {
                const int64_t p = p0 + lane;
                const bool ok = p < e;
                const unsigned cb = ok ? col[p] / TC_CB : 0xFFFFFFFFu;
                const unsigned peers = __match_any_sync(0xFFFFFFFFu, cb);
                if (ok && lane == __ffs(peers) - 1) atomicAdd(&hist[cb], (unsigned)__popc(peers));
            }
        }
        __syncthreads();
        // exclusive scan of the histogram (n_cb <= 4096: each thread owns a contiguous run)
        const int per = (n_cb + TC_BIN_THREADS - 1) / TC_BIN_THREADS;
        const int i0 = tid * per, i1 = (i0 + per < n_cb) ? i0 + per : n_cb;
        unsigned mine = 0;
        for (int i = i0; i < i1; i++) mine += hist[i];
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned woff = 0;
        for (int w = 0; w < warp; w++) woff += s_warp[w];
        unsigned run = woff + incl - mine;
        for (int i = i0; i < i1; i++) {
            const unsigned cnt = hist[i];
            hist[i] = run;                                    // cursor of tile (rb, i), relative to `base`
            tp[i] = base + run;
            run += cnt;
        }
        __syncthreads();
        // sweep 2: place the entries
        bool bad = false;
        float amax = 0.f;
        const unsigned rb1 = (unsigned)rb & 1u;
        for (int64_t r = r0 + warp; r < r1; r += NW) {
            const int64_t s = ptr[r], e = ptr[r + 1];
            const unsigned lr = (unsigned)(r - r0);
            for (int64_t p0 = s; p0 < e; p0 += 32) {
                const int64_t p = p0 + lane;
                const bool ok = p < e;
                const unsigned c = ok ? col[p] : 0u;
                const unsigned cb = ok ? c / TC_CB : 0xFFFFFFFFu;
                const unsigned peers = __match_any_sync(0xFFFFFFFFu, cb);
                const int leader = __ffs(peers) - 1;
                unsigned start = 0;
                if (ok && lane == leader) start = atomicAdd(&hist[cb], (unsigned)__popc(peers));
                start = __shfl_sync(0xFFFFFFFFu, start, leader);
                if (ok) {
                    const float v = (float)val[p];
                    const unsigned lc = c % TC_CB;
                    // positions inside the 32 KB sparse-operand buffers of the two kernels (see tc_scatter_role)
                    const uint32_t off_ax = canon_off(lr + 128u * rb1, lc, 4096u) >> 1;
                    const uint32_t off_aty = canon_off(lc + 64u * (cb & 1u), lr, 2048u) >> 1;
                    entries[base + start + __popc(peers & ((1u << lane) - 1u))] =
                        make_uint2(off_ax | (off_aty << 14), __float_as_uint(v));
                    bad |= (__half2float(__float2half_rn(v)) != v);
                    amax = fmaxf(amax, fabsf(v));
                }
            }
        }
        if (bad) atomicOr(&info[0], 1u);
        atomicMax(&info[1], __float_as_uint(amax));
    }
    if (blockIdx.x == 0 && tid == 0) tile_ptr[(int64_t)n_rb * n_cb] = nnz;
}

template <typename T>
void* tc_build(salg_ctx* ctx, const salg_csr* c) {
    cudaStream_t st = ctx->stream;
    SALG_REQUIRE(c->nnz < ((int64_t)1 << 31), SALG_ERR_UNSUPPORTED, "tile format supports < 2^31 stored entries per GPU shard");
    TcTiles* t = new TcTiles();
    try {
        t->n_rb = (int)(ceil_div(ceil_div(c->nrows, TC_RB), TC_RPAD) * TC_RPAD);
        t->n_cb = (int)ceil_div(c->ncols, TC_CB);
        t->nnz = c->nnz;
        int64_t n_tiles = (int64_t)t->n_rb * t->n_cb;
        SALG_REQUIRE(n_tiles < ((int64_t)1 << 31), SALG_ERR_UNSUPPORTED, "too many tiles");
        ProfScope ps(ctx, PROF_TRANSPOSE, (double)c->nnz * (sizeof(T) + 4 + 8));
        t->entries = (uint2*)dev_alloc(ctx, ((size_t)c->nnz + 1024) * sizeof(uint2));
        t->tile_ptr = (int64_t*)dev_alloc(ctx, (size_t)(n_tiles + 2) * 8);
        SALG_CUDA(cudaMemsetAsync(t->entries + c->nnz, 0, 1024 * sizeof(uint2), st));
        DevBuf<unsigned> info(2, st);
        SALG_CUDA(cudaMemsetAsync(info.get(), 0, 8, st));
        int64_t nnz = c->nnz;
        const bool binned = t->n_cb <= TC_BIN_MAX_CB && !getenv("SALG_TC_SORT_BUILD");
        if (binned) {
            const size_t sm = (size_t)t->n_cb * sizeof(unsigned);
            int grid = t->n_rb < ctx->sm_count * 4 ? t->n_rb : ctx->sm_count * 4;
            tc_bin_kernel<T><<<grid, TC_BIN_THREADS, sm, st>>>(c->row_ptr, c->col, (const T*)c->val, c->nrows, nnz, t->n_rb,
                                                              t->n_cb, t->entries, t->tile_ptr, info.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        DevBuf<uint32_t> keys(binned ? 0 : (size_t)nnz + 1, st), keys_out(binned ? 0 : (size_t)nnz + 1, st);
        if (nnz && !binned) {
            DevBuf<uint2> payload((size_t)nnz, st);
            int64_t want = ceil_div(c->nrows * 32, 256);
            int64_t cap = (int64_t)ctx->sm_count * 16;
            tc_keys_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(
                c->row_ptr, c->col, (const T*)c->val, c->nrows, t->n_cb, keys.get(), payload.get(), info.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
            int end_bit = 1;
            while (end_bit < 32 && ((int64_t)1 << end_bit) < n_tiles) end_bit++;
            size_t tmp_bytes = 0;
            SALG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t*)keys.get(), keys_out.get(),
                                                      (const uint64_t*)payload.get(), (uint64_t*)t->entries, (int)nnz, 0,
                                                      end_bit, st));
            DevBuf<uint8_t> tmp(tmp_bytes + 16, st);
            SALG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tmp_bytes, (const uint32_t*)keys.get(), keys_out.get(),
                                                      (const uint64_t*)payload.get(), (uint64_t*)t->entries, (int)nnz, 0,
                                                      end_bit, st));
        }
        if (!binned) {
            tc_tile_ptr_kernel<<<(unsigned)ceil_div(n_tiles + 1, 256), 256, 0, st>>>(keys_out.get(), nnz, n_tiles, t->tile_ptr);
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        unsigned h_info[2] = {0, 0};
        SALG_CUDA(cudaMemcpyAsync(h_info, info.get(), 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        float amax;
        memcpy(&amax, &h_info[1], 4);
        if (h_info[0]) {
            t->a_terms = 2;
            t->a_scale = tc_pow2_scale(amax);
        } else {
            t->a_terms = 1;
            t->a_scale = 1.f;
        }
    } catch (...) {
        cudaStreamSynchronize(st);
        tc_free(ctx, t);
        throw;
    }
    return t;
}
template void* tc_build<float>(salg_ctx*, const salg_csr*);

// ---- panel pre-split into the canonical dense-operand layout -------------------------------------------------------
__global__ void tc_absmax_kernel(const float* __restrict__ P, int64_t n_elems, unsigned* __restrict__ out_bits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(P[i]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}
// scales[0] = panel scale s, scales[1] = 1 / (s * a_scale)
__global__ void tc_scale_kernel(const unsigned* __restrict__ amax_bits, float a_scale, float* __restrict__ scales) {
    float s = tc_pow2_scale(__uint_as_float(*amax_bits));
    scales[0] = s;
    scales[1] = 1.f / (s * a_scale);
}
// Panel P (n x 64 f32, row-major; row index = K of the product).  Block b covers K rows [b*KB, (b+1)*KB);
// out[b] = canonical (M = 128: m = 2*column + term) x (K = KB) fp16 operand, 16 B K-chunks 2048 B apart.
// One CTA converts 16 panel rows: coalesced 256 B row loads into shared memory, then thread m gathers the 8
// K-values of its operand row and writes one 16 B chunk (128 threads x 16 B = one contiguous 2 KB K-chunk).
template <int KB>
__global__ void __launch_bounds__(256)
tc_prep_kernel(const float* __restrict__ P, int64_t n, int64_t n_blocks, const float* __restrict__ scales,
               unsigned short* __restrict__ out) {
    __shared__ float tile[16][LP + 1];
    const float s = scales[0];
    const int64_t k0 = (int64_t)blockIdx.x * 16;
    for (int i = threadIdx.x; i < 16 * LP; i += 256) {
        int r = i >> 6, c = i & 63;
        int64_t k = k0 + r;
        tile[r][c] = (k < n) ? P[k * LP + c] * s : 0.f;
    }
    __syncthreads();
    const int half = threadIdx.x >> 7;          // which 8-row K-chunk of the 16 rows
    const int m = threadIdx.x & 127;            // operand row: 2 * column + term
    const int pc = m >> 1, t = m & 1;
    const int64_t kc0 = k0 + half * 8;          // first K of this chunk
    const int64_t b = kc0 / KB;
    if (b >= n_blocks) return;
    const uint32_t kl = (uint32_t)(kc0 % KB);
    unsigned short h[8];
#pragma unroll
    for (int j = 0; j < 8; j++) h[j] = f16_term(tile[half * 8 + j][pc], t);
    uint4 v;
    v.x = h[0] | ((uint32_t)h[1] << 16);
    v.y = h[2] | ((uint32_t)h[3] << 16);
    v.z = h[4] | ((uint32_t)h[5] << 16);
    v.w = h[6] | ((uint32_t)h[7] << 16);
    unsigned short* base = out + (size_t)b * (128 * KB);
    *reinterpret_cast<uint4*>(base + (canon_off((uint32_t)m, kl, 2048u) >> 1)) = v;
}

// ---- work sequences ---------------------------------------------------------------------------------------------------
// Every role of a CTA walks the same private sequence of units s = 0 .. n_units-1; a unit (one scatter + MMA pass
// per operator term) is a pair of tiles whose entry lists sit in ring slot s % NS.
//   A X   : row-block pairs g = blockIdx.x, +gridDim.x, ...; order (pair, cb); tiles (2g, cb) and (2g+1, cb)
//   A^T Y : row blocks [rb0, rb1), units u of each: tiles cb_lo + 2u and cb_lo + 2u + 1 (one tile when odd)
struct TcSeq {
    int n_cb;
    bool aty;
    int n_pairs_mine;                  // A X
    int rb0, rb1, cb_lo, ntr, upr;     // A^T Y (upr = units per row block)
    __device__ __forceinline__ int64_t n_units() const {
        return aty ? (int64_t)(rb1 - rb0) * upr : (int64_t)n_pairs_mine * n_cb;
    }
    // tile ids of unit s (t1 < 0: the unit has a single tile)
    __device__ __forceinline__ void tiles(int64_t s, int64_t& t0, int64_t& t1) const {
        if (!aty) {
            int64_t i = s / n_cb;
            int cb = (int)(s - i * n_cb) + (int)(blockIdx.x % (unsigned)n_cb);   // staggered start: the CTAs do not all
            if (cb >= n_cb) cb -= n_cb;                                          // pull the same panel slice from L2 at once
            int64_t g = (int64_t)blockIdx.x + i * gridDim.x;
            t0 = (2 * g) * n_cb + cb;
            t1 = (2 * g + 1) * n_cb + cb;
        } else {
            int64_t rbi = s / upr;
            int u = (int)(s - rbi * upr);
            t0 = (rb0 + rbi) * n_cb + cb_lo + 2 * u;
            t1 = (2 * u + 1 < ntr) ? t0 + 1 : -1;
        }
    }
};

struct alignas(16) TcSlotMeta { long long e0[2]; int n[2]; int pad[2]; };   // per tile: first entry, entry count, leading pad (0/1)
constexpr int TC_LOADER_WARPS = 4;   // entry loaders: warp w serves units s = w (mod 4), one lane each (a suspended
                                     // try_wait parks the whole warp, so lanes of one warp cannot wait independently)

// ---- entry loader role ---------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void tc_entry_loader(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr,
                                                const TcSeq& seq, uint8_t* sRing, TcSlotMeta* sMeta, uint64_t* e_full,
                                                uint64_t* e_free, int w) {
    const int64_t nu = seq.n_units();
    long long p[4] = {0, 0, 0, 0};       // tile pointers of the current unit: [e0, e1) of tile 0, [e0, e1) of tile 1
    bool two = false;
    auto fetch = [&](int64_t s, long long (&o)[4], bool& o_two) {
        int64_t t0, t1;
        seq.tiles(s, t0, t1);
        o[0] = tile_ptr[t0];
        o[1] = tile_ptr[t0 + 1];
        o_two = t1 >= 0;
        o[2] = o_two ? tile_ptr[t1] : 0;
        o[3] = o_two ? tile_ptr[t1 + 1] : 0;
    };
    if (w < nu) fetch(w, p, two);
    for (int64_t s = w; s < nu; s += TC_LOADER_WARPS) {
        long long np[4] = {0, 0, 0, 0};
        bool ntwo = false;
        if (s + TC_LOADER_WARPS < nu) fetch(s + TC_LOADER_WARPS, np, ntwo);   // next unit's pointers, overlapped
        const int slot = (int)(s % NS);
        const uint32_t use = (uint32_t)(s / NS);
        if (use > 0) mbar_wait(&e_free[slot], (use - 1) & 1);
        uint32_t bytes[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const long long e0 = p[2 * k], e1 = p[2 * k + 1];
            const int n = (k == 0 || two) ? (int)(e1 - e0) : 0;
            const int pad = (int)(e0 & 1);
            int cnt = n > 0 ? n + pad : 0;
            if (cnt > TC_SLOT_ENTRIES) cnt = TC_SLOT_ENTRIES;
            cnt = (cnt + 1) & ~1;
            sMeta[slot].e0[k] = e0;
            sMeta[slot].n[k] = n;
            sMeta[slot].pad[k] = pad;
            bytes[k] = (uint32_t)cnt * 8u;
        }
        if (bytes[0] + bytes[1] > 0) {
            mbar_expect_tx(&e_full[slot], bytes[0] + bytes[1]);
            uint8_t* dst = sRing + (size_t)slot * (2 * TC_SLOT_BYTES);
            if (bytes[0]) bulk_g2s(dst, entries + (p[0] - (p[0] & 1)), bytes[0], &e_full[slot]);
            if (bytes[1]) bulk_g2s(dst + TC_SLOT_BYTES, entries + (p[2] - (p[2] & 1)), bytes[1], &e_full[slot]);
        } else {
            mbar_arrive(&e_full[slot]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = np[k];
        two = ntwo;
    }
}

// ---- scatter role -----------------------------------------------------------------------------------------------------------
// Per pass: wait until the MMA that read this buffer has retired, clear it with 128-bit stores, barrier among the
// scatter warps, scatter the unit's entries as fp16 term `term`, make the writes visible to the tensor core
// (async proxy) and signal the MMA thread (one arrival per warp).  A thread owns at most two entries of each tile
// (denser tiles take the slow loop); all of them are loaded before any is processed so the shared-memory latencies
// overlap.
// Operand coordinates: A X   n = 128*k + local_row (k = tile of the pair), K = local column, chunk stride 4096
//                      A^T Y n =  64*k + local_col,                         K = local row,    chunk stride 2048
template <bool ATY>
__device__ __forceinline__ void tc_scatter_one(uint8_t* S, uint2 en, float a_scale, int term) {
    const uint32_t off = ((ATY ? (en.x >> 14) : en.x) & 0x3FFFu) << 1;      // pre-computed at tile-build time
    *reinterpret_cast<unsigned short*>(S + off) = f16_term(__uint_as_float(en.y) * a_scale, term);
}

template <bool ATY, int NS>
__device__ __forceinline__ void tc_scatter_role(const uint2* __restrict__ entries, int64_t n_units, int a_terms, float a_scale,
                                                uint8_t* sS, const uint8_t* sRing, const TcSlotMeta* sMeta, uint64_t* e_full,
                                                uint64_t* e_free, uint64_t* s_full, uint64_t* s_free, uint64_t* d_free, int units_per_d, int nb,
                                                int tid) {
    // d_free: the dense-operand stage of a finished group of `units_per_d` units is released here (by warp 0, when it
    // sees the s_free of the group's last pass) instead of by a second tcgen05.commit of the single MMA thread
    const int lane = tid & 31;
    uint32_t pass = 0;
    int slot = 0;
    uint32_t slot_use = 0;
    const uint32_t passes_per_d = (uint32_t)units_per_d * (uint32_t)a_terms;
    for (int64_t s = 0; s < n_units; s++) {
        const uint2* sl0 = reinterpret_cast<const uint2*>(sRing + (size_t)slot * (2 * TC_SLOT_BYTES));
        const uint2* sl1 = reinterpret_cast<const uint2*>(sRing + (size_t)slot * (2 * TC_SLOT_BYTES) + TC_SLOT_BYTES);
        for (int term = 0; term < a_terms; term++) {
            const int sb = pass % TC_NSB;
            const uint32_t use = pass / TC_NSB;
            if (lane == 0) {
                if (use > 0) {
                    mbar_wait(&s_free[sb], (use - 1) & 1);
                    // pass - TC_NSB has retired; if it closed a dense-operand group, hand that stage back to its loader
                    const uint32_t done = pass - TC_NSB;
                    if (tid == 0 && d_free && (done + 1) % passes_per_d == 0) mbar_arrive(&d_free[(done / passes_per_d) % nb]);
                }
                if (term == 0) mbar_wait(&e_full[slot], slot_use & 1);
            }
            __syncwarp();
            uint8_t* S = sS + sb * TC_S_BYTES;
#pragma unroll
            for (int i = 0; i < TC_S_BYTES / 16 / TC_SCATTER_THREADS; i++)
                reinterpret_cast<uint4*>(S)[i * TC_SCATTER_THREADS + tid] = make_uint4(0, 0, 0, 0);
            // this thread's entries of both tiles (independent shared-memory loads, issued before the barrier)
            const int4 mt = *reinterpret_cast<const int4*>(&sMeta[slot].n[0]);     // n0, n1, pad0, pad1
            const int n0 = mt.x, n1 = mt.y;
            constexpr int lim = TC_SLOT_ENTRIES - 1;     // entries guaranteed to sit in the slot whatever the pad
            const bool ok0 = tid < n0, ok1 = tid + TC_SCATTER_THREADS < n0 && tid + TC_SCATTER_THREADS < lim;
            const bool ok2 = tid < n1, ok3 = tid + TC_SCATTER_THREADS < n1 && tid + TC_SCATTER_THREADS < lim;
            uint2 en0 = ok0 ? sl0[mt.z + tid] : make_uint2(0, 0);
            uint2 en1 = ok1 ? sl0[mt.z + tid + TC_SCATTER_THREADS] : make_uint2(0, 0);
            uint2 en2 = ok2 ? sl1[mt.w + tid] : make_uint2(0, 0);
            uint2 en3 = ok3 ? sl1[mt.w + tid + TC_SCATTER_THREADS] : make_uint2(0, 0);
            named_bar_sync(1, TC_SCATTER_THREADS);       // every thread's clearing stores precede every scatter store
            if (ok0) tc_scatter_one<ATY>(S, en0, a_scale, term);
            if (ok1) tc_scatter_one<ATY>(S, en1, a_scale, term);
            if (ok2) tc_scatter_one<ATY>(S, en2, a_scale, term);
            if (ok3) tc_scatter_one<ATY>(S, en3, a_scale, term);
            if (n0 > lim || n1 > lim) {
                // rare: a tile with more entries than two per thread / than the slot holds
                for (int k = 0; k < 2; k++) {
                    const int n = k ? n1 : n0;
                    const uint2* sl = (k ? sl1 : sl0) + (k ? mt.w : mt.z);
                    const long long e0 = sMeta[slot].e0[k];
                    for (int i = tid + 2 * TC_SCATTER_THREADS; i < n; i += TC_SCATTER_THREADS)
                        if (i < lim) tc_scatter_one<ATY>(S, sl[i], a_scale, term);
                    for (int i = lim + tid; i < n; i += TC_SCATTER_THREADS) tc_scatter_one<ATY>(S, entries[e0 + i], a_scale, term);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_full[sb]);
                if (term == a_terms - 1) mbar_arrive(&e_free[slot]);   // every lane of this warp has read the slot
            }
            pass++;
        }
        if (++slot == NS) { slot = 0; slot_use++; }
    }
}

// ---- Y = A X - 1 corr^T -----------------------------------------------------------------------------------------------------
struct AxSmem {
    static constexpr int D_BYTES = 128 * TC_CB * 2;          // 16 KB per stage: panel slice (M = 128) x (K = 64)
    static constexpr int NB = 4;                             // >= TC_NSB + 2: a stage is released TC_NSB passes late
    static constexpr int NS = 8;                             // ring slots (one unit = two tiles each)
    static constexpr int TOTAL = TC_NSB * TC_S_BYTES + NB * D_BYTES + NS * 2 * TC_SLOT_BYTES + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_ax_kernel(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr, int n_rb, int n_cb, int a_terms,
             float a_scale, int64_t nrows, const uint8_t* __restrict__ Xprep, const float* __restrict__ scales,
             float* __restrict__ Y, const double* __restrict__ corr, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sS = smem;
    uint8_t* sD = sS + TC_NSB * TC_S_BYTES;
    uint8_t* sRing = sD + AxSmem::NB * AxSmem::D_BYTES;
    __shared__ uint64_t s_full[TC_NSB], s_free[TC_NSB], d_full[AxSmem::NB], d_free[AxSmem::NB], acc_full[2], acc_free[2];
    __shared__ uint64_t e_full[AxSmem::NS], e_free[AxSmem::NS];
    __shared__ TcSlotMeta sMeta[AxSmem::NS];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_pairs = n_rb / 2;
    const int n_mine = ((int)blockIdx.x < n_pairs) ? (n_pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid == 0) {
        for (int i = 0; i < TC_NSB; i++) {
            mbar_init(&s_full[i], TC_SCATTER_WARPS);
            mbar_init(&s_free[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 4);
        }
        for (int i = 0; i < AxSmem::NB; i++) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_free[i], 1);
        }
        for (int i = 0; i < AxSmem::NS; i++) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_free[i], TC_SCATTER_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    TcSeq seq{n_cb, false, n_mine, 0, 0, 0, 0, 0};

    if (warp < TC_SCATTER_WARPS) {
        tc_scatter_role<false, AxSmem::NS>(entries, seq.n_units(), a_terms, a_scale, sS, sRing, sMeta, e_full, e_free, s_full,
                                           s_free, nullptr, 1, 1, tid);
    } else if (warp >= TC_W_ELOAD) {
        if (lane == 0) tc_entry_loader<AxSmem::NS>(entries, tile_ptr, seq, sRing, sMeta, e_full, e_free, warp - TC_W_ELOAD);
    } else if (warp == TC_W_BLOAD) {
        // ================= panel-slice loader =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int gi = 0; gi < n_mine; gi++) {
                for (int cb = 0; cb < n_cb; cb++, it++) {
                    const int bb = it % AxSmem::NB;
                    const uint32_t use = it / AxSmem::NB;
                    if (use > 0) mbar_wait(&d_free[bb], (use - 1) & 1);
                    int cbr = cb + (int)(blockIdx.x % (unsigned)n_cb);      // same staggered order as TcSeq::tiles
                    if (cbr >= n_cb) cbr -= n_cb;
                    if (dbg & 4) { mbar_arrive(&d_full[bb]); continue; }
                    mbar_expect_tx(&d_full[bb], AxSmem::D_BYTES);
                    bulk_g2s(sD + bb * AxSmem::D_BYTES, Xprep + (size_t)cbr * AxSmem::D_BYTES, AxSmem::D_BYTES, &d_full[bb]);
                }
            }
        }
    } else if (warp == TC_W_MMA) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t pass = 0, it = 0;
            uint64_t s_desc[TC_NSB], d_desc[AxSmem::NB];
            for (int i = 0; i < TC_NSB; i++) s_desc[i] = umma_desc(smem_u32(sS + i * TC_S_BYTES), 4096, 128);
            for (int i = 0; i < AxSmem::NB; i++) d_desc[i] = umma_desc(smem_u32(sD + i * AxSmem::D_BYTES), 2048, 128);
            constexpr uint32_t idesc = tc_idesc(256);
            long long c_acc = 0, c_d = 0, c_s = 0, c_issue = 0, c_commit = 0, t_prev = clock64(), t_begin = t_prev;
            for (int gi = 0; gi < n_mine; gi++) {
                const int as = gi & 1;
                if (gi >= 2) mbar_wait(&acc_free[as], ((gi >> 1) - 1) & 1);
                tc_fence_after();
                TC_T(c_acc);
                const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
                for (int cb = 0; cb < n_cb; cb++, it++) {
                    const int bb = it % AxSmem::NB;
                    mbar_wait(&d_full[bb], (it / AxSmem::NB) & 1);
                    TC_T(c_d);
                    for (int term = 0; term < a_terms; term++) {
                        const int sb = pass % TC_NSB;
                        mbar_wait(&s_full[sb], (pass / TC_NSB) & 1);
                        tc_fence_after();
                        TC_T(c_s);
                        // K = 64: four K-steps; dense operand advances 2 chunks x 2048 B, sparse operand 2 x 4096 B
                        umma_f16_run4(d_tmem, d_desc[bb], s_desc[sb], idesc, (cb | term) != 0, 256, 512);
                        TC_T(c_issue);
                        umma_commit(&s_free[sb]);
                        TC_T(c_commit);
                        pass++;
                    }
                    umma_commit(&d_free[bb]);
                }
                umma_commit(&acc_full[as]);
            }
            if (blockIdx.x == 0) {
                g_tc_dbg[10] = clock64() - t_begin; g_tc_dbg[11] = c_acc; g_tc_dbg[12] = c_d; g_tc_dbg[13] = c_s; g_tc_dbg[14] = c_issue; g_tc_dbg[15] = c_commit;
            }
        }
    } else if (warp >= TC_W_EPI && warp < TC_W_EPI + 4) {
        // ================= epilogue warps =================
        // TMEM lane m = 2*column + term: the two term rows of a panel column sit in adjacent lanes of one warp
        const int qd = warp & 3;
        const int pc = qd * 16 + (lane >> 1);
        const float inv = scales[1];
        const float cr = corr ? (float)corr[pc] : 0.f;
        long long c_wait = 0, c_work = 0, t_prev = clock64();
        for (int gi = 0; gi < n_mine; gi++) {
            const int as = gi & 1;
            mbar_wait_warp(&acc_full[as], (gi >> 1) & 1, lane);
            tc_fence_after();
            TC_T(c_wait);
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)gi * gridDim.x) * 256;
#pragma unroll 1
            for (int c = 0; c < 8; c++) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)as * 256u + c * 32, v);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    float x = __uint_as_float(v[j]);
                    x += __shfl_xor_sync(0xFFFFFFFFu, x, 1);
                    const int64_t row = row0 + c * 32 + j;
                    if (!(lane & 1) && row < nrows) Y[row * LP + pc] = x * inv - cr;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[as]);
            TC_T(c_work);
        }
        if (blockIdx.x == 0 && warp == TC_W_EPI && lane == 0) { g_tc_dbg[16] = c_wait; g_tc_dbg[17] = c_work; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- Z += A^T Y (Z pre-initialised with the centring term) ------------------------------------------------------------------
struct AtySmem {
    static constexpr int D_BYTES = 128 * TC_RB * 2;          // 32 KB per stage: Y row block (M = 128) x (K = 128 rows)
    static constexpr int NB = 2;
    static constexpr int NS = 8;                             // ring slots (one unit = two tiles each)
    static constexpr int G = 8;                               // operator column blocks per CTA: 4 units x 128 TMEM columns
    static constexpr int TOTAL = TC_NSB * TC_S_BYTES + NB * D_BYTES + NS * 2 * TC_SLOT_BYTES + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_aty_kernel(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr, int n_rb, int n_cb, int a_terms,
              float a_scale, int64_t n_eff, const uint8_t* __restrict__ Yprep, const float* __restrict__ scales,
              float* __restrict__ Z, int n_groups, int rb_per_range) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sS = smem;
    uint8_t* sD = sS + TC_NSB * TC_S_BYTES;
    uint8_t* sRing = sD + AtySmem::NB * AtySmem::D_BYTES;
    __shared__ uint64_t s_full[TC_NSB], s_free[TC_NSB], d_full[AtySmem::NB], d_free[AtySmem::NB], acc_full;
    __shared__ uint64_t e_full[AtySmem::NS], e_free[AtySmem::NS];
    __shared__ TcSlotMeta sMeta[AtySmem::NS];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x % n_groups, range = blockIdx.x / n_groups;
    const int rb0 = range * rb_per_range;
    const int rb1 = (rb0 + rb_per_range < n_rb) ? rb0 + rb_per_range : n_rb;
    const int cb_lo = g * AtySmem::G;
    const int cb_hi = (cb_lo + AtySmem::G < n_cb) ? cb_lo + AtySmem::G : n_cb;     // exclusive
    const int ntr = cb_hi - cb_lo;
    const int n_units = (ntr + 1) / 2;

    if (tid == 0) {
        for (int i = 0; i < TC_NSB; i++) {
            mbar_init(&s_full[i], TC_SCATTER_WARPS);
            mbar_init(&s_free[i], 1);
        }
        for (int i = 0; i < AtySmem::NB; i++) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_free[i], 1);
        }
        for (int i = 0; i < AtySmem::NS; i++) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_free[i], TC_SCATTER_WARPS);
        }
        mbar_init(&acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const bool active = rb0 < rb1 && ntr > 0;
    TcSeq seq{n_cb, true, 0, rb0, rb1, cb_lo, ntr, n_units};

    if (warp < TC_SCATTER_WARPS) {
        if (active)
            tc_scatter_role<true, AtySmem::NS>(entries, seq.n_units(), a_terms, a_scale, sS, sRing, sMeta, e_full, e_free, s_full,
                                               s_free, nullptr, 1, 1, tid);
    } else if (warp >= TC_W_ELOAD) {
        if (active && lane == 0)
            tc_entry_loader<AtySmem::NS>(entries, tile_ptr, seq, sRing, sMeta, e_full, e_free, warp - TC_W_ELOAD);
    } else if (warp == TC_W_BLOAD) {
        if (lane == 0 && active) {
            uint32_t it = 0;
            for (int rb = rb0; rb < rb1; rb++, it++) {
                const int bb = it % AtySmem::NB;
                const uint32_t use = it / AtySmem::NB;
                if (use > 0) mbar_wait(&d_free[bb], (use - 1) & 1);
                mbar_expect_tx(&d_full[bb], AtySmem::D_BYTES);
                bulk_g2s(sD + bb * AtySmem::D_BYTES, Yprep + (size_t)rb * AtySmem::D_BYTES, AtySmem::D_BYTES, &d_full[bb]);
            }
        }
    } else if (warp == TC_W_MMA) {
        if (lane == 0 && active) {
            uint32_t pass = 0, it = 0;
            uint64_t s_desc[TC_NSB], d_desc[AtySmem::NB];
            for (int i = 0; i < TC_NSB; i++) s_desc[i] = umma_desc(smem_u32(sS + i * TC_S_BYTES), 2048, 128);
            for (int i = 0; i < AtySmem::NB; i++) d_desc[i] = umma_desc(smem_u32(sD + i * AtySmem::D_BYTES), 2048, 128);
            constexpr uint32_t idesc = tc_idesc(128);
            for (int rb = rb0; rb < rb1; rb++, it++) {
                const int bb = it % AtySmem::NB;
                mbar_wait(&d_full[bb], (it / AtySmem::NB) & 1);
                for (int u = 0; u < n_units; u++) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)u * 128u;
                    for (int term = 0; term < a_terms; term++) {
                        const int sb = pass % TC_NSB;
                        mbar_wait(&s_full[sb], (pass / TC_NSB) & 1);
                        tc_fence_after();
                        // K = 128 rows: eight K-steps, both operands advance 2 chunks x 2048 B per step
                        umma_f16_run4(d_tmem, d_desc[bb], s_desc[sb], idesc, ((rb - rb0) | term) != 0, 256, 256);
                        umma_f16_run4(d_tmem, d_desc[bb] + 1024, s_desc[sb] + 1024, idesc, 1, 256, 256);
                        umma_commit(&s_free[sb]);
                        pass++;
                    }
                }
                umma_commit(&d_free[bb]);      // once per row block (amortised over the block's units)
            }
            umma_commit(&acc_full);
        }
    } else if (warp >= TC_W_EPI && warp < TC_W_EPI + 4) {
        if (active) {
            const int qd = warp & 3;
            const int pc = qd * 16 + (lane >> 1);
            const float inv = scales[1];
            mbar_wait_warp(&acc_full, 0, lane);
            tc_fence_after();
            for (int u = 0; u < n_units; u++) {
#pragma unroll 1
                for (int c = 0; c < 4; c++) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)u * 128u + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        float x = __uint_as_float(v[j]);
                        x += __shfl_xor_sync(0xFFFFFFFFu, x, 1);
                        const int64_t cA = (int64_t)(cb_lo + 2 * u) * TC_CB + c * 32 + j;   // operator column
                        if (!(lane & 1) && cA < n_eff) atomicAdd(Z + cA * LP + pc, x * inv);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// Z[r][j] = -mu[r] * cs[j]  (or 0)
__global__ void tc_init_z_kernel(float* __restrict__ Z, int64_t n_eff, const float* __restrict__ mu, const double* __restrict__ cs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_eff * LP) return;
    Z[i] = (mu && cs) ? (float)(-(double)mu[i >> 6] * cs[i & 63]) : 0.f;
}

bool tc_enabled(const salg_ctx* ctx) { return ctx->spmm_impl == 0; }

static TcTiles* tiles_of(salg_ctx* ctx, const salg_csr* c) {
    if (!c->tc) c->tc = tc_build<float>(ctx, c);
    return (TcTiles*)c->tc;
}

// scale + split the panel into the canonical dense operand (device-side scale: no host round trip)
template <int KB>
static void tc_prepare_panel(salg_ctx* ctx, const float* P, int64_t n, int64_t n_blocks, float a_scale, float* d_scales,
                             unsigned* d_amax, uint8_t* out) {
    cudaStream_t st = ctx->stream;
    SALG_CUDA(cudaMemsetAsync(d_amax, 0, 4, st));
    if (n > 0) {
        int64_t ne = n * LP;
        int64_t want = ceil_div(ne, 256 * 8);
        int64_t cap = (int64_t)ctx->sm_count * 8;
        tc_absmax_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, st>>>(P, ne, d_amax);
        ctx->n_launch++;
    }
    tc_scale_kernel<<<1, 1, 0, st>>>(d_amax, a_scale, d_scales);
    ctx->n_launch++;
    tc_prep_kernel<KB><<<(unsigned)ceil_div(n_blocks * KB, 16), 256, 0, st>>>(P, n, n_blocks, d_scales, (unsigned short*)out);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

static void tc_dbg_print(salg_ctx* ctx, const char* what) {
    const char* e = getenv("SALG_TC_DBG");
    if (!e || !(atoi(e) & 32)) return;
    unsigned long long h[32];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpyFromSymbol(h, g_tc_dbg, sizeof(h));
    fprintf(stderr, "[tc %s] scatter total %llu passes %llu: s_free %llu zero %llu bar %llu e_full %llu scatter %llu fence %llu arrive %llu e_free %llu\n",
            what, h[0], h[9], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
    fprintf(stderr, "[tc %s] mma total %llu: acc_free %llu d_full %llu s_full %llu issue %llu commit %llu | epilogue wait %llu work %llu\n", what,
            h[10], h[11], h[12], h[13], h[14], h[15], h[16], h[17]);
}

// Y (nrows x 64) = A X - 1 corr^T
void tc_spmm_A(salg_ctx* ctx, const salg_csr* c, const float* X, float* Y, const double* corr) {
    cudaStream_t st = ctx->stream;
    TcTiles* t = tiles_of(ctx, c);
    if (c->nrows == 0) return;
    double bytes = (double)c->nnz * 8 + (double)(c->nrows + 1) * 8 + (double)c->ncols * 60 * 4 + (double)c->nrows * 60 * 4;
    ProfScope ps(ctx, PROF_SPMM, bytes);   // includes the panel pre-split
    DevBuf<uint8_t> Xprep((size_t)t->n_cb * AxSmem::D_BYTES, st);
    DevBuf<float> scales(2, st);
    DevBuf<unsigned> amax(1, st);
    tc_prepare_panel<TC_CB>(ctx, X, c->ncols, t->n_cb, t->a_scale, scales.get(), amax.get(), Xprep.get());
    SALG_CUDA(cudaFuncSetAttribute(tc_ax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AxSmem::TOTAL));
    int n_pairs = t->n_rb / 2;
    int grid = n_pairs < ctx->sm_count ? n_pairs : ctx->sm_count;
    tc_ax_kernel<<<grid, TC_THREADS, AxSmem::TOTAL, st>>>(t->entries, t->tile_ptr, t->n_rb, t->n_cb, t->a_terms, t->a_scale,
                                                          c->nrows, Xprep.get(), scales.get(), Y, corr, getenv("SALG_TC_DBG") ? atoi(getenv("SALG_TC_DBG")) : 0);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    tc_dbg_print(ctx, "ax");
}

// Z (ncols x 64) = A^T Y - mu corr^T   (local rows only)
void tc_spmm_At(salg_ctx* ctx, const salg_csr* c, const float* Y, float* Z, const float* mu, const double* corr) {
    cudaStream_t st = ctx->stream;
    TcTiles* t = tiles_of(ctx, c);
    if (c->ncols == 0) return;
    double bytes = (double)c->nnz * 8 + (double)(c->nrows + 1) * 8 + (double)c->ncols * 60 * 4 + (double)c->nrows * 60 * 4;
    ProfScope ps(ctx, PROF_SPMMT, bytes);   // includes the Z initialisation and the panel pre-split
    tc_init_z_kernel<<<(unsigned)ceil_div(c->ncols * LP, 256), 256, 0, st>>>(Z, c->ncols, mu, corr);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    if (c->nrows == 0) return;
    DevBuf<uint8_t> Yprep((size_t)t->n_rb * AtySmem::D_BYTES, st);
    DevBuf<float> scales(2, st);
    DevBuf<unsigned> amax(1, st);
    tc_prepare_panel<TC_RB>(ctx, Y, c->nrows, t->n_rb, t->a_scale, scales.get(), amax.get(), Yprep.get());
    int n_groups = (int)ceil_div(t->n_cb, AtySmem::G);
    int n_rb_real = (int)ceil_div(c->nrows, TC_RB);
    int ranges = ctx->sm_count / n_groups;
    if (ranges < 1) ranges = 1;
    if (ranges > n_rb_real) ranges = n_rb_real;
    int rb_per_range = (int)ceil_div(n_rb_real, ranges);
    ranges = (int)ceil_div(n_rb_real, rb_per_range);
    SALG_CUDA(cudaFuncSetAttribute(tc_aty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AtySmem::TOTAL));
    tc_aty_kernel<<<n_groups * ranges, TC_THREADS, AtySmem::TOTAL, st>>>(t->entries, t->tile_ptr, n_rb_real, t->n_cb, t->a_terms,
                                                                         t->a_scale, c->ncols, Yprep.get(), scales.get(), Z,
                                                                         n_groups, rb_per_range);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

}  // namespace salg
