// tc.cu — tile-densified sparse x panel products on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Why: per stored entry a CUDA-core CSR kernel needs one distinct 256 B panel row from L1/L2;
// measured on B200 that gather caps the products of the randomized-SVD power iteration at 2-6 % of the HBM
// roofline (profiles/r01_v1_summary.md).  Here the operator is re-tiled once into 128-row x 64-column tiles
// (entry stream sorted by tile, 8 B per entry); a CTA scatters one tile at a time into a zeroed dense bf16
// tile in shared memory (UMMA canonical K-major layout, no swizzle) and contracts it with the matching
// 64 x 64 slice of the panel by tcgen05.mma, accumulating the whole row block in TMEM.  The panel slice is
// fetched ONCE per 128 rows x 64 columns instead of once per entry.
//
// Precision: bf16 x bf16 products are exact in f32 and accumulate in f32.  The panel is split into three
// bf16 terms (24 mantissa bits) and so is the operator unless every stored value is exactly representable
// in bf16 (raw counts <= 256), in which case one term suffices.  Products kept: (a1,x1..x3) for exact
// operators; (a1,x1..x3), (a2,x1..x2), (a3,x1) otherwise — relative error ~2^-22, f32-like.
//
// Roles inside a CTA (warp-specialised, all hand-offs through mbarriers):
//   warps 0-7  scatter: un-scatter the previous tile's positions, scatter the new tile, fence.proxy.async
//   warp  8    loader: cp.async.bulk of the pre-split panel slice (canonical layout) into a 2-3 stage ring
//   warp  9    one thread issues tcgen05.mma (M=128, N=64, K=16 per instruction) and tcgen05.commit
//   warps 10-13 epilogue: tcgen05.ld the f32 accumulator, apply the rank-1 centring term, store / atomically add
#include <cub/cub.cuh>
#include <cuda_bf16.h>

#include "common.cuh"

namespace salg {

constexpr int TC_RB = 128;          // tile rows
constexpr int TC_CB = 64;           // tile columns
constexpr int TC_SCATTER_WARPS = 16;
constexpr int TC_SCATTER_THREADS = TC_SCATTER_WARPS * 32;
constexpr int TC_THREADS = (TC_SCATTER_WARPS + 7) * 32;   // scatter warps + panel loader + mma + 4 epilogue + entry loader
constexpr int TC_W_BLOAD = TC_SCATTER_WARPS, TC_W_MMA = TC_SCATTER_WARPS + 1, TC_W_EPI = TC_SCATTER_WARPS + 2,
              TC_W_ELOAD = TC_SCATTER_WARPS + 6;
constexpr int TC_RMAX = 4;          // row blocks sharing one panel slice in the A X kernel
constexpr int TC_SLOT_ENTRIES = 768;  // entries per ring slot (one tile); denser tiles read their tail from global memory
constexpr int TC_SLOT_BYTES = (TC_SLOT_ENTRIES + 2) * 8;
constexpr uint32_t TC_SPIN_LIMIT = 1u << 24;

struct TcTiles {
    uint2* entries = nullptr;       // [nnz] .x = (local_row << 6) | local_col, .y = f32 bits of the value
    int64_t* tile_ptr = nullptr;    // [n_rb * n_cb + 1]
    int n_rb = 0, n_cb = 0;
    int a_terms = 3;                // 1 when every value is exact in bf16
    int64_t nnz = 0;
};

void tc_free(void* p) {
    TcTiles* t = (TcTiles*)p;
    if (!t) return;
    if (t->entries) cudaFree(t->entries);
    if (t->tile_ptr) cudaFree(t->tile_ptr);
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
__device__ unsigned long long g_tc_dbg[32];   // timing experiment counters of CTA 0
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
// One lane polls, the warp follows: 32x fewer try_wait instructions competing for issue slots with the
// single-thread MMA / loader roles.  `sleep_ns` > 0 backs the poll off for waits that are known to be long.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane, unsigned sleep_ns = 0) {
    if (lane == 0) {
        if (sleep_ns) {
            uint32_t addr = smem_u32(bar), ok = 0, spins = 0;
            while (true) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(addr), "r"(parity)
                    : "memory");
                if (ok) break;
                if (++spins > TC_SPIN_LIMIT) __trap();
                __nanosleep(sleep_ns);
            }
        } else {
            mbar_wait(bar, parity);
        }
    }
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
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> f32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// KSTEPS consecutive K-steps of one (operator term, panel term) product in ONE asm block: the issuing thread is the
// serial resource of the CTA, so per-MMA overhead is two 64-bit adds.  Descriptor start addresses advance by
// `a_step` / `b_step` (16 B units) per K-step; only the first MMA may overwrite the accumulator.
template <int KSTEPS>
__device__ __forceinline__ void umma_bf16_run(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
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
    if (KSTEPS == 8) {
        asm volatile(
            "{\n\t.reg .pred q;\n\t.reg .b64 da, db;\n\t"
            "setp.eq.b32 q, 0, 0;\n\t"
            "mad.lo.u64 da, %4, 4, %1;\n\tmad.lo.u64 db, %5, 4, %2;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
            "add.u64 da, da, %4;\n\tadd.u64 db, db, %5;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
            "add.u64 da, da, %4;\n\tadd.u64 db, db, %5;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
            "add.u64 da, da, %4;\n\tadd.u64 db, db, %5;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
            "}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "l"(a_step), "l"(b_step)
            : "memory");
    }
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B (128 B contiguous);
// LBO = byte distance between the two 16 B K-chunks of one instruction, SBO = distance between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D f32, A/B bf16, both K-major, N = 64, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

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

// bf16 term `t` (0,1,2) of a float: x ~ b0 + b1 + b2 with 8 significant bits each
__device__ __forceinline__ unsigned short bf16_term(float x, int t) {
    __nv_bfloat16 b = __float2bfloat16_rn(x);
    if (t > 0) {
        x -= __bfloat162float(b);
        b = __float2bfloat16_rn(x);
        if (t > 1) {
            x -= __bfloat162float(b);
            b = __float2bfloat16_rn(x);
        }
    }
    return __bfloat16_as_ushort(b);
}

// byte offset of element (mn, k) inside a canonical K-major no-swizzle operand whose 16 B K-chunks are
// `chunk_stride` bytes apart (8-row groups are 128 B apart)
__device__ __forceinline__ uint32_t canon_off(uint32_t mn, uint32_t k, uint32_t chunk_stride) {
    return (k >> 3) * chunk_stride + (mn >> 3) * 128u + (mn & 7u) * 16u + (k & 7u) * 2u;
}

// ---- tile format builder -----------------------------------------------------------------------------------------
template <typename T>
__global__ void tc_keys_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col, const T* __restrict__ val,
                               int64_t nrows, int n_cb, uint32_t* __restrict__ keys, uint2* __restrict__ payload,
                               int* __restrict__ inexact) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (int64_t r = w; r < nrows; r += nw) {
        int64_t s = ptr[r], e = ptr[r + 1];
        uint32_t rb = (uint32_t)(r / TC_RB), lr = (uint32_t)(r % TC_RB);
        for (int64_t p = s + lane; p < e; p += 32) {
            uint32_t c = col[p];
            float v = (float)val[p];
            keys[p] = rb * (uint32_t)n_cb + c / TC_CB;
            payload[p] = make_uint2((lr << 6) | (c % TC_CB), __float_as_uint(v));
            bad |= (__bfloat162float(__float2bfloat16_rn(v)) != v);
        }
    }
    if (bad) atomicOr(inexact, 1);
}

__global__ void tc_tile_ptr_kernel(const uint32_t* __restrict__ keys_sorted, int64_t nnz, int64_t n_tiles,
                                   int64_t* __restrict__ tile_ptr) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    // first position whose key >= t
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)keys_sorted[mid] < t) lo = mid + 1; else hi = mid;
    }
    tile_ptr[t] = lo;
}

template <typename T>
void* tc_build(salg_ctx* ctx, const salg_csr* c) {
    cudaStream_t st = ctx->stream;
    SALG_REQUIRE(c->nnz < ((int64_t)1 << 31), SALG_ERR_UNSUPPORTED, "tile format supports < 2^31 stored entries per GPU shard");
    TcTiles* t = new TcTiles();
    try {
        t->n_rb = (int)(ceil_div(ceil_div(c->nrows, TC_RB), TC_RMAX) * TC_RMAX);   // padded: the A X kernel walks groups of TC_RMAX row blocks
        t->n_cb = (int)ceil_div(c->ncols, TC_CB);
        t->nnz = c->nnz;
        int64_t n_tiles = (int64_t)t->n_rb * t->n_cb;
        SALG_REQUIRE(n_tiles < ((int64_t)1 << 31), SALG_ERR_UNSUPPORTED, "too many tiles");
        ProfScope ps(ctx, PROF_TRANSPOSE, (double)c->nnz * (sizeof(T) + 4 + 8));
        SALG_CUDA(cudaMalloc((void**)&t->entries, ((size_t)c->nnz + 1024) * sizeof(uint2)));
        SALG_CUDA(cudaMalloc((void**)&t->tile_ptr, (size_t)(n_tiles + 2) * 8));
        SALG_CUDA(cudaMemsetAsync(t->entries + c->nnz, 0, 1024 * sizeof(uint2), st));
        DevBuf<int> flag(1, st);
        SALG_CUDA(cudaMemsetAsync(flag.get(), 0, 4, st));
        int64_t nnz = c->nnz;
        DevBuf<uint32_t> keys((size_t)nnz + 1, st), keys_out((size_t)nnz + 1, st);
        if (nnz) {
            DevBuf<uint2> payload((size_t)nnz, st);
            int64_t want = ceil_div(c->nrows * 32, 256);
            int64_t cap = (int64_t)ctx->sm_count * 16;
            tc_keys_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(
                c->row_ptr, c->col, (const T*)c->val, c->nrows, t->n_cb, keys.get(), payload.get(), flag.get());
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
        tc_tile_ptr_kernel<<<(unsigned)ceil_div(n_tiles + 1, 256), 256, 0, st>>>(keys_out.get(), nnz, n_tiles, t->tile_ptr);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        int h_flag = 0;
        SALG_CUDA(cudaMemcpyAsync(&h_flag, flag.get(), 4, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        t->a_terms = h_flag ? 3 : 1;
    } catch (...) {
        cudaStreamSynchronize(st);
        tc_free(t);
        throw;
    }
    return t;
}
template void* tc_build<float>(salg_ctx*, const salg_csr*);

// ---- panel pre-split into the canonical B-operand layout ---------------------------------------------------------
// Panel P (n x 64 f32, row-major; row index = K of the product).  Block b covers K rows [b*KB, (b+1)*KB);
// out[b][term][canonical (N = 64 panel columns) x (K = KB)] bf16, 16 B K-chunks 1024 B apart.
template <int KB>
__global__ void tc_prep_kernel(const float* __restrict__ P, int64_t n, int64_t n_blocks, unsigned short* __restrict__ out) {
    // one thread per (k, 4 panel columns)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = n_blocks * KB * 16;
    if (i >= total) return;
    int nq = (int)(i & 15);
    int64_t k = i >> 4;
    int64_t b = k / KB;
    uint32_t kl = (uint32_t)(k % KB);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < n) v = *reinterpret_cast<const float4*>(P + k * LP + nq * 4);
    float x[4] = {v.x, v.y, v.z, v.w};
    unsigned short* base = out + (size_t)b * 3 * (KB * 64);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t off = canon_off((uint32_t)(nq * 4 + j), kl, 1024u) >> 1;
#pragma unroll
        for (int t = 0; t < 3; t++) base[(size_t)t * (KB * 64) + off] = bf16_term(x[j], t);
    }
}

// ---- work sequences ---------------------------------------------------------------------------------------------------
// Every role of a CTA walks the same private sequence of tiles q = 0 .. q_total-1 (ring slot q % NS holds the
// tile's entries) grouped into units (one scatter + MMA pass each).
//   A X   : groups of R row blocks gb = blockIdx.x, +gridDim.x, ...; order (group, cb, r); unit = one tile
//   A^T Y : row blocks [rb0, rb1), tiles cb_lo .. cb_hi-1 of each; unit = two adjacent tiles (one when odd)
struct TcSeq {
    int n_cb;
    bool aty;
    int R, n_groups_mine;          // A X
    int rb0, rb1, cb_lo, ntr;      // A^T Y
    __device__ __forceinline__ int64_t q_total() const {
        return aty ? (int64_t)(rb1 - rb0) * ntr : (int64_t)n_groups_mine * n_cb * R;
    }
    __device__ __forceinline__ int64_t tile_id(int64_t q) const {
        if (!aty) {
            int64_t per = (int64_t)n_cb * R;
            int64_t i = q / per;
            int rem = (int)(q - i * per);
            int cb = rem / R, r = rem - cb * R;
            int64_t rb = ((int64_t)blockIdx.x + i * gridDim.x) * R + r;
            return rb * n_cb + cb;
        }
        int64_t rbi = q / ntr;
        int c = (int)(q - rbi * ntr);
        return (rb0 + rbi) * n_cb + cb_lo + c;
    }
    __device__ __forceinline__ int64_t n_units() const {
        return aty ? (int64_t)(rb1 - rb0) * ((ntr + 1) / 2) : q_total();
    }
    // unit s -> first tile index and tile count
    __device__ __forceinline__ void unit(int64_t s, int64_t& q0, int& nt) const {
        if (!aty) { q0 = s; nt = 1; return; }
        int upr = (ntr + 1) / 2;
        int64_t rbi = s / upr;
        int u = (int)(s - rbi * upr);
        q0 = rbi * ntr + 2 * u;
        nt = (2 * u + 1 < ntr) ? 2 : 1;
    }
};

struct TcSlotMeta { long long e0; int n; int pad; };   // first entry (global index), entry count, leading pad (0/1)

// ---- entry loader role: one warp streams the tiles' entry lists into the shared-memory ring ----------------------------
// Lane j owns ring slot j: the NS tiles of a batch are handled in parallel (pointer fetch, slot wait, one
// cp.async.bulk per lane), so the per-tile cost of this single warp is a few cycles and up to NS tiles
// (~NS * 4.6 KB) are in flight per SM without holding registers.  Pointers are fetched one batch ahead.
template <int NS>
__device__ __forceinline__ void tc_entry_loader(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr,
                                                const TcSeq& seq, uint8_t* sRing, TcSlotMeta* sMeta, uint64_t* e_full,
                                                uint64_t* e_free, int lane, int dbg) {
    const int64_t qt = seq.q_total();
    if (lane >= NS) return;
    long long e0 = 0, e1 = 0;
    if (lane < qt) {
        int64_t t = seq.tile_id(lane);
        e0 = tile_ptr[t];
        e1 = tile_ptr[t + 1];
    }
    uint32_t use = 0;
    for (int64_t qb = 0; qb < qt; qb += NS, use++) {
        const int64_t q = qb + lane;
        long long ne0 = 0, ne1 = 0;
        if (q + NS < qt) {
            int64_t t = seq.tile_id(q + NS);
            ne0 = tile_ptr[t];
            ne1 = tile_ptr[t + 1];
        }
        if (q < qt) {
            if (use > 0) mbar_wait(&e_free[lane], (use - 1) & 1);
            const int n = (int)(e1 - e0);
            const int pad = (int)(e0 & 1);
            int cnt = n + pad;
            if (cnt > TC_SLOT_ENTRIES) cnt = TC_SLOT_ENTRIES;
            cnt = (cnt + 1) & ~1;
            sMeta[lane].e0 = e0;
            sMeta[lane].n = n;
            sMeta[lane].pad = pad;
            if (cnt > 0 && !(dbg & 128)) {
                mbar_expect_tx(&e_full[lane], (uint32_t)cnt * 8u);
                bulk_g2s(sRing + (size_t)lane * TC_SLOT_BYTES, entries + (e0 - pad), (uint32_t)cnt * 8u, &e_full[lane]);
            } else {
                mbar_arrive(&e_full[lane]);
            }
        }
        e0 = ne0;
        e1 = ne1;
    }
}

// ---- scatter role -----------------------------------------------------------------------------------------------------------
// Per pass: wait until the MMA that read this A buffer has retired, clear the buffer with 128-bit stores (cheaper in
// instructions than undoing the previous scatter), barrier among the scatter warps, scatter the unit's entries as
// bf16 term `term`, make the writes visible to the tensor core (async proxy) and signal the MMA thread.
template <bool ATY, int A_BYTES, int NS>
__device__ __forceinline__ void tc_scatter_role(const uint2* __restrict__ entries, const TcSeq& seq, int a_terms, uint8_t* sA,
                                                const uint8_t* sRing, const TcSlotMeta* sMeta, uint64_t* e_full, uint64_t* e_free,
                                                uint64_t* a_full, uint64_t* a_free, int tid, int dbg) {
    const int64_t nu = seq.n_units();
    const int lane = tid & 31;
    uint32_t pass = 0;
    long long c_unit = 0, c_afree = 0, c_zero = 0, c_efull = 0, c_scat = 0, c_arr = 0, c_efree = 0, t_prev = clock64(), t_begin = t_prev;
    for (int64_t s = 0; s < nu; s++) {
        int64_t q0;
        int nt;
        seq.unit(s, q0, nt);
        TC_T(c_unit);
        for (int term = 0; term < a_terms; term++) {
            const int ab = pass & 1;
            const uint32_t use = pass >> 1;
            if (use > 0) mbar_wait_warp(&a_free[ab], (use - 1) & 1, lane);
            TC_T(c_afree);
            uint8_t* A = sA + ab * A_BYTES;
            if (!(dbg & 512)) {
#pragma unroll
                for (int i = 0; i < A_BYTES / 16 / TC_SCATTER_THREADS; i++)
                    reinterpret_cast<uint4*>(A)[i * TC_SCATTER_THREADS + tid] = make_uint4(0, 0, 0, 0);
                named_bar_sync(1, TC_SCATTER_THREADS);
            }
            TC_T(c_zero);
#pragma unroll
            for (int k = 0; k < 2; k++) {
                if (k < nt) {
                    const int64_t q = q0 + k;
                    const int slot = (int)(q % NS);
                    if (term == 0) mbar_wait_warp(&e_full[slot], (uint32_t)(q / NS) & 1, lane);
                    TC_T(c_efull);
                    const int n = (dbg & 1) ? 0 : sMeta[slot].n;
                    const int pad = sMeta[slot].pad;
                    const uint2* sl = reinterpret_cast<const uint2*>(sRing + (size_t)slot * TC_SLOT_BYTES) + pad;
                    const int in_slot = TC_SLOT_ENTRIES - pad;
                    for (int i = tid; i < n; i += TC_SCATTER_THREADS) {
                        uint2 en = (i < in_slot) ? sl[i] : entries[sMeta[slot].e0 + i];
                        uint32_t lr = en.x >> 6, lc = en.x & 63u;
                        uint32_t off = ATY ? canon_off(lc + (k ? 64u : 0u), lr, 2048u) : canon_off(lr, lc, 2048u);
                        *reinterpret_cast<unsigned short*>(A + off) = bf16_term(__uint_as_float(en.y), term);
                    }
                    TC_T(c_scat);
                }
            }
            if (!(dbg & 8)) fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[ab]);
            TC_T(c_arr);
            pass++;
        }
        // the unit's ring slots can be refilled once every lane of this warp has read them
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&e_free[(int)(q0 % NS)]);
            if (nt > 1) mbar_arrive(&e_free[(int)((q0 + 1) % NS)]);
        }
        TC_T(c_efree);
    }
    if ((dbg & 32) && blockIdx.x == 0 && tid == 0) {
        g_tc_dbg[0] = clock64() - t_begin; g_tc_dbg[1] = c_unit; g_tc_dbg[2] = c_afree; g_tc_dbg[3] = c_zero; g_tc_dbg[4] = c_efull;
        g_tc_dbg[5] = c_scat; g_tc_dbg[6] = c_arr; g_tc_dbg[7] = c_efree; g_tc_dbg[8] = pass;
    }
}

// ---- Y = A X - 1 corr^T -----------------------------------------------------------------------------------------------------
struct AxSmem {
    static constexpr int A_BYTES = TC_RB * TC_CB * 2;        // 16 KB per buffer (one bf16 term)
    static constexpr int B_BYTES = 3 * TC_CB * 64 * 2;       // 24 KB per stage (three terms)
    static constexpr int NB = 3;
    static constexpr int NS = 16;
    static constexpr int TOTAL = 2 * A_BYTES + NB * B_BYTES + NS * TC_SLOT_BYTES + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_ax_kernel(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr, int n_rb, int n_cb, int a_terms, int R,
             int64_t nrows, const uint8_t* __restrict__ Xprep, float* __restrict__ Y, const double* __restrict__ corr, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + 2 * AxSmem::A_BYTES;
    uint8_t* sRing = sB + AxSmem::NB * AxSmem::B_BYTES;
    __shared__ uint64_t a_full[2], a_free[2], b_full[AxSmem::NB], b_free[AxSmem::NB], acc_full[2], acc_free[2];
    __shared__ uint64_t e_full[AxSmem::NS], e_free[AxSmem::NS];
    __shared__ TcSlotMeta sMeta[AxSmem::NS];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_groups = n_rb / R;
    const int n_mine = ((int)blockIdx.x < n_groups) ? (n_groups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t tmem_cols = 2u * (uint32_t)R * 64u;
    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], TC_SCATTER_WARPS);   // one arrival per scatter warp: per-thread arrivals serialise on the barrier word
            mbar_init(&a_free[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 4);
        }
        for (int i = 0; i < AxSmem::NB; i++) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_free[i], 1);
        }
        for (int i = 0; i < AxSmem::NS; i++) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_free[i], TC_SCATTER_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, tmem_cols);
    for (int i = tid; i < 2 * AxSmem::A_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    TcSeq seq{n_cb, false, R, n_mine, 0, 0, 0, 0};

    if (warp < TC_SCATTER_WARPS) {
        tc_scatter_role<false, AxSmem::A_BYTES, AxSmem::NS>(entries, seq, a_terms, sA, sRing, sMeta, e_full, e_free, a_full,
                                                            a_free, tid, dbg);
    } else if (warp == TC_W_ELOAD) {
        tc_entry_loader<AxSmem::NS>(entries, tile_ptr, seq, sRing, sMeta, e_full, e_free, lane, dbg);
    } else if (warp == TC_W_BLOAD) {
        // ================= panel-slice loader =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int gi = 0; gi < n_mine; gi++) {
                for (int cb = 0; cb < n_cb; cb++, it++) {
                    const int bb = it % AxSmem::NB;
                    const uint32_t use = it / AxSmem::NB;
                    if (use > 0) mbar_wait(&b_free[bb], (use - 1) & 1);
                    if (dbg & 4) { mbar_arrive(&b_full[bb]); continue; }
                    mbar_expect_tx(&b_full[bb], AxSmem::B_BYTES);
                    bulk_g2s(sB + bb * AxSmem::B_BYTES, Xprep + (size_t)cb * AxSmem::B_BYTES, AxSmem::B_BYTES, &b_full[bb]);
                }
            }
        }
    } else if (warp == TC_W_MMA) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t pass = 0, it = 0;
            long long c_acc = 0, c_b = 0, c_a = 0, c_issue = 0, c_commit = 0, t_prev = clock64(), t_begin = t_prev;
            uint64_t a_desc[2], b_desc[AxSmem::NB];
            for (int i = 0; i < 2; i++) a_desc[i] = umma_desc(smem_u32(sA + i * AxSmem::A_BYTES), 2048, 128);
            for (int i = 0; i < AxSmem::NB; i++) b_desc[i] = umma_desc(smem_u32(sB + i * AxSmem::B_BYTES), 1024, 128);
            for (int gi = 0; gi < n_mine; gi++) {
                const int as = gi & 1;
                if (gi >= 2) mbar_wait(&acc_free[as], ((gi >> 1) - 1) & 1);
                tc_fence_after();
                TC_T(c_acc);
                for (int cb = 0; cb < n_cb; cb++, it++) {
                    const int bb = it % AxSmem::NB;
                    mbar_wait(&b_full[bb], (it / AxSmem::NB) & 1);
                    TC_T(c_b);
                    const uint64_t db0 = b_desc[bb];
                    for (int r = 0; r < R; r++) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(as * R + r) * 64u;
                        for (int term = 0; term < a_terms; term++) {
                            const int ab = pass & 1;
                            mbar_wait(&a_full[ab], (pass >> 1) & 1);
                            tc_fence_after();
                            TC_T(c_a);
                            const int nx = (a_terms == 1) ? 3 : 3 - term;
                            if (!(dbg & 2)) {
                                const uint64_t da0 = a_desc[ab];
                                umma_bf16_run<4>(d_tmem, da0, db0, TC_IDESC, (cb | term) != 0, 256, 128);
                                if (nx > 1) umma_bf16_run<4>(d_tmem, da0, db0 + 512, TC_IDESC, 1, 256, 128);
                                if (nx > 2) umma_bf16_run<4>(d_tmem, da0, db0 + 1024, TC_IDESC, 1, 256, 128);
                            }
                            TC_T(c_issue);
                            if (dbg & 16) mbar_arrive(&a_free[ab]); else umma_commit(&a_free[ab]);
                            TC_T(c_commit);
                            pass++;
                        }
                    }
                    if (dbg & 16) mbar_arrive(&b_free[bb]); else umma_commit(&b_free[bb]);
                }
                umma_commit(&acc_full[as]);
            }
            if ((dbg & 32) && blockIdx.x == 0) {
                g_tc_dbg[10] = clock64() - t_begin; g_tc_dbg[11] = c_acc; g_tc_dbg[12] = c_b; g_tc_dbg[13] = c_a; g_tc_dbg[14] = c_issue; g_tc_dbg[15] = c_commit;
            }
        }
    } else if (warp >= TC_W_EPI && warp < TC_W_EPI + 4) {
        // ================= epilogue warps =================
        const int q = warp & 3;
        for (int gi = 0; gi < n_mine; gi++) {
            const int as = gi & 1;
            mbar_wait_warp(&acc_full[as], (gi >> 1) & 1, lane, 256);
            tc_fence_after();
            const int64_t g = (int64_t)blockIdx.x + (int64_t)gi * gridDim.x;
            for (int r = 0; r < ((dbg & 256) ? 0 : R); r++) {
                const int64_t row = (g * R + r) * TC_RB + q * 32 + lane;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * R + r) * 64u + h * 32, v);
                    if (row < nrows) {
                        float4* o = reinterpret_cast<float4*>(Y + row * LP + h * 32);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 rr;
                            rr.x = __uint_as_float(v[4 * j + 0]) - (corr ? (float)corr[h * 32 + 4 * j + 0] : 0.f);
                            rr.y = __uint_as_float(v[4 * j + 1]) - (corr ? (float)corr[h * 32 + 4 * j + 1] : 0.f);
                            rr.z = __uint_as_float(v[4 * j + 2]) - (corr ? (float)corr[h * 32 + 4 * j + 2] : 0.f);
                            rr.w = __uint_as_float(v[4 * j + 3]) - (corr ? (float)corr[h * 32 + 4 * j + 3] : 0.f);
                            o[j] = rr;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---- Z += A^T Y (Z pre-initialised with the centring term) ------------------------------------------------------------------
struct AtySmem {
    static constexpr int A_BYTES = 128 * TC_RB * 2;          // 32 KB: M = 128 operator columns x K = 128 rows
    static constexpr int B_BYTES = 3 * 64 * TC_RB * 2;       // 48 KB: three terms of the Y row block
    static constexpr int NB = 2;
    static constexpr int NS = 10;
    static constexpr int TOTAL = 2 * A_BYTES + NB * B_BYTES + NS * TC_SLOT_BYTES + 1024;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_aty_kernel(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr, int n_rb, int n_cb, int a_terms,
              int64_t n_eff, const uint8_t* __restrict__ Yprep, float* __restrict__ Z, int G, int n_groups, int rb_per_range,
              uint32_t tmem_cols, int dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = sA + 2 * AtySmem::A_BYTES;
    uint8_t* sRing = sB + AtySmem::NB * AtySmem::B_BYTES;
    __shared__ uint64_t a_full[2], a_free[2], b_full[AtySmem::NB], b_free[AtySmem::NB], acc_full;
    __shared__ uint64_t e_full[AtySmem::NS], e_free[AtySmem::NS];
    __shared__ TcSlotMeta sMeta[AtySmem::NS];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x % n_groups, range = blockIdx.x / n_groups;
    const int rb0 = range * rb_per_range;
    const int rb1 = (rb0 + rb_per_range < n_rb) ? rb0 + rb_per_range : n_rb;
    const int cb_lo = g * G;
    const int cb_hi = (cb_lo + G < n_cb) ? cb_lo + G : n_cb;     // exclusive
    const int ntr = cb_hi - cb_lo;
    const int n_units = (ntr + 1) / 2;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a_full[i], TC_SCATTER_WARPS);   // one arrival per scatter warp: per-thread arrivals serialise on the barrier word
            mbar_init(&a_free[i], 1);
        }
        for (int i = 0; i < AtySmem::NB; i++) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_free[i], 1);
        }
        for (int i = 0; i < AtySmem::NS; i++) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_free[i], TC_SCATTER_WARPS);
        }
        mbar_init(&acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, tmem_cols);
    for (int i = tid; i < 2 * AtySmem::A_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const bool active = rb0 < rb1 && ntr > 0;
    TcSeq seq{n_cb, true, 0, 0, rb0, rb1, cb_lo, ntr};

    if (warp < TC_SCATTER_WARPS) {
        if (active)
            tc_scatter_role<true, AtySmem::A_BYTES, AtySmem::NS>(entries, seq, a_terms, sA, sRing, sMeta, e_full, e_free, a_full,
                                                                 a_free, tid, dbg);
    } else if (warp == TC_W_ELOAD) {
        if (active) tc_entry_loader<AtySmem::NS>(entries, tile_ptr, seq, sRing, sMeta, e_full, e_free, lane, dbg);
    } else if (warp == TC_W_BLOAD) {
        if (lane == 0 && active) {
            uint32_t it = 0;
            for (int rb = rb0; rb < rb1; rb++, it++) {
                const int bb = it % AtySmem::NB;
                const uint32_t use = it / AtySmem::NB;
                if (use > 0) mbar_wait(&b_free[bb], (use - 1) & 1);
                if (dbg & 4) { mbar_arrive(&b_full[bb]); continue; }
                mbar_expect_tx(&b_full[bb], AtySmem::B_BYTES);
                bulk_g2s(sB + bb * AtySmem::B_BYTES, Yprep + (size_t)rb * AtySmem::B_BYTES, AtySmem::B_BYTES, &b_full[bb]);
            }
        }
    } else if (warp == TC_W_MMA) {
        if (lane == 0 && active) {
            uint32_t pass = 0, it = 0;
            uint64_t a_desc[2], b_desc[AtySmem::NB];
            for (int i = 0; i < 2; i++) a_desc[i] = umma_desc(smem_u32(sA + i * AtySmem::A_BYTES), 2048, 128);
            for (int i = 0; i < AtySmem::NB; i++) b_desc[i] = umma_desc(smem_u32(sB + i * AtySmem::B_BYTES), 1024, 128);
            for (int rb = rb0; rb < rb1; rb++, it++) {
                const int bb = it % AtySmem::NB;
                mbar_wait(&b_full[bb], (it / AtySmem::NB) & 1);
                const uint64_t db0 = b_desc[bb];
                for (int u = 0; u < n_units; u++) {
                    const uint32_t d_tmem = tmem_base + u * 64;
                    for (int term = 0; term < a_terms; term++) {
                        const int ab = pass & 1;
                        mbar_wait(&a_full[ab], (pass >> 1) & 1);
                        tc_fence_after();
                        const int nx = (a_terms == 1) ? 3 : 3 - term;
                        if (!(dbg & 2)) {
                            const uint64_t da0 = a_desc[ab];
                            umma_bf16_run<8>(d_tmem, da0, db0, TC_IDESC, ((rb - rb0) | term) != 0, 256, 128);
                            if (nx > 1) umma_bf16_run<8>(d_tmem, da0, db0 + 1024, TC_IDESC, 1, 256, 128);
                            if (nx > 2) umma_bf16_run<8>(d_tmem, da0, db0 + 2048, TC_IDESC, 1, 256, 128);
                        }
                        if (dbg & 16) mbar_arrive(&a_free[ab]); else umma_commit(&a_free[ab]);
                        pass++;
                    }
                }
                if (dbg & 16) mbar_arrive(&b_free[bb]); else umma_commit(&b_free[bb]);
            }
            umma_commit(&acc_full);
        }
    } else if (warp >= TC_W_EPI && warp < TC_W_EPI + 4) {
        if (active) {
            const int q = warp & 3;
            mbar_wait_warp(&acc_full, 0, lane, 1024);
            tc_fence_after();
            for (int u = 0; u < n_units; u++) {
                const int64_t cA = (int64_t)(cb_lo + 2 * u) * TC_CB + q * 32 + lane;   // operator column of this TMEM lane
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + u * 64 + h * 32, v);
                    if (cA < n_eff && !(dbg & 2)) {
                        float* o = Z + cA * LP + h * 32;
#pragma unroll
                        for (int j = 0; j < 32; j++) atomicAdd(o + j, __uint_as_float(v[j]));
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// Z[r][j] = -mu[r] * cs[j]  (or 0)
__global__ void tc_init_z_kernel(float* __restrict__ Z, int64_t n_eff, const float* __restrict__ mu, const double* __restrict__ cs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_eff * LP) return;
    Z[i] = (mu && cs) ? (float)(-(double)mu[i >> 6] * cs[i & 63]) : 0.f;
}

bool tc_enabled(const salg_ctx* ctx) { return ctx->spmm_impl == 0; }

static int tc_dbg() {
    const char* e = getenv("SALG_TC_DBG");   // timing experiments only: 1 skip scatter, 2 skip MMA, 4 skip panel loads, 8 skip proxy fence
    return e ? atoi(e) : 0;
}

static void tc_dbg_print(salg_ctx* ctx, const char* what) {
    if (!(tc_dbg() & 32)) return;
    unsigned long long h[32];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpyFromSymbol(h, g_tc_dbg, sizeof(h));
    fprintf(stderr, "[tc %s] scatter total %llu passes %llu: unit %llu a_free %llu zero+bar %llu e_full %llu scatter %llu fence+arrive %llu e_free %llu\n",
            what, h[0], h[8], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    fprintf(stderr, "[tc %s] mma total %llu: acc_free %llu b_full %llu a_full %llu issue %llu commit %llu\n", what, h[10], h[11], h[12], h[13], h[14], h[15]);
}

static TcTiles* tiles_of(salg_ctx* ctx, const salg_csr* c) {
    if (!c->tc) c->tc = tc_build<float>(ctx, c);
    return (TcTiles*)c->tc;
}

// Y (nrows x 64) = A X - 1 corr^T
void tc_spmm_A(salg_ctx* ctx, const salg_csr* c, const float* X, float* Y, const double* corr) {
    cudaStream_t st = ctx->stream;
    TcTiles* t = tiles_of(ctx, c);
    if (c->nrows == 0) return;
    double bytes = (double)c->nnz * 8 + (double)(c->nrows + 1) * 8 + (double)c->ncols * 60 * 4 + (double)c->nrows * 60 * 4;
    ProfScope ps(ctx, PROF_SPMM, bytes);   // includes the panel pre-split
    DevBuf<uint8_t> Xprep((size_t)t->n_cb * AxSmem::B_BYTES, st);
    {
        int64_t total = (int64_t)t->n_cb * TC_CB * 16;
        tc_prep_kernel<TC_CB><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(X, c->ncols, t->n_cb, (unsigned short*)Xprep.get());
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    SALG_CUDA(cudaFuncSetAttribute(tc_ax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AxSmem::TOTAL));
    // R row blocks share one panel slice (cuts the L2 -> shared-memory panel traffic by R) as long as every SM
    // still gets several groups
    int R = TC_RMAX;
    while (R > 1 && t->n_rb / R < 4 * ctx->sm_count) R >>= 1;
    int n_groups = t->n_rb / R;
    int grid = n_groups < ctx->sm_count ? n_groups : ctx->sm_count;
    tc_ax_kernel<<<grid, TC_THREADS, AxSmem::TOTAL, st>>>(t->entries, t->tile_ptr, t->n_rb, t->n_cb, t->a_terms, R, c->nrows,
                                                          Xprep.get(), Y, corr, tc_dbg());
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
    DevBuf<uint8_t> Yprep((size_t)t->n_rb * AtySmem::B_BYTES, st);
    {
        int64_t total = (int64_t)t->n_rb * TC_RB * 16;
        tc_prep_kernel<TC_RB><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(Y, c->nrows, t->n_rb, (unsigned short*)Yprep.get());
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    int G = t->n_cb < 16 ? t->n_cb : 16;
    int n_groups = (int)ceil_div(t->n_cb, G);
    int n_rb_real = (int)ceil_div(c->nrows, TC_RB);
    int ranges = ctx->sm_count / n_groups;
    if (ranges < 1) ranges = 1;
    if (ranges > n_rb_real) ranges = n_rb_real;
    int rb_per_range = (int)ceil_div(n_rb_real, ranges);
    ranges = (int)ceil_div(n_rb_real, rb_per_range);
    int n_units = (G + 1) / 2;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)n_units * 64) tmem_cols <<= 1;
    SALG_CUDA(cudaFuncSetAttribute(tc_aty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AtySmem::TOTAL));
    tc_aty_kernel<<<n_groups * ranges, TC_THREADS, AtySmem::TOTAL, st>>>(t->entries, t->tile_ptr, n_rb_real, t->n_cb, t->a_terms,
                                                                         c->ncols, Yprep.get(), Z, G, n_groups, rb_per_range,
                                                                         tmem_cols, tc_dbg());
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

}  // namespace salg
