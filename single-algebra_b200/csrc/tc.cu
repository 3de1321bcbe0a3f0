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
#include <algorithm>
#include <cub/cub.cuh>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace salg {

constexpr int TC_RB = 128;          // tile rows
constexpr int TC_CB = 64;           // tile columns
// The scatter role is split into independent GROUPS: group g builds every TC_GROUPS-th unit while the other groups build
// theirs and the tensor core consumes earlier ones.  Buffers are NOT tied to groups: pass P (unit * terms + term) goes to
// buffer P % NSB, so a group can start its next unit while its previous one still waits for the tensor core.
// Measured on B200 (profiles/r02_tc_variants.md): 2 groups x 8 warps with 4 buffers (A X; two 16 KB panel stages instead
// of four pay for the fourth buffer) / 3 buffers (A^T Y) beat 3 x 5 with 3 buffers by 8 % / 8 %; one group of 15 or 16
// warps building unit after unit is 30 % slower (the units' hand-off chains no longer overlap).
#ifndef TC_GROUP_WARPS_
#define TC_GROUP_WARPS_ 8
#endif
#ifndef TC_UNSCATTER_
#define TC_UNSCATTER_ 0
#endif
#ifndef TC_GROUPS_
#define TC_GROUPS_ 2
#endif
#ifndef TC_SLOT_ENTRIES_
#define TC_SLOT_ENTRIES_ 768
#endif
#ifndef TC_NS_
#define TC_NS_ 5
#endif
#ifndef TC_AX_NB_
#define TC_AX_NB_ 2
#endif
#ifndef TC_ATY_NS_
#define TC_ATY_NS_ TC_NS_
#endif
constexpr int TC_GROUPS = TC_GROUPS_;
constexpr int TC_GROUP_WARPS = TC_GROUP_WARPS_;
constexpr bool TC_UNSCATTER = TC_UNSCATTER_ != 0;   // zero only what the previous unit wrote instead of clearing the buffer
constexpr int TC_GROUP_THREADS = TC_GROUP_WARPS * 32;
constexpr int TC_SCATTER_WARPS = TC_GROUPS * TC_GROUP_WARPS;
constexpr int TC_NS = TC_NS_;              // entry-ring slots (one unit = two tiles each)
constexpr int TC_LOADER_WARPS = TC_NS;  // entry loaders: warp w serves units s = w (mod TC_NS), i.e. ALWAYS slot w — a slot's
                                        // uses are then ordered by one thread, so its e_free parity wait can never be a
                                        // whole phase behind (a suspended try_wait parks the whole warp: one lane per warp)
constexpr int TC_THREADS = (TC_SCATTER_WARPS + 6 + TC_LOADER_WARPS) * 32;   // + panel loader, mma, 4 epilogue, entry loaders
constexpr int TC_W_BLOAD = TC_SCATTER_WARPS, TC_W_MMA = TC_SCATTER_WARPS + 1, TC_W_EPI = TC_SCATTER_WARPS + 2,
              TC_W_ELOAD = TC_SCATTER_WARPS + 6;
constexpr int TC_RPAD = 4;            // row blocks are padded to a multiple of this (the A X kernel walks pairs)
constexpr int TC_SLOT_ENTRIES = TC_SLOT_ENTRIES_;  // entries per ring slot (one tile); denser tiles read their tail from global memory
constexpr int TC_SLOT_BYTES = (TC_SLOT_ENTRIES + 2) * 8;
constexpr int TC_S_BYTES = 32768;     // sparse operand buffer: 256 x 64 (A X) or 128 x 128 (A^T Y) fp16
#ifndef TC_NSB_
#define TC_NSB_ 4
#endif
#ifndef TC_ATY_NSB_
#define TC_ATY_NSB_ 3
#endif
#ifndef TC_OCC_
#define TC_OCC_ 1
#endif
#ifndef TC_ATY_OCC_
#define TC_ATY_OCC_ 1
#endif
constexpr int TC_OCC = TC_OCC_;       // CTAs per SM (A X): 2 halves every per-CTA resource (TMEM 256 columns: one accumulator
                                      // stage) and lets two independent pipelines hide each other's hand-off latencies
constexpr int TC_ATY_OCC = TC_ATY_OCC_;   // the same for A^T Y (4 operator column blocks per CTA instead of 8)
constexpr int TC_ACC = TC_OCC == 1 ? 2 : 1;          // A X accumulator stages
constexpr int TC_TMEM_COLS = 512 / TC_OCC;
constexpr int TC_ATY_TMEM_COLS = 512 / TC_ATY_OCC;
constexpr int TC_NSB = TC_NSB_;       // sparse operand buffers (A X): pass P (unit * terms + term) uses buffer P % NSB
constexpr int TC_ATY_NSB = TC_ATY_NSB_;   // the same for A^T Y (its dense stages are twice as large)
constexpr int TC_EPT = (TC_SLOT_ENTRIES + TC_GROUP_THREADS - 1) / TC_GROUP_THREADS;   // ring entries per thread and tile

struct TcTiles {
    uint2* entries = nullptr;       // [nnz] .x = half byte-offsets in the A X (bits 0-13) and A^T Y (bits 14-27) operand
                                    //       buffers, .y = f32 bits of the value
    int64_t* tile_ptr = nullptr;    // [n_rb * n_cb + 1]
    int n_rb = 0, n_cb = 0;
    int a_terms = 2;                // 1 when every value is exact in fp16
    float a_scale = 1.f;            // power of two applied to the operator values before the split
    int64_t nnz = 0;
    void* tm = nullptr;             // TMEM-operand format (tm.cu) instead of entries / tile_ptr
};

// tm.cu
void tm_free(salg_ctx* owner, void* tiles);
bool tm_supported(const salg_csr* c);
void* tm_build(salg_ctx* ctx, const salg_csr* c, const int64_t* in_ptr, const uint32_t* in_col, const float* in_val, int in_shift);
int tm_n_rb(const void* tiles);
int tm_terms(const void* tiles);
float tm_scale(const void* tiles);
void tm_spmm_A(salg_ctx* ctx, const salg_csr* c, void* tiles, const float* X, float* Y, const double* corr, unsigned* d_amax, int b_terms);
void tm_aty_launch(salg_ctx* ctx, const salg_csr* c, void* tiles, const uint8_t* Yprep, const float* scales, float* Z);
void tm_spmm_A_prepped(salg_ctx* ctx, const salg_csr* c, void* tiles, const uint8_t* Xprep, const float* scales, float* Y,
                       const double* corr, unsigned* d_amax, int b_terms);
void tm_zside_apply(salg_ctx* ctx, const salg_csr* c, void* tiles, float* Z, const float* d_M, const float* mu, const float* scales,
                    uint8_t* Xprep, double* corr);
size_t tm_xprep_bytes(const void* tiles);
static bool tm_wanted(const salg_ctx* ctx, const salg_csr* c) { return ctx->spmm_impl == 2 && tm_supported(c); }

void tc_free(salg_ctx* owner, void* p) {
    TcTiles* t = (TcTiles*)p;
    if (!t) return;
    tm_free(owner, t->tm);
    dev_free(owner, t->entries);
    dev_free(owner, t->tile_ptr);
    delete t;
}

__device__ unsigned long long g_tc_dbg[32];   // timing experiment counters of CTA 0 (SALG_TC_DBG=32)
#define TC_T(acc) do { long long _t = clock64(); acc += _t - t_prev; t_prev = _t; } while (0)

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
__global__ void __launch_bounds__(TC_BIN_THREADS, 2)
tc_bin_kernel(const int64_t* __restrict__ in_ptr, int in_shift, const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col,
              const T* __restrict__ val, int64_t nrows,
              int64_t nnz, int n_rb, int n_cb, uint2* __restrict__ entries, int64_t* __restrict__ tile_ptr,
              unsigned* __restrict__ info /* [0] inexact flag, [1] bits of max |v| */, int rot) {
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
        // The warp's rows (r0 + warp + NW i): start and length fetched once by lanes i, broadcast per row — the sweeps
        // below never wait on a pointer load.  Each sweep reads a row in batches of BU x 32 entries with every load issued
        // before the first use (indices clamped to the row's last entry: no per-load predicates); one load per trip left
        // this kernel latency bound (1.29 ms for the cfg3 operator).
        constexpr int BU = 8;
        constexpr int RPW = TC_RB / NW;                        // rows per warp
        int64_t my_s = 0;
        uint32_t my_len = 0;
        // rows of this warp: interleaved (warp + NW i), or — `rot` — the 8 consecutive rows 8 warp .. 8 warp + 7 visited
        // in an order rotated by the warp index, so that warps running side by side place entries of DIFFERENT
        // (row & 7) classes into a tile (the shared-memory bank of an entry in the products is 4 (row & 7) + ...)
        auto local_row = [&](int i) -> unsigned { return rot ? (unsigned)(RPW * warp + ((i + warp) & (RPW - 1))) : (unsigned)(warp + NW * i); };
        if (lane < RPW) {
            const int64_t r = r0 + local_row(lane);
            if (r < r1) {
                const int64_t a0 = ptr[r];
                my_s = in_ptr[r] >> in_shift;
                my_len = (uint32_t)(ptr[r + 1] - a0);
            }
        }
        // sweep 1: column-block histogram (columns ascend inside a row: equal blocks sit in adjacent lanes)
        for (int i = 0; i < RPW; i++) {
            const int64_t s = __shfl_sync(0xFFFFFFFFu, my_s, i);
            const uint32_t len = __shfl_sync(0xFFFFFFFFu, my_len, i);
            const uint32_t* __restrict__ cr = col + s;
            const uint32_t last = len - 1;
            for (uint32_t p0 = 0; p0 < len; p0 += 32 * BU) {
                uint32_t c[BU];
#pragma unroll
                for (int u = 0; u < BU; u++) {
                    const uint32_t q = p0 + lane + 32 * u;
                    c[u] = cr[q < last ? q : last];
                }
#pragma unroll
                for (int u = 0; u < BU; u++) {
                    if (p0 + 32 * u >= len) break;            // warp-uniform
                    const bool ok = p0 + lane + 32 * u < len;
                    const unsigned cb = ok ? c[u] / TC_CB : 0xFFFFFFFFu;
                    const unsigned peers = __match_any_sync(0xFFFFFFFFu, cb);
                    if (ok && lane == __ffs(peers) - 1) atomicAdd(&hist[cb], (unsigned)__popc(peers));
                }
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
        for (int i = 0; i < RPW; i++) {
            const int64_t s = __shfl_sync(0xFFFFFFFFu, my_s, i);
            const uint32_t len = __shfl_sync(0xFFFFFFFFu, my_len, i);
            const uint32_t* __restrict__ cr = col + s;
            const T* __restrict__ vr = val + s;
            const uint32_t last = len - 1;
            const unsigned lr = local_row(i);
            for (uint32_t p0 = 0; p0 < len; p0 += 32 * BU) {
                uint32_t cc[BU];
                T vv[BU];
#pragma unroll
                for (int u = 0; u < BU; u++) {
                    uint32_t q = p0 + lane + 32 * u;
                    q = q < last ? q : last;
                    cc[u] = cr[q];
                    vv[u] = vr[q];
                }
#pragma unroll
                for (int u = 0; u < BU; u++) {
                    if (p0 + 32 * u >= len) break;            // warp-uniform
                    const bool ok = p0 + lane + 32 * u < len;
                    const unsigned c = cc[u];
                    const unsigned cb = ok ? c / TC_CB : 0xFFFFFFFFu;
                    const unsigned peers = __match_any_sync(0xFFFFFFFFu, cb);
                    const int leader = __ffs(peers) - 1;
                    unsigned start = 0;
                    if (ok && lane == leader) start = atomicAdd(&hist[cb], (unsigned)__popc(peers));
                    start = __shfl_sync(0xFFFFFFFFu, start, leader);
                    if (ok) {
                        const float v = (float)vv[u];
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
        }
        if (bad) atomicOr(&info[0], 1u);
        atomicMax(&info[1], __float_as_uint(amax));
    }
    if (blockIdx.x == 0 && tid == 0) tile_ptr[(int64_t)n_rb * n_cb] = nnz;
}

// ---- bank-aware order of the entries inside a tile --------------------------------------------------------------------
// The scatter warps take 32 CONSECUTIVE entries of a tile per store instruction.  The shared-memory bank of an entry is
// 4 (row & 7) + ((col & 7) >> 1) in the A X operand and 4 (col & 7) + ((row & 7) >> 1) in the A^T Y operand; in CSR
// arrival order (runs of one row) ncu counted 0.44 bank-conflict wavefronts per useful one, and shared-memory bandwidth is
// what bounds the products.  Entries of class (a, b) = (row & 7, col & 7) with a + b even ("colour" 0) map one-to-one to the
// 32 banks in BOTH operands (kappa = 4a + (b >> 1)), and so do the odd ones.  One warp per tile deals the entries out in
// rounds: round j of a colour holds the j-th entry of every class of that colour that has one, classes ascending, so 32
// consecutive entries almost always hit 32 different banks in both kernels.  The order inside a tile never affects results.
constexpr int TC_ORD_MAX = 1024;      // larger tiles keep their arrival order
constexpr int TC_ORD_WARPS = 4;
constexpr int TC_ORD_J = 64;          // rounds tracked per colour; entries beyond go to the tile's tail
__global__ void __launch_bounds__(TC_ORD_WARPS * 32)
tc_tile_order_kernel(uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr, int64_t n_tiles) {
    __shared__ uint2 s_in[TC_ORD_WARPS][TC_ORD_MAX];
    __shared__ unsigned char s_j[TC_ORD_WARPS][TC_ORD_MAX];
    __shared__ unsigned s_cnt[TC_ORD_WARPS][64], s_mask[TC_ORD_WARPS][2 * TC_ORD_J], s_start[TC_ORD_WARPS][2 * TC_ORD_J];
    __shared__ unsigned s_tail[TC_ORD_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint2* in = s_in[w];
    unsigned char* jb = s_j[w];
    unsigned *cnt = s_cnt[w], *mask = s_mask[w], *start = s_start[w];
    for (int64_t t = (int64_t)blockIdx.x * TC_ORD_WARPS + w; t < n_tiles; t += (int64_t)gridDim.x * TC_ORD_WARPS) {
        const int64_t e0 = tile_ptr[t];
        const int n = (int)(tile_ptr[t + 1] - e0);
        if (n < 64 || n > TC_ORD_MAX) continue;                 // warp-uniform
        __syncwarp();
        cnt[lane] = 0;
        cnt[lane + 32] = 0;
#pragma unroll
        for (int i = 0; i < 2 * TC_ORD_J / 32; i++) mask[lane + 32 * i] = 0;
        if (lane == 0) s_tail[w] = 0;
        for (int i = lane; i < n; i += 32) in[i] = entries[e0 + i];
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const unsigned x = in[i].x;                          // bits 0-2: col & 7, bits 3-5: row & 7 (A X half-offset)
            const unsigned a = (x >> 3) & 7u, b = x & 7u;
            const unsigned col = (a ^ b) & 1u, kap = a * 4u + (b >> 1);
            const unsigned j = atomicAdd(&cnt[col * 32 + kap], 1u);
            jb[i] = (unsigned char)(j < 255u ? j : 255u);
            if (j < TC_ORD_J) atomicOr(&mask[col * TC_ORD_J + j], 1u << kap);
        }
        __syncwarp();
        // start of round j of colour c = entries in all earlier rounds (colour 0 first)
        unsigned run = 0;
#pragma unroll
        for (int i = 0; i < 2 * TC_ORD_J / 32; i++) {
            const unsigned v = (unsigned)__popc(mask[lane + 32 * i]);
            unsigned incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += u;
            }
            start[lane + 32 * i] = run + incl - v;
            run += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        __syncwarp();
        const unsigned ordered = run;                            // entries with j < TC_ORD_J
        for (int i = lane; i < n; i += 32) {
            const uint2 e = in[i];
            const unsigned a = (e.x >> 3) & 7u, b = e.x & 7u;
            const unsigned col = (a ^ b) & 1u, kap = a * 4u + (b >> 1);
            const unsigned j = jb[i];
            unsigned pos;
            if (j < TC_ORD_J) pos = start[col * TC_ORD_J + j] + (unsigned)__popc(mask[col * TC_ORD_J + j] & ((1u << kap) - 1u));
            else pos = ordered + atomicAdd(&s_tail[w], 1u);
            entries[e0 + pos] = e;
        }
    }
}

template <typename T>
void* tc_build(salg_ctx* ctx, const salg_csr* c, const int64_t* in_ptr = nullptr, const uint32_t* in_col = nullptr,
               const T* in_val = nullptr, int in_shift = 0) {
    cudaStream_t st = ctx->stream;
    if (!in_ptr) {
        in_ptr = c->row_ptr;
        in_col = c->col;
        in_val = (const T*)c->val;
    }
    if (tm_wanted(ctx, c)) {
        TcTiles* t = new TcTiles();
        try {
            t->tm = tm_build(ctx, c, in_ptr, in_col, (const float*)in_val, in_shift);
        } catch (...) {
            delete t;
            throw;
        }
        t->n_rb = tm_n_rb(t->tm);
        t->n_cb = (int)ceil_div(c->ncols, TC_CB);
        t->nnz = c->nnz;
        t->a_terms = tm_terms(t->tm);
        t->a_scale = tm_scale(t->tm);
        return t;
    }
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
        const bool binned = in_ptr != c->row_ptr || (t->n_cb <= TC_BIN_MAX_CB && !getenv("SALG_TC_SORT_BUILD"));
        SALG_REQUIRE(t->n_cb <= TC_BIN_MAX_CB || in_ptr == c->row_ptr, SALG_ERR_UNSUPPORTED, "too many column blocks");
        if (binned) {
            const size_t sm = (size_t)t->n_cb * sizeof(unsigned);
            int grid = t->n_rb < ctx->sm_count * 4 ? t->n_rb : ctx->sm_count * 4;
            tc_bin_kernel<T><<<grid, TC_BIN_THREADS, sm, st>>>(in_ptr, in_shift, c->row_ptr, in_col, in_val, c->nrows, nnz, t->n_rb,
                                                              t->n_cb, t->entries, t->tile_ptr, info.get(),
                                                              getenv("SALG_TC_ROT") ? atoi(getenv("SALG_TC_ROT")) : 0);
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
        if (nnz && getenv("SALG_TC_ORDER")) {   // opt-in: measured +0.9 ms build for -2 % product time at cfg3
            int64_t want = ceil_div(n_tiles, TC_ORD_WARPS), cap = (int64_t)ctx->sm_count * 20;
            tc_tile_order_kernel<<<(unsigned)(want < cap ? want : cap), TC_ORD_WARPS * 32, 0, st>>>(t->entries, t->tile_ptr, n_tiles);
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
template void* tc_build<float>(salg_ctx*, const salg_csr*, const int64_t*, const uint32_t*, const float*, int);

void tc_attach_tiles_f32(salg_ctx* ctx, salg_csr* c, const int64_t* in_ptr, int in_shift, const uint32_t* col, const float* val) {
    if (c->tc) tc_free(ctx, c->tc);
    c->tc = tc_build<float>(ctx, c, in_ptr, col, val, in_shift);
}

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

// L2 prefetch of the entry lists of one unit (tile pointers o[0..3] as in tc_entry_loader)
__device__ __forceinline__ void tc_prefetch_unit(const uint2* __restrict__ entries, const long long (&o)[4], bool two) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (k == 1 && !two) break;
        const long long e0 = o[2 * k] & ~1LL, e1 = o[2 * k + 1];
        if (e1 > e0) {
            const uint32_t bytes = (uint32_t)(((e1 - e0) * 8 + 15) & ~15LL);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(entries + e0), "r"(bytes) : "memory");
        }
    }
}

// ---- entry loader role ---------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void tc_entry_loader(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr,
                                                const TcSeq& seq, uint8_t* sRing, TcSlotMeta* sMeta, uint64_t* e_full,
                                                uint64_t* e_free, int w, int pfd, int dbg = 0) {
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
    // L2 prefetch of the entry lists `pfd` units ahead: the ring holds NS units (~9 KB each), far less than the bytes a
    // B200 SM must keep in flight to cover DRAM latency at its share of the HBM bandwidth; with the lists already in L2
    // the ring's bulk copies complete in L2 latency instead
    long long pf[4] = {0, 0, 0, 0};
    bool pf_two = false;
    if (pfd > 0 && w + pfd < nu) fetch(w + pfd, pf, pf_two);
    if (w < nu) fetch(w, p, two);
    for (int64_t s = w; s < nu; s += NS) {           // loader w owns ring slot w
        long long np[4] = {0, 0, 0, 0};
        bool ntwo = false;
        if (s + NS < nu) fetch(s + NS, np, ntwo);   // next unit's pointers, overlapped
        if (pfd > 0) {
            if (s + pfd < nu) tc_prefetch_unit(entries, pf, pf_two);
            if (s + NS + pfd < nu) fetch(s + NS + pfd, pf, pf_two);
        }
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
        if (bytes[0] + bytes[1] > 0 && !(dbg & 64)) {                  // (dbg 64: timing experiment, no entry loads)
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

// the first `pfd` units of a CTA are prefetched by all lanes of the loader warps at kernel start
__device__ __forceinline__ void tc_prefetch_prologue(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr,
                                                     const TcSeq& seq, int pfd, int t) {
    const int64_t nu = seq.n_units();
    for (int64_t u = t; u < pfd && u < nu; u += TC_LOADER_WARPS * 32) {
        int64_t t0, t1;
        seq.tiles(u, t0, t1);
        long long o[4];
        o[0] = tile_ptr[t0];
        o[1] = tile_ptr[t0 + 1];
        const bool two = t1 >= 0;
        o[2] = two ? tile_ptr[t1] : 0;
        o[3] = two ? tile_ptr[t1 + 1] : 0;
        tc_prefetch_unit(entries, o, two);
    }
    __syncwarp();
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

template <bool ATY, int NS, int NSB>
__device__ __forceinline__ void tc_scatter_role(const uint2* __restrict__ entries, int64_t n_units, int a_terms, float a_scale,
                                                uint8_t* sS, const uint8_t* sRing, const TcSlotMeta* sMeta, uint64_t* e_full,
                                                uint64_t* e_free, uint64_t* s_full, uint64_t* s_free, int tid, int dbg = 0) {
    const int lane = tid & 31;
    const int grp = (tid >> 5) / TC_GROUP_WARPS;
    const int gt = tid - grp * TC_GROUP_THREADS;           // thread index inside the group
    uint32_t lp = 0;                                        // passes this group has built so far
    uint32_t prev[TC_EPT];                                  // half-offsets this thread wrote for the previous unit (tile 0 | tile 1 << 16)
#pragma unroll
    for (int k = 0; k < TC_EPT; k++) prev[k] = 0xFFFFFFFFu;
    bool full_clear = true;                                 // first unit, or the previous one had entries beyond the slot
    for (int64_t s = grp; s < n_units; s += TC_GROUPS) {
        const int slot = (int)(s % NS);
        const uint32_t slot_use = (uint32_t)(s / NS);
        const uint2* sl0 = reinterpret_cast<const uint2*>(sRing + (size_t)slot * (2 * TC_SLOT_BYTES));
        const uint2* sl1 = reinterpret_cast<const uint2*>(sRing + (size_t)slot * (2 * TC_SLOT_BYTES) + TC_SLOT_BYTES);
        for (int term = 0; term < a_terms; term++, lp++) {
            const uint32_t P = (uint32_t)s * (uint32_t)a_terms + (uint32_t)term;   // global pass index
            const uint32_t sb = P % NSB, sb_use = P / NSB;
            uint8_t* S = sS + sb * TC_S_BYTES;
            if (lane == 0) {
                if (sb_use > 0) mbar_wait_crit(&s_free[sb], (sb_use - 1) & 1);   // the MMAs of this buffer's previous pass have retired
                if (term == 0) mbar_wait(&e_full[slot], slot_use & 1);
            }
            __syncwarp();
            if (dbg & 1) {                                                   // (timing experiment: no clear)
            } else if (!TC_UNSCATTER || (term == 0 && full_clear)) {
                for (int i = gt; i < TC_S_BYTES / 16; i += TC_GROUP_THREADS) reinterpret_cast<uint4*>(S)[i] = make_uint4(0, 0, 0, 0);
            } else if (term == 0) {
#pragma unroll
                for (int k = 0; k < TC_EPT; k++) {
                    const uint32_t o0 = prev[k] & 0xFFFFu, o1 = prev[k] >> 16;
                    if (o0 != 0xFFFFu) *reinterpret_cast<unsigned short*>(S + (o0 << 1)) = 0;
                    if (o1 != 0xFFFFu) *reinterpret_cast<unsigned short*>(S + (o1 << 1)) = 0;
                }
            }   // (a second term overwrites exactly the first term's positions: nothing to clear)
            // this thread's entries of both tiles (independent shared-memory loads, issued before the barrier)
            const int4 mt = *reinterpret_cast<const int4*>(&sMeta[slot].n[0]);     // n0, n1, pad0, pad1
            const int n0 = mt.x, n1 = mt.y;
            constexpr int lim = TC_SLOT_ENTRIES - 1;     // entries guaranteed to sit in the slot whatever the pad
            const int m0 = n0 < lim ? n0 : lim, m1 = n1 < lim ? n1 : lim;
            uint2 en0[TC_EPT], en1[TC_EPT];
#pragma unroll
            for (int k = 0; k < TC_EPT; k++) {
                const int i = gt + k * TC_GROUP_THREADS;
                en0[k] = (i < m0) ? sl0[mt.z + i] : make_uint2(0, 0);
                en1[k] = (i < m1) ? sl1[mt.w + i] : make_uint2(0, 0);
            }
            if (!TC_UNSCATTER || term == 0)
                named_bar_sync(1 + grp, TC_GROUP_THREADS);   // every thread's clearing stores precede every scatter store
            if (!(dbg & 8)) {                                                // (timing experiment: no scatter stores)
#pragma unroll
            for (int k = 0; k < TC_EPT; k++) {
                const int i = gt + k * TC_GROUP_THREADS;
                if (i < m0) tc_scatter_one<ATY>(S, en0[k], a_scale, term);
                if (i < m1) tc_scatter_one<ATY>(S, en1[k], a_scale, term);
            }
            }
            if (n0 > lim || n1 > lim) {
                // rare: a tile with more entries than the slot holds reads its tail from global memory
                for (int k = 0; k < 2; k++) {
                    const int n = k ? n1 : n0;
                    const long long e0 = sMeta[slot].e0[k];
                    for (int i = lim + gt; i < n; i += TC_GROUP_THREADS) tc_scatter_one<ATY>(S, entries[e0 + i], a_scale, term);
                }
            }
            if (!(dbg & 16)) fence_proxy_async();                           // (timing experiment: no proxy fence)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_full[sb]);
                if (term == a_terms - 1) mbar_arrive(&e_free[slot]);   // every lane of this warp has read the slot
            }
            if (TC_UNSCATTER && term == a_terms - 1) {
                full_clear = n0 > lim || n1 > lim;
#pragma unroll
                for (int k = 0; k < TC_EPT; k++) {
                    const int i = gt + k * TC_GROUP_THREADS;
                    const uint32_t o0 = (i < m0) ? (((ATY ? (en0[k].x >> 14) : en0[k].x) & 0x3FFFu)) : 0xFFFFu;
                    const uint32_t o1 = (i < m1) ? (((ATY ? (en1[k].x >> 14) : en1[k].x) & 0x3FFFu)) : 0xFFFFu;
                    prev[k] = o0 | (o1 << 16);
                }
            }
        }
    }
}

// ---- Y = A X - 1 corr^T -----------------------------------------------------------------------------------------------------
struct AxSmem {
    static constexpr int D_BYTES = 128 * TC_CB * 2;          // 16 KB per stage: panel slice (M = 128) x (K = 64)
    static constexpr int NB = TC_AX_NB_;
    static constexpr int NS = TC_NS;                         // ring slots
    static constexpr int TOTAL = TC_NSB * TC_S_BYTES + NB * D_BYTES + NS * 2 * TC_SLOT_BYTES + 128;
};

__global__ void __launch_bounds__(TC_THREADS, TC_OCC)
tc_ax_kernel(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr, int n_rb, int n_cb, int a_terms,
             float a_scale, int64_t nrows, const uint8_t* __restrict__ Xprep, const float* __restrict__ scales,
             float* __restrict__ Y, const double* __restrict__ corr, unsigned* __restrict__ amax_out, int dbg, int pfd) {
    // (no integer round trip on this pointer: the compiler must keep seeing the shared address space, or every access
    // below becomes a generic LD/ST; the no-swizzle operand layouts only need 16 B alignment)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    uint8_t* sS = smem;
    uint8_t* sD = sS + TC_NSB * TC_S_BYTES;
    uint8_t* sRing = sD + AxSmem::NB * AxSmem::D_BYTES;
    __shared__ uint64_t s_full[TC_NSB], s_free[TC_NSB], d_full[AxSmem::NB], d_free[AxSmem::NB], acc_full[TC_ACC], acc_free[TC_ACC];
    __shared__ uint64_t e_full[AxSmem::NS], e_free[AxSmem::NS];
    __shared__ TcSlotMeta sMeta[AxSmem::NS];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_pairs = n_rb / 2;
    const int n_mine = ((int)blockIdx.x < n_pairs) ? (n_pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid == 0) {
        for (int i = 0; i < TC_NSB; i++) {
            mbar_init(&s_full[i], TC_GROUP_WARPS);
            mbar_init(&s_free[i], 1);
        }
        for (int i = 0; i < TC_ACC; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 4);
        }
        for (int i = 0; i < AxSmem::NB; i++) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_free[i], 1);
        }
        for (int i = 0; i < AxSmem::NS; i++) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_free[i], TC_GROUP_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    TcSeq seq{n_cb, false, n_mine, 0, 0, 0, 0, 0};

    if (warp < TC_SCATTER_WARPS) {
        tc_scatter_role<false, AxSmem::NS, TC_NSB>(entries, seq.n_units(), a_terms, a_scale, sS, sRing, sMeta, e_full, e_free, s_full,
                                           s_free, tid, dbg);
    } else if (warp >= TC_W_ELOAD) {
        tc_prefetch_prologue(entries, tile_ptr, seq, pfd, (warp - TC_W_ELOAD) * 32 + lane);
        if (lane == 0 && warp - TC_W_ELOAD < AxSmem::NS)
            tc_entry_loader<AxSmem::NS>(entries, tile_ptr, seq, sRing, sMeta, e_full, e_free, warp - TC_W_ELOAD, pfd, dbg);
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
            uint32_t it = 0;
            // descriptors of buffer i = descriptor of buffer 0 + i * (buffer bytes >> 4) (start-address field, no carry out)
            const uint64_t s_desc0 = umma_desc(smem_u32(sS), 4096, 128), d_desc0 = umma_desc(smem_u32(sD), 2048, 128);
            constexpr uint32_t idesc = tc_idesc(256);
            long long c_acc = 0, c_d = 0, c_s = 0, c_issue = 0, c_commit = 0, t_prev = clock64(), t_begin = t_prev;
            for (int gi = 0; gi < n_mine; gi++) {
                const int as = gi % TC_ACC;
                if (gi >= TC_ACC) mbar_wait(&acc_free[as], ((gi / TC_ACC) - 1) & 1);
                tc_fence_after();
                TC_T(c_acc);
                const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
                for (int cb = 0; cb < n_cb; cb++, it++) {
                    const int bb = it % AxSmem::NB;
                    mbar_wait(&d_full[bb], (it / AxSmem::NB) & 1);
                    TC_T(c_d);
                    for (int term = 0; term < a_terms; term++) {
                        const uint32_t P = it * (uint32_t)a_terms + (uint32_t)term;
                        const uint32_t sb = P % TC_NSB;
                        mbar_wait_crit(&s_full[sb], (P / TC_NSB) & 1);
                        tc_fence_after();
                        TC_T(c_s);
                        // K = 64: four K-steps; dense operand advances 2 chunks x 2048 B, sparse operand 2 x 4096 B
                        if (!(dbg & 2))                                       // (timing experiment: no MMA)
                        umma_f16_run4(d_tmem, d_desc0 + (uint64_t)bb * (AxSmem::D_BYTES >> 4), s_desc0 + (uint64_t)sb * (TC_S_BYTES >> 4), idesc,
                                      (cb | term) != 0, 256, 512);
                        TC_T(c_issue);
                        if (dbg & 128) mbar_arrive(&s_free[sb]);             // (timing experiment with bit 2: plain arrive, no commit)
                        else umma_commit(&s_free[sb]);
                        TC_T(c_commit);
                    }
                    if (dbg & 128) mbar_arrive(&d_free[bb]);
                    else umma_commit(&d_free[bb]);
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
        float amax = 0.f;                       // max |Y| of this thread's outputs (scale of the next pre-split)
        for (int gi = 0; gi < n_mine; gi++) {
            const int as = gi % TC_ACC;
            mbar_wait_warp(&acc_full[as], (gi / TC_ACC) & 1, lane);
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
                    if (!(lane & 1) && row < nrows) {
                        const float y = x * inv - cr;
                        Y[row * LP + pc] = y;
                        amax = fmaxf(amax, fabsf(y));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[as]);
            TC_T(c_work);
        }
        if (amax_out) {
#pragma unroll
            for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
            if (lane == 0 && amax > 0.f) atomicMax(amax_out, __float_as_uint(amax));
        }
        if (blockIdx.x == 0 && warp == TC_W_EPI && lane == 0) { g_tc_dbg[16] = c_wait; g_tc_dbg[17] = c_work; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ---- fused Gram + pre-split of a tall panel ---------------------------------------------------------------------------------
// One pass over Y (m x 64 f32): every 128-row block is fetched by one bulk copy, scaled and split into the canonical
// fp16 two-term operand (M = 128: m = 2*column + term, K = 128 rows) ONCE — the block is stored to global memory as the
// dense operand of the following A^T Y product and, from the same shared-memory copy, contracted with itself on the
// tensor core: D[m][n] += sum_k P[m][k] P[n][k] (M = N = 128, fp16 products exact, f32 accumulation in TMEM).  Every
// GP_DRAIN row blocks the accumulator is drained into f64 shared-memory sums (G[c][c'] = sum over the four term pairs),
// so the f32 accumulation never spans more than GP_DRAIN * 128 rows.  Replaces four passes over the panel (Gram, R^{-1}
// apply, |max|, pre-split) of the explicit CholeskyQR step: the power iteration applies R^{-1} to the SMALL side instead
// (A^T (Y R^{-1}) = (A^T Y) R^{-1}).  Column sums 1^T Y ride along in f64.
#ifndef GP_STAGES_
#define GP_STAGES_ 3
#endif
#ifndef GP_OPS_
#define GP_OPS_ 2
#endif
struct GpSmem {
    static constexpr int STAGES = GP_STAGES_;
    static constexpr int STAGE_BYTES = TC_RB * LP * 4;        // 32 KB of f32 rows
    static constexpr int OPS = GP_OPS_;
    static constexpr int OP_BYTES = 128 * TC_RB * 2;          // 32 KB canonical operand
    static constexpr int G_LD = LP + 1;
    static constexpr int G_BYTES = LP * G_LD * 8;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + OPS * OP_BYTES + G_BYTES + 128;
};
constexpr int GP_CONV_WARPS = 16;    // (8 warps left the pass latency-bound on the conversion chain: 0.20 ms per 1M-row panel)
constexpr int GP_W_EPI = GP_CONV_WARPS, GP_W_LOAD = GP_CONV_WARPS + 4, GP_W_MMA = GP_CONV_WARPS + 5,
              GP_W_STORE = GP_CONV_WARPS + 6;
constexpr int GP_THREADS = (GP_CONV_WARPS + 7) * 32;
constexpr int GP_DRAIN = 8;

__global__ void __launch_bounds__(GP_THREADS, 1)
tc_gram_prep_kernel(const float* __restrict__ Y, int64_t m, int n_rb, const float* __restrict__ scales,
                    uint8_t* __restrict__ Yprep, double* __restrict__ G /* GRAM_BUF, pre-zeroed */, int dbg) {
    // (no integer round trip on this pointer: the compiler must keep seeing the shared address space, or every access
    // below becomes a generic LD/ST; the no-swizzle operand layouts only need 16 B alignment)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    uint8_t* sStage = smem;
    uint8_t* sOp = sStage + GpSmem::STAGES * GpSmem::STAGE_BYTES;
    double* sG = reinterpret_cast<double*>(sOp + GpSmem::OPS * GpSmem::OP_BYTES);
    __shared__ uint64_t st_full[GpSmem::STAGES], st_free[GpSmem::STAGES], op_full[GpSmem::OPS], op_free[GpSmem::OPS],
        acc_full[2], acc_free[2];
    __shared__ double s_cs[GP_CONV_WARPS][LP];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_mine = ((int)blockIdx.x < n_rb) ? (n_rb - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_groups = (n_mine + GP_DRAIN - 1) / GP_DRAIN;
    if (tid == 0) {
        for (int i = 0; i < GpSmem::STAGES; i++) {
            mbar_init(&st_full[i], 1);
            mbar_init(&st_free[i], GP_CONV_WARPS);
        }
        for (int i = 0; i < GpSmem::OPS; i++) {
            mbar_init(&op_full[i], GP_CONV_WARPS);
            mbar_init(&op_free[i], 2);              // the MMAs' commit + the bulk store of the block
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 4);
        }
        fence_barrier_init();
    }
    for (int i = tid; i < LP * GpSmem::G_LD; i += GP_THREADS) sG[i] = 0.0;
    if (warp == 0) tmem_alloc(&s_tmem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp < GP_CONV_WARPS) {
        // ================= converters: f32 rows -> canonical fp16 two-term operand =================
        // One item per thread and block: K-octet o (8 rows) x column pair cp -> the four 16 B chunks of operand rows
        // m = 4cp .. 4cp+3 (column 2cp term 0/1, column 2cp+1 term 0/1), 64 B contiguous.  Both terms come from one load
        // and one scaling; conversions are done two values at a time (the pass is instruction-issue bound otherwise).
        static_assert(GP_CONV_WARPS == 16, "one (octet, column pair) item per thread");
        const int cp = tid & 31, o = tid >> 5;
        const float s = scales[0];
        double cs0 = 0.0, cs1 = 0.0;
        for (int it = 0; it < n_mine; it++) {
            const int64_t rb = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
            const int stg = it % GpSmem::STAGES, ob = it % GpSmem::OPS;
            if (lane == 0) {
                mbar_wait(&st_full[stg], (it / GpSmem::STAGES) & 1);
                if (it >= GpSmem::OPS) mbar_wait(&op_free[ob], ((it / GpSmem::OPS) - 1) & 1);
            }
            __syncwarp();
            const float2* src = reinterpret_cast<const float2*>(sStage + stg * GpSmem::STAGE_BYTES) + (size_t)(8 * o) * (LP / 2) + cp;
            const int rows_left = (int)((m - rb * TC_RB < TC_RB) ? m - rb * TC_RB : TC_RB) - 8 * o;   // valid rows of this octet
            float2 v[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) v[jj] = src[jj * (LP / 2)];
            if (rows_left < 8) {                      // last block of the panel: the stage's tail is stale
#pragma unroll
                for (int jj = 0; jj < 8; jj++)
                    if (jj >= rows_left) v[jj] = make_float2(0.f, 0.f);
            }
            float p0 = 0.f, p1 = 0.f;
            uint32_t w[4][4];                         // [operand row 4cp + r][packed pair of K values]
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float2 a = v[2 * q], b = v[2 * q + 1];
                p0 += a.x + b.x;
                p1 += a.y + b.y;
                const float ax = a.x * s, bx = b.x * s, ay = a.y * s, by = b.y * s;
                const __half2 hx = __floats2half2_rn(ax, bx), hy = __floats2half2_rn(ay, by);       // term 0
                const float2 fx = __half22float2(hx), fy = __half22float2(hy);
                const __half2 lx = __floats2half2_rn(ax - fx.x, bx - fx.y), ly = __floats2half2_rn(ay - fy.x, by - fy.y);   // term 1
                w[0][q] = *reinterpret_cast<const uint32_t*>(&hx);
                w[1][q] = *reinterpret_cast<const uint32_t*>(&lx);
                w[2][q] = *reinterpret_cast<const uint32_t*>(&hy);
                w[3][q] = *reinterpret_cast<const uint32_t*>(&ly);
            }
            cs0 += (double)p0;
            cs1 += (double)p1;
            uint8_t* dst = sOp + ob * GpSmem::OP_BYTES + (uint32_t)o * 2048u + (uint32_t)cp * 64u;   // canon_off(4cp, 8o, 2048)
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int rr = (r + cp) & 3;          // rotate the chunk order across lanes (bank spread of the 64 B stride)
                uint4 qv;
                qv.x = rr == 0 ? w[0][0] : rr == 1 ? w[1][0] : rr == 2 ? w[2][0] : w[3][0];
                qv.y = rr == 0 ? w[0][1] : rr == 1 ? w[1][1] : rr == 2 ? w[2][1] : w[3][1];
                qv.z = rr == 0 ? w[0][2] : rr == 1 ? w[1][2] : rr == 2 ? w[2][2] : w[3][2];
                qv.w = rr == 0 ? w[0][3] : rr == 1 ? w[1][3] : rr == 2 ? w[2][3] : w[3][3];
                *reinterpret_cast<uint4*>(dst + rr * 16) = qv;
            }
            if (!(dbg & 16)) fence_proxy_async();      // (timing experiment switch)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&op_full[ob]);
                mbar_arrive(&st_free[stg]);
            }
        }
        // column sums: the 16 converter warps of the CTA are summed in shared memory first.  One atomic per warp put 16 x 148
        // f64 atomics on each of only 64 addresses: same-address atomics serialise in L2, and that was the ~40 us of this
        // pass that did not shrink with the panel (0.052 ms for a 125k-row panel with loads, stores and MMAs switched off)
        s_cs[o][2 * cp] = cs0;
        s_cs[o][2 * cp + 1] = cs1;
        named_bar_sync(1, GP_CONV_WARPS * 32);
        if (o < 2) {
            double a = 0.0;
#pragma unroll
            for (int w = 0; w < GP_CONV_WARPS; w++) a += s_cs[w][o * 32 + cp];
            if (a != 0.0) atomicAdd(&G[LP * LP + o * 32 + cp], a);
        }
    } else if (warp == GP_W_LOAD) {
        if (lane == 0) {
            for (int it = 0; it < n_mine; it++) {
                const int64_t rb = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                const int stg = it % GpSmem::STAGES;
                if (it >= GpSmem::STAGES) mbar_wait(&st_free[stg], ((it / GpSmem::STAGES) - 1) & 1);
                int64_t rows = m - rb * TC_RB;
                if (rows > TC_RB) rows = TC_RB;
                const uint32_t bytes = (uint32_t)rows * LP * 4u;
                if (dbg & 4) { mbar_arrive(&st_full[stg]); continue; }      // timing experiment: no load
                mbar_expect_tx(&st_full[stg], bytes);
                bulk_g2s(sStage + stg * GpSmem::STAGE_BYTES, Y + rb * TC_RB * LP, bytes, &st_full[stg]);
            }
        }
    } else if (warp == GP_W_STORE) {
        // the finished operand block goes to global memory as ONE 32 KB bulk store (it is the dense operand of the next
        // A^T Y product); per-thread stores here would sit in front of the converters' proxy fence
        if (lane == 0) {
            for (int it = 0; it < n_mine; it++) {
                const int64_t rb = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
                const int ob = it % GpSmem::OPS;
                mbar_wait(&op_full[ob], (it / GpSmem::OPS) & 1);
                if (dbg & 1) { mbar_arrive(&op_free[ob]); continue; }      // timing experiment: no store
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Yprep + (size_t)rb * GpSmem::OP_BYTES),
                             "r"(smem_u32(sOp + ob * GpSmem::OP_BYTES)), "r"((uint32_t)GpSmem::OP_BYTES)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(&op_free[ob]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == GP_W_MMA) {
        if (lane == 0) {
            const uint64_t desc0 = umma_desc(smem_u32(sOp), 2048, 128);
            constexpr uint32_t idesc = tc_idesc(128);
            for (int it = 0; it < n_mine; it++) {
                const int ob = it % GpSmem::OPS;
                const int grp = it / GP_DRAIN, as = grp & 1;
                if (it % GP_DRAIN == 0 && grp >= 2) mbar_wait(&acc_free[as], ((grp >> 1) - 1) & 1);
                mbar_wait(&op_full[ob], (it / GpSmem::OPS) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)as * 128u;
                // K = 128 rows: eight K-steps of 16, both operands are the SAME buffer (P P^T)
                const uint64_t desc = desc0 + (uint64_t)ob * (GpSmem::OP_BYTES >> 4);
                if (!(dbg & (2 | 32))) {                                      // (2: timing experiment, 32: caller needs no Gram)
                    umma_f16_run4(d_tmem, desc, desc, idesc, (it % GP_DRAIN) != 0, 256, 256);
                    umma_f16_run4(d_tmem, desc + 1024, desc + 1024, idesc, 1, 256, 256);
                }
                umma_commit(&op_free[ob]);
                if ((it + 1) % GP_DRAIN == 0 || it + 1 == n_mine) umma_commit(&acc_full[as]);
            }
        }
    } else if (warp >= GP_W_EPI && warp < GP_W_EPI + 4) {
        // ================= drain: TMEM lane = operand row 2c + t; columns = operand rows 2c' + t' =================
        const int qd = warp & 3;
        const int c = (qd * 32 + lane) >> 1;
        for (int grp = 0; grp < n_groups; grp++) {
            const int as = grp & 1;
            mbar_wait_warp(&acc_full[as], (grp >> 1) & 1, lane);
            tc_fence_after();
#pragma unroll 1
            for (int c4 = 0; c4 < ((dbg & 32) ? 0 : 4); c4++) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)as * 128u + c4 * 32, v);
#pragma unroll
                for (int jj = 0; jj < 16; jj++) {
                    float x = __uint_as_float(v[2 * jj]) + __uint_as_float(v[2 * jj + 1]);
                    x += __shfl_xor_sync(0xFFFFFFFFu, x, 1);
                    if (!(lane & 1)) sG[c * GpSmem::G_LD + c4 * 16 + jj] += (double)x;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[as]);
        }
        if (n_mine > 0 && !(lane & 1) && !(dbg & 32)) {
            const double s = (double)scales[0];
            const double inv_s2 = 1.0 / (s * s);
            for (int cc = 0; cc < LP; cc++) {
                const double g = sG[c * GpSmem::G_LD + cc];
                if (g != 0.0) atomicAdd(&G[c * LP + cc], g * inv_s2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---- Z += A^T Y (Z pre-initialised with the centring term) ------------------------------------------------------------------
struct AtySmem {
    static constexpr int D_BYTES = 128 * TC_RB * 2;          // 32 KB per stage: Y row block (M = 128) x (K = 128 rows)
    static constexpr int NB = 2;
    static constexpr int NS = TC_ATY_NS_;                         // ring slots
    static constexpr int G = 8 / TC_ATY_OCC;                      // operator column blocks per CTA: G / 2 units x 128 TMEM columns
    static constexpr int TOTAL = TC_ATY_NSB * TC_S_BYTES + NB * D_BYTES + NS * 2 * TC_SLOT_BYTES + 128;
};

__global__ void __launch_bounds__(TC_THREADS, TC_ATY_OCC)
tc_aty_kernel(const uint2* __restrict__ entries, const int64_t* __restrict__ tile_ptr, int n_rb, int n_cb, int a_terms,
              float a_scale, int64_t n_eff, const uint8_t* __restrict__ Yprep, const float* __restrict__ scales,
              float* __restrict__ Z, int n_groups, int rb_per_range, int dbg, int pfd) {
    // (no integer round trip on this pointer: the compiler must keep seeing the shared address space, or every access
    // below becomes a generic LD/ST; the no-swizzle operand layouts only need 16 B alignment)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    uint8_t* sS = smem;
    uint8_t* sD = sS + TC_ATY_NSB * TC_S_BYTES;
    uint8_t* sRing = sD + AtySmem::NB * AtySmem::D_BYTES;
    __shared__ uint64_t s_full[TC_ATY_NSB], s_free[TC_ATY_NSB], d_full[AtySmem::NB], d_free[AtySmem::NB], acc_full;
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
        for (int i = 0; i < TC_ATY_NSB; i++) {
            mbar_init(&s_full[i], TC_GROUP_WARPS);
            mbar_init(&s_free[i], 1);
        }
        for (int i = 0; i < AtySmem::NB; i++) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_free[i], 1);
        }
        for (int i = 0; i < AtySmem::NS; i++) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_free[i], TC_GROUP_WARPS);
        }
        mbar_init(&acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s_tmem, TC_ATY_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const bool active = rb0 < rb1 && ntr > 0;
    TcSeq seq{n_cb, true, 0, rb0, rb1, cb_lo, ntr, n_units};

    if (warp < TC_SCATTER_WARPS) {
        if (active)
            tc_scatter_role<true, AtySmem::NS, TC_ATY_NSB>(entries, seq.n_units(), a_terms, a_scale, sS, sRing, sMeta, e_full, e_free, s_full,
                                               s_free, tid, dbg);
    } else if (warp >= TC_W_ELOAD) {
        if (active) tc_prefetch_prologue(entries, tile_ptr, seq, pfd, (warp - TC_W_ELOAD) * 32 + lane);
        if (active && lane == 0 && warp - TC_W_ELOAD < AtySmem::NS)
            tc_entry_loader<AtySmem::NS>(entries, tile_ptr, seq, sRing, sMeta, e_full, e_free, warp - TC_W_ELOAD, pfd, dbg);
    } else if (warp == TC_W_BLOAD) {
        if (lane == 0 && active) {
            uint32_t it = 0;
            for (int rb = rb0; rb < rb1; rb++, it++) {
                const int bb = it % AtySmem::NB;
                const uint32_t use = it / AtySmem::NB;
                if (use > 0) mbar_wait(&d_free[bb], (use - 1) & 1);
                if (dbg & 4) { mbar_arrive(&d_full[bb]); continue; }        // (timing experiment: no panel loads)
                mbar_expect_tx(&d_full[bb], AtySmem::D_BYTES);
                bulk_g2s(sD + bb * AtySmem::D_BYTES, Yprep + (size_t)rb * AtySmem::D_BYTES, AtySmem::D_BYTES, &d_full[bb]);
            }
        }
    } else if (warp == TC_W_MMA) {
        if (lane == 0 && active) {
            uint32_t it = 0, unit = 0;
            const uint64_t s_desc0 = umma_desc(smem_u32(sS), 2048, 128), d_desc0 = umma_desc(smem_u32(sD), 2048, 128);
            constexpr uint32_t idesc = tc_idesc(128);
            for (int rb = rb0; rb < rb1; rb++, it++) {
                const int bb = it % AtySmem::NB;
                mbar_wait(&d_full[bb], (it / AtySmem::NB) & 1);
                for (int u = 0; u < n_units; u++, unit++) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)u * 128u;
                    for (int term = 0; term < a_terms; term++) {
                        const uint32_t P = unit * (uint32_t)a_terms + (uint32_t)term;
                        const uint32_t sb = P % TC_ATY_NSB;
                        mbar_wait_crit(&s_full[sb], (P / TC_ATY_NSB) & 1);
                        tc_fence_after();
                        // K = 128 rows: eight K-steps, both operands advance 2 chunks x 2048 B per step
                        const uint64_t dd = d_desc0 + (uint64_t)bb * (AtySmem::D_BYTES >> 4);
                        const uint64_t sd = s_desc0 + (uint64_t)sb * (TC_S_BYTES >> 4);
                        if (!(dbg & 2)) {                                     // (timing experiment: no MMA)
                        umma_f16_run4(d_tmem, dd, sd, idesc, ((rb - rb0) | term) != 0, 256, 256);
                        umma_f16_run4(d_tmem, dd + 1024, sd + 1024, idesc, 1, 256, 256);
                        }
                        umma_commit(&s_free[sb]);
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
    if (warp == 0) tmem_dealloc(tmem_base, TC_ATY_TMEM_COLS);
}

// Z[r][j] = -mu[r] * cs[j]  (or 0)
__global__ void tc_init_z_kernel(float* __restrict__ Z, int64_t n_eff, const float* __restrict__ mu, const double* __restrict__ cs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_eff * LP) return;
    Z[i] = (mu && cs) ? (float)(-(double)mu[i >> 6] * cs[i & 63]) : 0.f;
}

bool tc_enabled(const salg_ctx* ctx) { return ctx->spmm_impl != 1; }

// L2 prefetch distance of the entry lists, in units (SALG_TC_PFD overrides; 0 = off)
static int tc_pfd() {
    const char* e = getenv("SALG_TC_PFD");
    const int v = e ? atoi(e) : 8;      // (measured: 8-16 units ahead -1..2 %, 32 and more +1 %; the ring is not the limit)
    return v < 0 ? 0 : v;
}

static TcTiles* tiles_of(salg_ctx* ctx, const salg_csr* c) {
    if (c->tc && (((TcTiles*)c->tc)->tm != nullptr) != tm_wanted(ctx, c)) {   // the context switched product generations
        SALG_CUDA(cudaStreamSynchronize(ctx->stream));
        tc_free(ctx, c->tc);
        c->tc = nullptr;
    }
    if (!c->tc) c->tc = tc_build<float>(ctx, c);
    return (TcTiles*)c->tc;
}

// d_scales = {s, 1 / (s a_scale)} with s the power of two that puts max |P| just below 2^14 (tm.cu uses it too)
void tc_panel_scales(salg_ctx* ctx, const float* P, int64_t n, float a_scale, float* d_scales, unsigned* d_amax) {
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
    SALG_CUDA(cudaGetLastError());
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
    fprintf(stderr, "[tc %s] scatter(grp 0) total %llu passes %llu: wait s_free %llu wait e_full %llu clear+loads %llu bar %llu scatter %llu fence %llu | loader0 wait e_free %llu over %llu units\n",
            what, h[0], h[9], h[1], h[2], h[3], h[4], h[5], h[6], h[20], h[21]);
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_tc_dbg, z, sizeof(z));
    fprintf(stderr, "[tc %s] mma total %llu: acc_free %llu d_full %llu s_full %llu issue %llu commit %llu | epilogue wait %llu work %llu\n", what,
            h[10], h[11], h[12], h[13], h[14], h[15], h[16], h[17]);
}

// Y (nrows x 64) = A X - 1 corr^T; d_amax (optional, device, zeroed here) receives the bits of max |Y|
void tc_spmm_A(salg_ctx* ctx, const salg_csr* c, const float* X, float* Y, const double* corr, unsigned* d_amax, int b_terms) {
    cudaStream_t st = ctx->stream;
    TcTiles* t = tiles_of(ctx, c);
    if (c->nrows == 0) return;
    double bytes = (double)c->nnz * 8 + (double)(c->nrows + 1) * 8 + (double)c->ncols * 60 * 4 + (double)c->nrows * 60 * 4;
    ProfScope ps(ctx, PROF_SPMM, bytes);   // includes the panel pre-split
    if (t->tm) {
        tm_spmm_A(ctx, c, t->tm, X, Y, corr, d_amax, b_terms);
        return;
    }
    DevBuf<uint8_t> Xprep((size_t)t->n_cb * AxSmem::D_BYTES, st);
    DevBuf<float> scales(2, st);
    DevBuf<unsigned> amax(1, st);
    tc_prepare_panel<TC_CB>(ctx, X, c->ncols, t->n_cb, t->a_scale, scales.get(), amax.get(), Xprep.get());
    if (d_amax) SALG_CUDA(cudaMemsetAsync(d_amax, 0, 4, st));
    set_max_dyn_smem(tc_ax_kernel, (int)(AxSmem::TOTAL));
    int n_pairs = t->n_rb / 2;
    int grid = n_pairs < ctx->sm_count * TC_OCC ? n_pairs : ctx->sm_count * TC_OCC;
    tc_ax_kernel<<<grid, TC_THREADS, AxSmem::TOTAL, st>>>(t->entries, t->tile_ptr, t->n_rb, t->n_cb, t->a_terms, t->a_scale,
                                                          c->nrows, Xprep.get(), scales.get(), Y, corr, d_amax,
                                                          getenv("SALG_TC_DBG") ? atoi(getenv("SALG_TC_DBG")) : 0, tc_pfd());
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    tc_dbg_print(ctx, "ax");
}

// ---- fused small-side step (TMEM-operand generation only): zside_solve (dense.cu) -> tc_zside_apply -> tc_spmm_A_prepped ----
bool tc_zside_supported(salg_ctx* ctx, const salg_csr* c) { return tiles_of(ctx, c)->tm != nullptr; }
size_t tc_xprep_bytes(salg_ctx* ctx, const salg_csr* c) { return tm_xprep_bytes(tiles_of(ctx, c)->tm); }
float tc_a_scale(salg_ctx* ctx, const salg_csr* c) { return tiles_of(ctx, c)->a_scale; }
void tc_zside_apply(salg_ctx* ctx, const salg_csr* c, float* Z, const float* d_M, const float* mu, const float* d_scales,
                    uint8_t* Xprep, double* corr) {
    TcTiles* t = tiles_of(ctx, c);
    ProfScope ps(ctx, PROF_PANELMUL, 2.0 * (double)c->ncols * 60 * 4);
    tm_zside_apply(ctx, c, t->tm, Z, d_M, mu, d_scales, Xprep, corr);
}
void tc_spmm_A_prepped(salg_ctx* ctx, const salg_csr* c, const uint8_t* Xprep, const float* d_scales, float* Y, const double* corr,
                       unsigned* d_amax, int b_terms) {
    TcTiles* t = tiles_of(ctx, c);
    if (c->nrows == 0) return;
    double bytes = (double)c->nnz * 8 + (double)(c->nrows + 1) * 8 + (double)c->ncols * 60 * 4 + (double)c->nrows * 60 * 4;
    ProfScope ps(ctx, PROF_SPMM, bytes);
    tm_spmm_A_prepped(ctx, c, t->tm, Xprep, d_scales, Y, corr, d_amax, b_terms);
}

// Fused pass over a tall panel Y (c->nrows x 64): Yprep = canonical two-term fp16 operand of Y (scale from d_amax, the
// bits of max |Y| written by tc_spmm_A), d_scales = {s, 1 / (s a_scale)}, G (GRAM_BUF f64) = [Y^T Y, 1^T Y] (local rows).
size_t tc_yprep_bytes(salg_ctx* ctx, const salg_csr* c) { return (size_t)tiles_of(ctx, c)->n_rb * AtySmem::D_BYTES; }
__global__ void tc_set_bits_kernel(unsigned* p, unsigned v) { *p = v; }
// d_amax[0] = bits of a known bound on max |Y| (panels with unit-norm columns: 1)
void tc_set_amax(salg_ctx* ctx, unsigned* d_amax, float bound) {
    unsigned bits;
    memcpy(&bits, &bound, 4);
    tc_set_bits_kernel<<<1, 1, 0, ctx->stream>>>(d_amax, bits);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
void tc_gram_prep(salg_ctx* ctx, const salg_csr* c, const float* Y, const unsigned* d_amax, uint8_t* Yprep, float* d_scales,
                  double* G, bool no_gram) {
    cudaStream_t st = ctx->stream;
    TcTiles* t = tiles_of(ctx, c);
    SALG_CUDA(cudaMemsetAsync(G, 0, GRAM_BUF * sizeof(double), st));
    if (c->nrows == 0) return;
    ProfScope ps(ctx, PROF_GRAM, (double)c->nrows * 60 * 4);
    tc_scale_kernel<<<1, 1, 0, st>>>(d_amax, t->a_scale, d_scales);
    ctx->n_launch++;
    const int n_rb_real = (int)ceil_div(c->nrows, TC_RB);
    set_max_dyn_smem(tc_gram_prep_kernel, (int)(GpSmem::TOTAL));
    int grid = n_rb_real < ctx->sm_count ? n_rb_real : ctx->sm_count;
    tc_gram_prep_kernel<<<grid, GP_THREADS, GpSmem::TOTAL, st>>>(Y, c->nrows, n_rb_real, d_scales, Yprep, G, no_gram ? 32 : 0);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

// Test / probe entry: G (GRAM_BUF) = [Y^T Y, 1^T Y] of a device panel through the fused pass; returns the average kernel
// time over `iters` launches.  yprep_out (optional, device, ceil(m/128) * 32 KB) receives the pre-split operand.
void tc_gram_probe(salg_ctx* ctx, const float* Y, int64_t m, double* G, uint8_t* yprep_out, int iters, double* avg_ms) {
    cudaStream_t st = ctx->stream;
    const int n_rb = (int)ceil_div(m, TC_RB);
    DevBuf<uint8_t> prep(yprep_out ? 0 : (size_t)std::max(n_rb, 1) * GpSmem::OP_BYTES, st);
    uint8_t* out = yprep_out ? yprep_out : prep.get();
    DevBuf<float> scales(2, st);
    DevBuf<unsigned> amax(1, st);
    SALG_CUDA(cudaMemsetAsync(amax.get(), 0, 4, st));
    if (m > 0) {
        int64_t ne = m * LP, want = ceil_div(ne, 256 * 8), cap = (int64_t)ctx->sm_count * 8;
        tc_absmax_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, st>>>(Y, ne, amax.get());
        ctx->n_launch++;
    }
    tc_scale_kernel<<<1, 1, 0, st>>>(amax.get(), 1.f, scales.get());
    ctx->n_launch++;
    set_max_dyn_smem(tc_gram_prep_kernel, (int)(GpSmem::TOTAL));
    const int grid = std::max(1, n_rb < ctx->sm_count ? n_rb : ctx->sm_count);
    const int dbg = getenv("SALG_GP_DBG") ? atoi(getenv("SALG_GP_DBG")) : 0;
    cudaEvent_t e0, e1;
    SALG_CUDA(cudaEventCreate(&e0));
    SALG_CUDA(cudaEventCreate(&e1));
    float ms = 0.f;
    for (int i = 0; i < iters + 1; i++) {          // first launch untimed
        SALG_CUDA(cudaMemsetAsync(G, 0, GRAM_BUF * sizeof(double), st));
        if (i == 1) SALG_CUDA(cudaEventRecord(e0, st));
        tc_gram_prep_kernel<<<grid, GP_THREADS, GpSmem::TOTAL, st>>>(Y, m, n_rb, scales.get(), out, G, dbg);
        ctx->n_launch++;
    }
    SALG_CUDA(cudaEventRecord(e1, st));
    SALG_CUDA(cudaGetLastError());
    SALG_CUDA(cudaStreamSynchronize(st));
    if (iters > 0) SALG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (avg_ms) *avg_ms = iters > 0 ? (double)ms / iters : 0.0;
}

static void tc_aty_launch(salg_ctx* ctx, const salg_csr* c, TcTiles* t, const uint8_t* Yprep, const float* scales, float* Z);

// Z (ncols x 64) = A^T Y - mu corr^T   (local rows only), Y given pre-split (tc_gram_prep)
void tc_spmm_At_prepped(salg_ctx* ctx, const salg_csr* c, const uint8_t* Yprep, const float* d_scales, float* Z,
                        const float* mu, const double* corr) {
    cudaStream_t st = ctx->stream;
    TcTiles* t = tiles_of(ctx, c);
    if (c->ncols == 0) return;
    double bytes = (double)c->nnz * 8 + (double)(c->nrows + 1) * 8 + (double)c->ncols * 60 * 4 + (double)c->nrows * 60 * 4;
    ProfScope ps(ctx, PROF_SPMMT, bytes);   // includes the Z initialisation
    tc_init_z_kernel<<<(unsigned)ceil_div(c->ncols * LP, 256), 256, 0, st>>>(Z, c->ncols, mu, corr);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    if (c->nrows == 0) return;
    tc_aty_launch(ctx, c, t, Yprep, d_scales, Z);
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
    tc_aty_launch(ctx, c, t, Yprep.get(), scales.get(), Z);
}

static void tc_aty_launch(salg_ctx* ctx, const salg_csr* c, TcTiles* t, const uint8_t* Yprep, const float* scales, float* Z) {
    cudaStream_t st = ctx->stream;
    if (t->tm) {
        tm_aty_launch(ctx, c, t->tm, Yprep, scales, Z);
        return;
    }
    int n_groups = (int)ceil_div(t->n_cb, AtySmem::G);
    int n_rb_real = (int)ceil_div(c->nrows, TC_RB);
    int ranges = ctx->sm_count * TC_ATY_OCC / n_groups;
    if (ranges < 1) ranges = 1;
    if (ranges > n_rb_real) ranges = n_rb_real;
    int rb_per_range = (int)ceil_div(n_rb_real, ranges);
    ranges = (int)ceil_div(n_rb_real, rb_per_range);
    set_max_dyn_smem(tc_aty_kernel, (int)(AtySmem::TOTAL));
    tc_aty_kernel<<<n_groups * ranges, TC_THREADS, AtySmem::TOTAL, st>>>(t->entries, t->tile_ptr, n_rb_real, t->n_cb, t->a_terms,
                                                                         t->a_scale, c->ncols, Yprep, scales, Z,
                                                                         n_groups, rb_per_range,
                                                                         getenv("SALG_TC_DBG") ? atoi(getenv("SALG_TC_DBG")) : 0, tc_pfd());
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    tc_dbg_print(ctx, "aty");
}

}  // namespace salg
