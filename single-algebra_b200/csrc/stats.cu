// stats.cu — MatrixSum::sum_col / sum_col_squared (src/sparse/csr.rs:259-312, 558-608) fused into ONE
// pass over the CSR, plus the per-column stored-entry count (MatrixNonZero::nonzero_col,
// csr.rs:23-64) and MatrixSum::sum_row (csr.rs:314-392).
//
// Column statistics are a scatter.  Each CTA privatises the accumulators of a column tile in shared
// memory (native shared-memory f32 atomics; f64 for f64 matrices), streams its share of the entries
// with 128-bit loads and flushes the tile once with f64 global atomics, so the HBM stream is the
// only large traffic: algorithmic bytes = nnz*(S+I) + 2*ncols*S  (SURVEY §8d).
#include "common.cuh"
#include <type_traits>

namespace salg {

template <typename A>
__device__ __forceinline__ void smem_add(A* p, A v) { atomicAdd(p, v); }

// {sum, sumsq} of one column live side by side.  Shared-memory floating-point adds are compare-and-swap loops
// (ATOMS.CAST.SPIN; only integer adds are native).
#ifndef STATS_MHOIST_
#define STATS_MHOIST_ 0     // batch the mask lookups ahead of the atomics: measured 3 % slower (64 registers)
#endif
// The shared-memory atomic unit retires ~2 lane-operations per clock and SM whatever the operation (native integer add
// or the compare-and-swap loop of a floating-point add: measured 1.9-2.0 per clock in all three kernels below), which
// caps a one-atomic-per-entry pass at ~72 % of the HBM roofline; these passes run at 90 % of THAT bound.  Measured
// alternatives, both worse (profiles/r01_v4_summary.md): STATS_L2_SLOTS_ of every 8 entries bypass shared memory and go to
// the f64 global accumulators as L2 reductions (REDG.ADD.F64) — the chip-wide L2 atomic rate is only ~1e11 per second,
// 1 slot: +-0, 2 slots: +35 %, 3 slots: +75 % time; STATS_PAIR64_: {sum, sumsq} updated by one 64-bit compare-and-swap
// loop instead of two 32-bit ones — ATOMS.CAST.SPIN.64 is ~8x slower per operation (4.1 -> 15.6 ms).
#ifndef STATS_L2_SLOTS_
#define STATS_L2_SLOTS_ 0
#endif
#ifndef STATS_PAIR64_
#define STATS_PAIR64_ 0
#endif
__device__ __forceinline__ void pair_add(float* pair, float x) {
#if STATS_PAIR64_
    unsigned long long* a = reinterpret_cast<unsigned long long*>(pair);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        const float s = __uint_as_float((unsigned)assumed) + x;
        const float q = fmaf(x, x, __uint_as_float((unsigned)(assumed >> 32)));
        old = atomicCAS(a, assumed, ((unsigned long long)__float_as_uint(q) << 32) | __float_as_uint(s));
    } while (old != assumed);
#else
    atomicAdd(pair, x);
    atomicAdd(pair + 1, x * x);
#endif
}
__device__ __forceinline__ void pair_add(double* pair, double x) {
    atomicAdd(pair, x);
    atomicAdd(pair + 1, x * x);
}

// ---- variant 1: the whole column range fits one shared-memory tile: flat stream over the entries ------
template <typename T, typename A, bool CNT>
__global__ void __launch_bounds__(1024, 1)
col_stats_flat_kernel(const uint32_t* __restrict__ col, const T* __restrict__ val, int64_t nnz, int ncols,
                      double* __restrict__ g_sum, double* __restrict__ g_sumsq, double* __restrict__ g_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A* acc = reinterpret_cast<A*>(smem_raw);                      // [2*ncols]: sum, sumsq interleaved
    unsigned* cnt = reinterpret_cast<unsigned*>(acc + 2 * (size_t)ncols);  // [ncols] when CNT
    for (int i = threadIdx.x; i < 2 * ncols; i += blockDim.x) acc[i] = A(0);
    if (CNT) for (int i = threadIdx.x; i < ncols; i += blockDim.x) cnt[i] = 0u;
    __syncthreads();

    // contiguous chunk per CTA; 2 x 4 entries per thread per step, all loads issued before the first atomic
    // (the arrays carry 16 entries of slack)
    int64_t n4 = (nnz + 3) >> 2;
    int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
    int64_t b0 = (int64_t)blockIdx.x * per, b1 = b0 + per;
    if (b1 > n4) b1 = n4;
    const uint4* col4 = reinterpret_cast<const uint4*>(col);
    constexpr int UF = 2;
    for (int64_t i0 = b0 + threadIdx.x; i0 < b1; i0 += (int64_t)UF * blockDim.x) {
        uint32_t cc[UF][4];
        T v[UF][4];
#pragma unroll
        for (int u = 0; u < UF; u++) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            if (i < b1) {
                const uint4 c = __ldcs(col4 + i);
                cc[u][0] = c.x; cc[u][1] = c.y; cc[u][2] = c.z; cc[u][3] = c.w;
                if (sizeof(T) == 4) {
                    float4 f = __ldcs(reinterpret_cast<const float4*>(val) + i);
                    v[u][0] = (T)f.x; v[u][1] = (T)f.y; v[u][2] = (T)f.z; v[u][3] = (T)f.w;
                } else {
                    double2 d0 = __ldcs(reinterpret_cast<const double2*>(val) + 2 * i);
                    double2 d1 = __ldcs(reinterpret_cast<const double2*>(val) + 2 * i + 1);
                    v[u][0] = (T)d0.x; v[u][1] = (T)d0.y; v[u][2] = (T)d1.x; v[u][3] = (T)d1.y;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UF; u++) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            if (i < b1) {
                const int64_t e = i << 2;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (e + k < nnz) {
                        if (4 * u + k >= 4 * UF - STATS_L2_SLOTS_) {                 // compile-time per unrolled slot
                            const double x = (double)v[u][k];
                            atomicAdd(&g_sum[cc[u][k]], x);
                            if (g_sumsq) atomicAdd(&g_sumsq[cc[u][k]], x * x);
                        } else {
                            pair_add(&acc[2 * cc[u][k]], (A)v[u][k]);
                        }
                        if (CNT) atomicAdd(&cnt[cc[u][k]], 1u);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) {
        A s = acc[2 * i], q = acc[2 * i + 1];
        if (s != A(0) || q != A(0)) {
            atomicAdd(&g_sum[i], (double)s);
            if (g_sumsq) atomicAdd(&g_sumsq[i], (double)q);
        }
        if (CNT) {
            unsigned n = cnt[i];
            if (n) atomicAdd(&g_cnt[i], (double)n);
        }
    }
}

// ---- variant 2: column tiles ----------------------------------------------------------------------------------
// The accumulators of all columns do not fit one CTA's shared memory: every CTA keeps its row block and sweeps it once
// per column tile.  The column indices of a row ascend, so the entries of tile t are a contiguous run of the row: pass t
// starts at the position pass t-1 stopped at (row_pos, one u32 per row, written and read by the same warp) and stops at
// the first batch that reaches the next tile.  Every entry is read from HBM once and no row is searched (the first
// version gave each (tile, row block) its own CTA and paid two dependent binary searches per row and tile).
template <typename T, typename A, bool CNT>
__global__ void __launch_bounds__(1024, 1)
col_stats_tiled_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col, const T* __restrict__ val,
                       int64_t nrows, int ncols, int tile_cols, int n_tiles, double* __restrict__ g_sum,
                       double* __restrict__ g_sumsq, double* __restrict__ g_cnt, const uint32_t* __restrict__ keepbits,
                       unsigned long long* __restrict__ row_kept, uint32_t* __restrict__ row_pos) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A* acc = reinterpret_cast<A*>(smem_raw);
    unsigned* cnt = reinterpret_cast<unsigned*>(acc + 2 * (size_t)tile_cols);
    unsigned* kb = cnt + (CNT ? tile_cols : 0);        // keep-bitmask of the whole column range (fused compaction count)
    if (keepbits) for (int i = threadIdx.x; i < (ncols + 31) / 32; i += blockDim.x) kb[i] = keepbits[i];
    const int64_t per = (nrows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = r0 + per < nrows ? r0 + per : nrows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    constexpr int U = 8;                                // 8 x 128 B per array in flight per warp
    for (int tile = 0; tile < n_tiles; tile++) {
        const uint32_t c0 = (uint32_t)tile * tile_cols;
        const bool last = tile == n_tiles - 1;
        const uint32_t c1 = last ? (uint32_t)ncols : c0 + tile_cols;
        for (int i = threadIdx.x; i < 2 * tile_cols; i += blockDim.x) acc[i] = A(0);
        if (CNT) for (int i = threadIdx.x; i < tile_cols; i += blockDim.x) cnt[i] = 0u;
        __syncthreads();
        for (int64_t r = r0 + warp; r < r1; r += nwarp) {
            const int64_t s = ptr[r];
            const uint32_t len = (uint32_t)(ptr[r + 1] - s);
            const uint32_t start = tile ? row_pos[r] : 0u;
            const uint32_t* __restrict__ colr = col + s;
            const T* __restrict__ valr = val + s;
            int kept = 0, n_in = 0;
            // one batch of 32*U entries from p0; FULL batches carry no per-load predicate (seven predicate registers
            // would cap the loads in flight at six), the tail batch clamps its index to the row's last entry.
            // Returns true when the batch reached the next tile (ascending columns: its last entry is its largest).
            auto batch = [&](uint32_t p0, auto full_tag) -> bool {
                constexpr bool FULL = decltype(full_tag)::value;
                uint32_t cc[U];
                T vv[U];
                const uint32_t last_e = len - 1;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    uint32_t q = p0 + lane + 32 * u;
                    if (!FULL) q = q < last_e ? q : last_e;
                    cc[u] = __ldcs(colr + q);
                    vv[u] = __ldcs(valr + q);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const bool ok = FULL || p0 + lane + 32 * u < len;
                    if (ok && (last || cc[u] < c1)) {
                        const uint32_t c = cc[u] - c0;
                        const A x = (A)vv[u];
                        n_in++;
                        const bool to_l2 = u >= U - STATS_L2_SLOTS_;                 // compile-time per unrolled slot
                        bool sq = true;
                        if (keepbits) {
                            // masked fit: sum of squares is only needed for the kept columns (total_var,
                            // pca/sparse_masked/mod.rs:303-307)
                            const unsigned kbit = (kb[cc[u] >> 5] >> (cc[u] & 31)) & 1u;
                            kept += kbit;
                            sq = kbit != 0u;
                        }
                        if (to_l2) {
                            atomicAdd(&g_sum[cc[u]], (double)x);
                            if (sq && g_sumsq) atomicAdd(&g_sumsq[cc[u]], (double)x * (double)x);
                        } else if (sq) {
                            pair_add(&acc[2 * c], x);
                        } else {
                            smem_add(&acc[2 * c], x);
                        }
                        if (CNT) atomicAdd(&cnt[c], 1u);
                    }
                }
                return !last && __any_sync(0xFFFFFFFFu, cc[U - 1] >= c1);
            };
            uint32_t p0 = start;
            bool done = false;
            for (; !done && p0 + 32 * U <= len; p0 += 32 * U) done = batch(p0, std::true_type{});
            if (!done && p0 < len) batch(p0, std::false_type{});
            if (!last) {
#pragma unroll
                for (int o = 16; o; o >>= 1) n_in += __shfl_xor_sync(0xFFFFFFFFu, n_in, o);
                if (lane == 0) row_pos[r] = start + (uint32_t)n_in;
            }
            if (keepbits) {
#pragma unroll
                for (int o = 16; o; o >>= 1) kept += __shfl_xor_sync(0xFFFFFFFFu, kept, o);
                if (lane == 0) row_kept[r] = (tile ? row_kept[r] : 0ull) + (unsigned long long)kept;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < (int)(c1 - c0); i += blockDim.x) {
            A s = acc[2 * i], q = acc[2 * i + 1];
            if (s != A(0) || q != A(0)) {
                atomicAdd(&g_sum[c0 + i], (double)s);
                if (g_sumsq) atomicAdd(&g_sumsq[c0 + i], (double)q);
            }
            if (CNT) {
                unsigned n = cnt[i];
                if (n) atomicAdd(&g_cnt[c0 + i], (double)n);
            }
        }
        __syncthreads();
    }
}

// ---- variant 3 (masked fits): one shared-memory tile for the whole column range --------------------------------
// A masked fit needs the sum of every column (mean_ has full length, pca/sparse_masked/mod.rs:280-291) but the sum of
// squares only of the kept ones (total_var, :303-307): sum[ncols] + sumsq[n_kept] fit one tile where sum + sumsq of
// every column would need two (and with them a binary search per row and tile).  Warp per row, 8 independent
// 128 B loads per array in flight per warp; the kept-entry count of the row (the compaction's count pass) rides along.
// kb = keep bitmask words followed by the exclusive prefix popcounts of the words (rank of a kept column).
// INTSUM: the column sums are accumulated as u32 with the NATIVE shared-memory integer atomic (f32 adds are a
// compare-and-swap loop, and their rate — not HBM — bounded this pass).  Valid while every value is an integer in
// [0, 65536) and a CTA sees at most 65536 rows (raw count matrices); any other value raises flags[1] and the caller
// repeats the pass with f32 accumulators.  Integer sums are exact.
__global__ void values_integral_probe_kernel(const float* __restrict__ val, int64_t n, int* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = val[i];
    if (!(x >= 0.f && x < 65536.f && x == truncf(x))) atomicOr(flag, 1);
}

template <typename T, bool INTSUM>
__global__ void __launch_bounds__(1024, 1)
col_stats_masked_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col, const T* __restrict__ val,
                        int64_t nrows, int ncols, int n_kept, double* __restrict__ g_sum, double* __restrict__ g_sumsq,
                        const uint32_t* __restrict__ keepbits, unsigned long long* __restrict__ row_kept,
                        uint32_t* __restrict__ kept_col, T* __restrict__ kept_val, int kept_shift,
                        int* __restrict__ flags /* [0] kept-slot overflow, [1] INTSUM violated */) {
    // kept_col / kept_val (optional, (nnz >> kept_shift) + 16 long): the kept entries of row r, renumbered to compact
    // column ids and in their original order, are written from ptr[r] >> kept_shift on — the compaction's WRITE pass
    // fused in as well; the tile builder (tc.cu) consumes them from there, so the 16.8 GB operator is read once, not
    // twice.  A row whose kept entries do not fit before the next row's slot raises *kept_overflow (the caller then
    // falls back to the separate compaction pass).
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nw32 = (ncols + 31) / 32;
    float* sum = reinterpret_cast<float*>(smem_raw);          // [ncols (+1 pad)]
    float* sq = sum + ((ncols + 1) & ~1);                      // [n_kept] f32, or [n_kept] u64 when INTSUM (8-byte aligned)
    // {mask word, exclusive prefix popcount} pairs: one 8-byte shared-memory lookup per entry (the random-bank lookups
    // and the accumulator atomics, not HBM, bound this pass)
    const int acc_words = ((ncols + 1) & ~1) + 2 * n_kept;     // accumulator area in 4-byte words (sq sized for u64)
    uint2* kb2 = reinterpret_cast<uint2*>(smem_raw + (size_t)acc_words * 4);   // [nw32]
    for (int i = threadIdx.x; i < acc_words; i += blockDim.x) sum[i] = 0.f;
    for (int i = threadIdx.x; i < nw32; i += blockDim.x) kb2[i] = make_uint2(keepbits[i], keepbits[nw32 + i]);
    __syncthreads();
    const int64_t per = (nrows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = r0 + per < nrows ? r0 + per : nrows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    constexpr int U = 8;
    bool not_int = false;
    unsigned* sum_u = reinterpret_cast<unsigned*>(sum);
    // INTSUM: exact integer sums of squares as two u32 accumulators per kept column (low / high 16 bits of x^2 < 2^32,
    // each below 2^32 over <= 65536 rows): native atomics, where a 64-bit shared add would be a compare-and-swap loop
    unsigned* sq_lo = reinterpret_cast<unsigned*>(sq);
    unsigned* sq_hi = sq_lo + n_kept;
    unsigned or_bits = 0, or_diff = 0;                          // INTSUM validity, checked once at the end
    // The pass is instruction-issue bound (ncu: 73 % issue utilisation, 82 instructions per entry in the first version), so
    // the body is kept branch-free: lanes past the end of the row add 0 to column 0 instead of diverging, offsets inside
    // a row are 32-bit, and raw counts use native integer atomics for both accumulators (no compare-and-swap loops).
    for (int64_t r = r0 + warp; r < r1; r += nwarp) {
        const int64_t s = ptr[r];
        const uint32_t len = (uint32_t)(ptr[r + 1] - s);
        const uint32_t* __restrict__ colr = col + s;
        const T* __restrict__ valr = val + s;
        int kept = 0;                                          // warp-uniform running count of kept entries
        const int64_t ks = s >> kept_shift;
        const int cap = (int)(((s + len) >> kept_shift) - ks); // slot of this row in the scaled scratch
        uint32_t* __restrict__ kc = kept_col ? kept_col + ks : nullptr;
        T* __restrict__ kv = kept_col ? kept_val + ks : nullptr;
        // one batch of 32*U entries (warp-uniform trip count: ballots inside).  FULL batches carry no per-load predicate
        // (seven predicate registers would cap the loads in flight at six); the tail batch clamps its index to the row's
        // last entry and lanes past the end add 0 to that entry's column instead of diverging.
        auto batch = [&](uint32_t p0, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            uint32_t cc[U];
            T vv[U];
            const uint32_t last_e = len - 1;
#pragma unroll
            for (int u = 0; u < U; u++) {
                uint32_t q = p0 + lane + 32 * u;
                if (!FULL) q = q < last_e ? q : last_e;
                cc[u] = __ldcs(colr + q);
                vv[u] = __ldcs(valr + q);
            }
            // mask lookups of the whole batch first (independent shared-memory reads in flight instead of one exposed
            // read latency per entry): {kept bit, rank among the kept columns} packed into one register
#if STATS_MHOIST_
            unsigned kr[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const bool ok = FULL || p0 + lane + 32 * u < len;
                const unsigned b = cc[u] & 31u;
                const uint2 wp = kb2[cc[u] >> 5];
                const unsigned kb1 = ok ? (wp.x >> b) & 1u : 0u;
                kr[u] = (kb1 << 31) | (wp.y + __popc(wp.x & ((1u << b) - 1u)));
            }
#endif
#pragma unroll
            for (int u = 0; u < U; u++) {
                const bool ok = FULL || p0 + lane + 32 * u < len;
                const float x = ok ? (float)vv[u] : 0.f;
#if STATS_MHOIST_
                const bool kbit = (kr[u] >> 31) != 0u;
                const unsigned rank = kr[u] & 0x7FFFFFFFu;
#else
                const unsigned b = cc[u] & 31u;
                const uint2 wp = kb2[cc[u] >> 5];
                const bool kbit = ok && ((wp.x >> b) & 1u);
                const unsigned rank = wp.y + __popc(wp.x & ((1u << b) - 1u));
#endif
                unsigned xi = 0;
                const bool to_l2 = u >= U - STATS_L2_SLOTS_;                         // compile-time per unrolled slot
                if (INTSUM) {
                    xi = __float2uint_rz(x);
                    or_bits |= xi;                                                   // any value >= 65536 sets a high bit
                    or_diff |= __float_as_uint(__uint2float_rn(xi)) ^ __float_as_uint(x);   // any non-integer / negative
                    if (!to_l2) atomicAdd(&sum_u[cc[u]], xi);
                } else {
                    if (!to_l2) atomicAdd(&sum[cc[u]], x);
                }
                if (to_l2 && ok) atomicAdd(&g_sum[cc[u]], (double)x);
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, kbit);
                if (kbit) {
                    if (INTSUM) {
                        const unsigned x2 = xi * xi;
                        atomicAdd(&sq_lo[rank], x2 & 0xFFFFu);
                        atomicAdd(&sq_hi[rank], x2 >> 16);
                    } else {
                        atomicAdd(&sq[rank], x * x);
                    }
                    if (kc) {
                        const int k = kept + __popc(bal & ((1u << lane) - 1u));
                        if (k < cap) {
                            kc[k] = rank;
                            kv[k] = vv[u];
                        }
                    }
                }
                kept += __popc(bal);
            }
        };
        uint32_t p0 = 0;
        for (; p0 + 32 * U <= len; p0 += 32 * U) batch(p0, std::true_type{});
        if (p0 < len) batch(p0, std::false_type{});
        if (lane == 0) {
            row_kept[r] = (unsigned long long)kept;
            if (kept_col && kept > cap) atomicOr(&flags[0], 1);
        }
    }
    if (INTSUM && (not_int || (or_bits >> 16) != 0u || or_diff != 0u)) atomicOr(&flags[1], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) {
        const double sv = INTSUM ? (double)sum_u[i] : (double)sum[i];
        if (sv != 0.0) atomicAdd(&g_sum[i], sv);
        const unsigned b = (unsigned)i & 31u;
        const uint2 wp = kb2[(unsigned)i >> 5];
        if (g_sumsq && ((wp.x >> b) & 1u)) {
            const unsigned k = wp.y + __popc(wp.x & ((1u << b) - 1u));
            const double q = INTSUM ? (double)reinterpret_cast<const unsigned*>(sq)[k] +
                                          65536.0 * (double)reinterpret_cast<const unsigned*>(sq)[n_kept + k]
                                    : (double)sq[k];
            if (q != 0.0) atomicAdd(&g_sumsq[i], q);
        }
    }
}

template <typename T, typename A, bool CNT>
static void col_stats_launch(salg_ctx* ctx, const salg_csr* c, double* d_sum, double* d_sumsq, double* d_cnt,
                             const uint32_t* keepbits, int64_t* row_kept, int64_t n_kept, uint32_t* kept_col, T* kept_val,
                             int kept_shift, int* kept_overflow) {
    cudaStream_t st = ctx->stream;
    const size_t kMaxSmem = 200 * 1024;
    size_t per_col = 2 * sizeof(A) + (CNT ? 4 : 0);
    int ncols = (int)c->ncols;
    size_t kb_bytes = keepbits ? (size_t)((ncols + 31) / 32) * 4 : 0;
    size_t need = per_col * (size_t)ncols;
    const size_t need_masked = ((size_t)ncols + 2 + 2 * (size_t)n_kept) * 4 + 2 * kb_bytes + 8;
    if (keepbits && row_kept && !CNT && sizeof(T) == 4 && need_masked <= kMaxSmem && !getenv("SALG_STATS_TILED")) {
        int grid = (int)(c->nrows < ctx->sm_count ? (c->nrows > 0 ? c->nrows : 1) : ctx->sm_count);
        DevBuf<int> own_flags(kept_overflow ? 0 : 2, st);
        int* flags = kept_overflow ? kept_overflow : own_flags.get();
        if (!kept_overflow) SALG_CUDA(cudaMemsetAsync(flags, 0, 8, st));
        // integer accumulators when the values look like raw counts (probe on a prefix; the kernel re-checks every value)
        bool try_int = ceil_div(c->nrows, grid) <= 65536 && !getenv("SALG_STATS_NO_INT");
        if (try_int) {
            const int64_t np = c->nnz < 65536 ? c->nnz : 65536;
            values_integral_probe_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, st>>>((const float*)c->val, np, flags + 1);
            ctx->n_launch++;
            int h = 0;
            SALG_CUDA(cudaMemcpyAsync(&h, flags + 1, 4, cudaMemcpyDeviceToHost, st));
            SALG_CUDA(cudaStreamSynchronize(st));
            try_int = h == 0;
            if (h) SALG_CUDA(cudaMemsetAsync(flags + 1, 0, 4, st));
        }
        for (int attempt = 0; attempt < 2; attempt++) {
            if (try_int) {
                auto k = col_stats_masked_kernel<T, true>;
                set_max_dyn_smem(k, (int)((int)kMaxSmem));
                k<<<grid, 1024, need_masked, st>>>(c->row_ptr, c->col, (const T*)c->val, c->nrows, ncols, (int)n_kept, d_sum,
                                                   d_sumsq, keepbits, (unsigned long long*)row_kept, kept_col, kept_val,
                                                   kept_shift, flags);
            } else {
                auto k = col_stats_masked_kernel<T, false>;
                set_max_dyn_smem(k, (int)((int)kMaxSmem));
                k<<<grid, 1024, need_masked, st>>>(c->row_ptr, c->col, (const T*)c->val, c->nrows, ncols, (int)n_kept, d_sum,
                                                   d_sumsq, keepbits, (unsigned long long*)row_kept, kept_col, kept_val,
                                                   kept_shift, flags);
            }
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
            if (!try_int) break;
            int h = 0;
            SALG_CUDA(cudaMemcpyAsync(&h, flags + 1, 4, cudaMemcpyDeviceToHost, st));
            SALG_CUDA(cudaStreamSynchronize(st));
            if (!h) break;
            // a non-integral value past the probed prefix: redo with f32 accumulators
            try_int = false;
            SALG_CUDA(cudaMemsetAsync(d_sum, 0, (size_t)ncols * 8, st));
            if (d_sumsq) SALG_CUDA(cudaMemsetAsync(d_sumsq, 0, (size_t)ncols * 8, st));
            SALG_CUDA(cudaMemsetAsync(flags, 0, 8, st));
        }
    } else if (kept_col) {
        throw Error(SALG_ERR_UNSUPPORTED, "fused compaction needs the single-tile masked statistics kernel");
    } else if (need <= kMaxSmem && !keepbits) {
        auto k = col_stats_flat_kernel<T, A, CNT>;
        set_max_dyn_smem(k, (int)((int)kMaxSmem));
        int64_t n4 = (c->nnz + 3) / 4;
        int grid = (int)(n4 < (int64_t)ctx->sm_count * 1024 ? ceil_div(n4 > 0 ? n4 : 1, 1024) : ctx->sm_count);
        k<<<grid, 1024, need, st>>>(c->col, (const T*)c->val, c->nnz, ncols, d_sum, d_sumsq, d_cnt);
        ctx->n_launch++;
    } else {
        int tile_cols = (int)((kMaxSmem - kb_bytes) / per_col);
        int n_tiles = (int)ceil_div(ncols, tile_cols);
        tile_cols = (int)ceil_div(ncols, n_tiles);  // balance the tiles
        int grid = (int)(c->nrows < ctx->sm_count ? (c->nrows > 0 ? c->nrows : 1) : ctx->sm_count);
        auto k = col_stats_tiled_kernel<T, A, CNT>;
        set_max_dyn_smem(k, (int)((int)kMaxSmem));
        DevBuf<uint32_t> row_pos(n_tiles > 1 ? (size_t)c->nrows : 0, st);   // where the next tile's pass resumes in each row
        k<<<grid, 1024, per_col * (size_t)tile_cols + kb_bytes, st>>>(
            c->row_ptr, c->col, (const T*)c->val, c->nrows, ncols, tile_cols, n_tiles, d_sum, d_sumsq, d_cnt, keepbits,
            (unsigned long long*)row_kept, row_pos.get());
        ctx->n_launch++;
    }
    SALG_CUDA(cudaGetLastError());
}

template <typename T>
void col_stats_device(salg_ctx* ctx, const salg_csr* c, double* d_sum, double* d_sumsq, double* d_cnt,
                      const uint32_t* keepbits, int64_t* row_kept, int64_t n_kept, uint32_t* kept_col, void* kept_val,
                      int kept_shift, int* kept_overflow) {
    cudaStream_t st = ctx->stream;
    int64_t ncols = c->ncols;
    if (ncols == 0) return;
    SALG_CUDA(cudaMemsetAsync(d_sum, 0, (size_t)ncols * 8, st));
    if (d_sumsq) SALG_CUDA(cudaMemsetAsync(d_sumsq, 0, (size_t)ncols * 8, st));
    if (d_cnt) SALG_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)ncols * 8, st));
    if (row_kept) SALG_CUDA(cudaMemsetAsync(row_kept, 0, (size_t)(c->nrows + 1) * 8, st));
    if (c->nnz > 0) {
        ProfScope ps(ctx, PROF_STATS, (double)c->nnz * (sizeof(T) + 4) + 2.0 * (double)ncols * sizeof(T));
        // f32 matrices accumulate per-CTA partials in f32 (exact for integer counts below 2^24 per CTA
        // chunk), f64 matrices in f64; the cross-CTA reduction is always f64.
        if (sizeof(T) == 4) {
            if (d_cnt) col_stats_launch<T, float, true>(ctx, c, d_sum, d_sumsq, d_cnt, keepbits, row_kept, n_kept, kept_col, (T*)kept_val,
                                                            kept_shift, kept_overflow);
            else col_stats_launch<T, float, false>(ctx, c, d_sum, d_sumsq, d_cnt, keepbits, row_kept, n_kept, kept_col, (T*)kept_val,
                                                            kept_shift, kept_overflow);
        } else {
            if (d_cnt) col_stats_launch<T, double, true>(ctx, c, d_sum, d_sumsq, d_cnt, keepbits, row_kept, n_kept, kept_col, (T*)kept_val,
                                                            kept_shift, kept_overflow);
            else col_stats_launch<T, double, false>(ctx, c, d_sum, d_sumsq, d_cnt, keepbits, row_kept, n_kept, kept_col, (T*)kept_val,
                                                            kept_shift, kept_overflow);
        }
    }
    // row-sharded context: global column sums
    if (ctx->nranks > 1) {      // one NCCL launch for the two or three vectors
        ProfScope ps(ctx, PROF_ALLREDUCE, (double)ncols * 8 * (1 + (d_sumsq ? 1 : 0) + (d_cnt ? 1 : 0)));
        SALG_NCCL(ncclGroupStart());
        SALG_NCCL(ncclAllReduce(d_sum, d_sum, (size_t)ncols, ncclDouble, ncclSum, ctx->comm, ctx->stream));
        if (d_sumsq) SALG_NCCL(ncclAllReduce(d_sumsq, d_sumsq, (size_t)ncols, ncclDouble, ncclSum, ctx->comm, ctx->stream));
        if (d_cnt) SALG_NCCL(ncclAllReduce(d_cnt, d_cnt, (size_t)ncols, ncclDouble, ncclSum, ctx->comm, ctx->stream));
        SALG_NCCL(ncclGroupEnd());
    }
}
template void col_stats_device<float>(salg_ctx*, const salg_csr*, double*, double*, double*, const uint32_t*, int64_t*, int64_t,
                                      uint32_t*, void*, int, int*);
template void col_stats_device<double>(salg_ctx*, const salg_csr*, double*, double*, double*, const uint32_t*, int64_t*, int64_t,
                                       uint32_t*, void*, int, int*);
// can col_stats_device write the kept entries (kept_col / kept_val) for this matrix / mask?
// ---- streamed fits: the masked statistics + fused compaction pass over ONE ROW CHUNK of a host-resident matrix ------
// ptr = device row offsets of the chunk's first row (global offsets), col / val = the chunk's staged entries rebased so
// that col[ptr[r]] is row r's first entry (the pointers may lie below the staging buffer; only offsets inside the chunk
// are dereferenced).  Accumulates into d_sum / d_sumsq / row_kept / the kept-entry scratch without zeroing anything, on the
// given stream.  probe = true: returns whether the chunk's leading values look like raw counts (integer accumulators).
bool col_stats_probe_int_f32(salg_ctx* ctx, cudaStream_t st, const float* d_val, int64_t n, int* d_flag1) {
    const int64_t np = n < 65536 ? n : 65536;
    if (np <= 0) return true;
    values_integral_probe_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, st>>>(d_val, np, d_flag1);
    ctx->n_launch++;
    int h = 0;
    SALG_CUDA(cudaMemcpyAsync(&h, d_flag1, 4, cudaMemcpyDeviceToHost, st));
    SALG_CUDA(cudaStreamSynchronize(st));
    if (h) SALG_CUDA(cudaMemsetAsync(d_flag1, 0, 4, st));
    return h == 0;
}

void col_stats_masked_chunk_f32(salg_ctx* ctx, cudaStream_t st, const int64_t* ptr, const uint32_t* col, const float* val,
                                int64_t nrows, int64_t ncols, int64_t n_kept, double* d_sum, double* d_sumsq,
                                const uint32_t* keepbits, int64_t* row_kept, uint32_t* kept_col, float* kept_val, int kept_shift,
                                int* flags, bool intsum) {
    if (nrows <= 0) return;
    const size_t kb_bytes = (size_t)((ncols + 31) / 32) * 4;
    const size_t need_masked = ((size_t)ncols + 2 + 2 * (size_t)n_kept) * 4 + 2 * kb_bytes + 8;
    const int grid = (int)(nrows < ctx->sm_count ? nrows : ctx->sm_count);
    SALG_REQUIRE(!intsum || ceil_div(nrows, grid) <= 65536, SALG_ERR_UNSUPPORTED, "row chunk too long for integer accumulators");
    if (intsum) {
        auto k = col_stats_masked_kernel<float, true>;
        set_max_dyn_smem(k, 200 * 1024);
        k<<<grid, 1024, need_masked, st>>>(ptr, col, val, nrows, (int)ncols, (int)n_kept, d_sum, d_sumsq, keepbits,
                                           (unsigned long long*)row_kept, kept_col, kept_val, kept_shift, flags);
    } else {
        auto k = col_stats_masked_kernel<float, false>;
        set_max_dyn_smem(k, 200 * 1024);
        k<<<grid, 1024, need_masked, st>>>(ptr, col, val, nrows, (int)ncols, (int)n_kept, d_sum, d_sumsq, keepbits,
                                           (unsigned long long*)row_kept, kept_col, kept_val, kept_shift, flags);
    }
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

bool col_stats_can_fuse_compaction(const salg_csr* c, int64_t n_kept) {
    const size_t kb = (size_t)((c->ncols + 31) / 32) * 4;
    return c->dtype == SALG_F32 && ((size_t)c->ncols + 2 + 2 * (size_t)n_kept) * 4 + 2 * kb + 8 <= 200 * 1024 && !getenv("SALG_STATS_TILED");
}

// ---- sum_row ----------------------------------------------------------------------------------------------
// per-row sum and sum of squares (the CSC twins: a column of A is a row of the stored CSR of A^T)
// One warp per row, kSumU independent 128 B loads in flight per lane (one load per trip left the pass latency bound).
constexpr int kSumU = 16;

// Whole batches carry no per-load predicate (ptxas has seven predicate registers: predicated loads are issued six at
// a time, then wait); the tail batch clamps its index to the last entry of the row and zeroes the value afterwards.
template <typename T, bool SQ>
__device__ __forceinline__ void warp_row_sums(const T* __restrict__ vr, uint32_t len, int lane, double& a, double& q) {
    constexpr int RU = sizeof(T) == 4 ? kSumU : kSumU / 2;
    a = 0.0;
    q = 0.0;
    uint32_t p0 = 0;
    for (; p0 + 32 * RU <= len; p0 += 32 * RU) {
        T v[RU];
#pragma unroll
        for (int u = 0; u < RU; u++) v[u] = __ldcs(vr + p0 + lane + 32 * u);
#pragma unroll
        for (int u = 0; u < RU; u++) {
            const double x = (double)v[u];
            a += x;
            if (SQ) q = fma(x, x, q);
        }
    }
    if (p0 < len) {
        T v[RU];
        const uint32_t last = len - 1;
#pragma unroll
        for (int u = 0; u < RU; u++) {
            const uint32_t i = p0 + lane + 32 * u;
            v[u] = __ldcs(vr + (i < last ? i : last));
        }
#pragma unroll
        for (int u = 0; u < RU; u++) {
            const double x = p0 + lane + 32 * u < len ? (double)v[u] : 0.0;
            a += x;
            if (SQ) q = fma(x, x, q);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        if (SQ) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
    }
}

template <typename T>
__global__ void __launch_bounds__(256, 4) sum_row_sq_kernel(const int64_t* __restrict__ ptr, const T* __restrict__ val, int64_t nrows,
                                  T* __restrict__ out, T* __restrict__ out_sq) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < nrows; r += nw) {
        const int64_t s = ptr[r];
        double a, q;
        warp_row_sums<T, true>(val + s, (uint32_t)(ptr[r + 1] - s), lane, a, q);
        if (lane == 0) {
            if (out) out[r] = (T)a;
            if (out_sq) out_sq[r] = (T)q;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256, 4) sum_row_kernel(const int64_t* __restrict__ ptr, const T* __restrict__ val, int64_t nrows,
                               T* __restrict__ out) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < nrows; r += nw) {
        const int64_t s = ptr[r];
        double a, q;
        warp_row_sums<T, false>(val + s, (uint32_t)(ptr[r + 1] - s), lane, a, q);
        if (lane == 0) out[r] = (T)a;
    }
}

template <typename T>
void sum_row_device(salg_ctx* ctx, const salg_csr* c, T* d_out) {
    if (c->nrows == 0) return;
    ProfScope ps(ctx, PROF_STATS, (double)c->nnz * sizeof(T) + (double)(c->nrows + 1) * 8 + (double)c->nrows * sizeof(T));
    int64_t want = ceil_div(c->nrows * 32, 256);
    int64_t cap = (int64_t)ctx->sm_count * 16;
    sum_row_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(c->row_ptr, (const T*)c->val,
                                                                                    c->nrows, d_out);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void sum_row_device<float>(salg_ctx*, const salg_csr*, float*);
template void sum_row_device<double>(salg_ctx*, const salg_csr*, double*);

template <typename T>
__global__ void cast_from_f64_kernel(const double* __restrict__ src, T* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (T)src[i];
}

template <typename T>
static void sum_col_api(salg_ctx* ctx, const salg_csr* c, T* sum, T* sumsq) {
    SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
    SALG_REQUIRE(sum, SALG_ERR_BAD_ARG, "sum is NULL");
    SALG_REQUIRE(c->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "csr value type does not match the entry point");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int64_t n = c->ncols;
    if (n == 0) return;
    DevBuf<double> d_sum((size_t)n, st), d_sq((size_t)(sumsq ? n : 0), st);
    col_stats_device<T>(ctx, c, d_sum.get(), sumsq ? d_sq.get() : nullptr, nullptr);
    DevBuf<T> o((size_t)n, st);
    cast_from_f64_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_sum.get(), o.get(), n);
    ctx->n_launch++;
    SALG_CUDA(cudaMemcpyAsync(sum, o.get(), (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
    if (sumsq) {
        SALG_CUDA(cudaStreamSynchronize(st));
        cast_from_f64_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_sq.get(), o.get(), n);
        ctx->n_launch++;
        SALG_CUDA(cudaMemcpyAsync(sumsq, o.get(), (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
    SALG_CUDA(cudaGetLastError());
    SALG_CUDA(cudaStreamSynchronize(st));
}

template <typename T>
static void sum_row_api(salg_ctx* ctx, const salg_csr* c, T* out) {
    SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
    SALG_REQUIRE(out || c->nrows == 0, SALG_ERR_BAD_ARG, "out is NULL");
    SALG_REQUIRE(c->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "csr value type does not match the entry point");
    SALG_CUDA(cudaSetDevice(ctx->device));
    if (c->nrows == 0) return;
    DevBuf<T> o((size_t)c->nrows, ctx->stream);
    sum_row_device<T>(ctx, c, o.get());
    SALG_CUDA(cudaMemcpyAsync(out, o.get(), (size_t)c->nrows * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    SALG_CUDA(cudaStreamSynchronize(ctx->stream));
}

// CscMatrix::sum_col / sum_col_squared (src/sparse/csc.rs:157-197, 323-335): one pass over the stored rows of A^T
template <typename T>
static void csc_sum_col_api(salg_ctx* ctx, const salg_csr* ct, T* sum, T* sumsq) {
    SALG_REQUIRE(ctx && ct, SALG_ERR_BAD_ARG, "ctx/csc is NULL");
    SALG_REQUIRE(sum || sumsq, SALG_ERR_BAD_ARG, "sum and sumsq are both NULL");
    SALG_REQUIRE(ct->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "csc value type does not match the entry point");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n = ct->nrows;      // columns of A
    if (n == 0) return;
    DevBuf<T> o((size_t)n, st), q((size_t)(sumsq ? n : 0), st);
    {
        ProfScope ps(ctx, PROF_STATS, (double)ct->nnz * sizeof(T) + (double)(n + 1) * 8 + (double)n * sizeof(T) * (sumsq ? 2 : 1));
        int64_t want = ceil_div(n * 32, 256), cap = (int64_t)ctx->sm_count * 16;
        sum_row_sq_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(ct->row_ptr, (const T*)ct->val, n, o.get(),
                                                                                 sumsq ? q.get() : nullptr);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    if (ctx->nranks > 1) {   // a CSC handle sharded over ranks holds a block of COLUMNS: nothing to reduce
    }
    if (sum) SALG_CUDA(cudaMemcpyAsync(sum, o.get(), (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
    if (sumsq) SALG_CUDA(cudaMemcpyAsync(sumsq, q.get(), (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
    SALG_CUDA(cudaStreamSynchronize(st));
}

__global__ void var_col_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double n_rows,
                               double* __restrict__ var, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // MatrixVariance::var_col (src/sparse/csr.rs:649-657): (sumsq/n - mean^2) * n/(n-1)
    double mean = sum[i] / n_rows;
    double v = sumsq[i] / n_rows - mean * mean;
    var[i] = n_rows > 1.0 ? v * (n_rows / (n_rows - 1.0)) : v;
}

int64_t global_nrows(salg_ctx* ctx, int64_t local_rows) {
    if (ctx->nranks <= 1) return local_rows;
    DevBuf<double> d(1, ctx->stream);
    double h = (double)local_rows;
    SALG_CUDA(cudaMemcpyAsync(d.get(), &h, 8, cudaMemcpyHostToDevice, ctx->stream));
    allreduce_f64(ctx, d.get(), 1);
    SALG_CUDA(cudaMemcpyAsync(&h, d.get(), 8, cudaMemcpyDeviceToHost, ctx->stream));
    SALG_CUDA(cudaStreamSynchronize(ctx->stream));
    return (int64_t)(h + 0.5);
}

}  // namespace salg

using namespace salg;

extern "C" {

int salg_sum_col_f32(salg_ctx* ctx, const salg_csr* c, float* sum, float* sumsq) {
    return guarded([&] { sum_col_api<float>(ctx, c, sum, sumsq); });
}
int salg_sum_col_f64(salg_ctx* ctx, const salg_csr* c, double* sum, double* sumsq) {
    return guarded([&] { sum_col_api<double>(ctx, c, sum, sumsq); });
}
int salg_sum_row_f32(salg_ctx* ctx, const salg_csr* c, float* out) {
    return guarded([&] { sum_row_api<float>(ctx, c, out); });
}
int salg_sum_row_f64(salg_ctx* ctx, const salg_csr* c, double* out) {
    return guarded([&] { sum_row_api<double>(ctx, c, out); });
}

/* CSC twins: the handle holds the CSR of A^T (csc_upload), so a column of A is a stored row */
int salg_csc_sum_col_f32(salg_ctx* ctx, const salg_csr* csc, float* sum, float* sumsq) {
    return guarded([&] { csc_sum_col_api<float>(ctx, csc, sum, sumsq); });
}
int salg_csc_sum_col_f64(salg_ctx* ctx, const salg_csr* csc, double* sum, double* sumsq) {
    return guarded([&] { csc_sum_col_api<double>(ctx, csc, sum, sumsq); });
}
int salg_csc_sum_row_f32(salg_ctx* ctx, const salg_csr* csc, float* out) {
    return guarded([&] { sum_col_api<float>(ctx, csc, out, nullptr); });
}
int salg_csc_sum_row_f64(salg_ctx* ctx, const salg_csr* csc, double* out) {
    return guarded([&] { sum_col_api<double>(ctx, csc, out, nullptr); });
}

int salg_col_stats_f64(salg_ctx* ctx, const salg_csr* c, double* sum, double* sumsq, double* nnz_col,
                       double* var_col) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        int64_t n = c->ncols;
        if (n == 0) return;
        DevBuf<double> d_sum((size_t)n, st), d_sq((size_t)n, st), d_cnt((size_t)(nnz_col ? n : 0), st);
        if (c->dtype == SALG_F64)
            col_stats_device<double>(ctx, c, d_sum.get(), d_sq.get(), nnz_col ? d_cnt.get() : nullptr);
        else
            col_stats_device<float>(ctx, c, d_sum.get(), d_sq.get(), nnz_col ? d_cnt.get() : nullptr);
        if (sum) SALG_CUDA(cudaMemcpyAsync(sum, d_sum.get(), (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (sumsq) SALG_CUDA(cudaMemcpyAsync(sumsq, d_sq.get(), (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (nnz_col) SALG_CUDA(cudaMemcpyAsync(nnz_col, d_cnt.get(), (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (var_col) {
            int64_t n_rows = global_nrows(ctx, c->nrows);
            DevBuf<double> d_var((size_t)n, st);
            var_col_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_sum.get(), d_sq.get(), (double)n_rows,
                                                                       d_var.get(), n);
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
            SALG_CUDA(cudaMemcpyAsync(var_col, d_var.get(), (size_t)n * 8, cudaMemcpyDeviceToHost, st));
            SALG_CUDA(cudaStreamSynchronize(st));
        }
        SALG_CUDA(cudaStreamSynchronize(st));
    });
}

}  // extern "C"
