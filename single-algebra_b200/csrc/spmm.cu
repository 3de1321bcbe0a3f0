// spmm.cu — sparse x dense-panel products of the randomized-SVD power iteration:
//     out = S * X - alpha * corr^T
// with S a CSR operand (A itself, or the transposed copy for A^T Y), X a (ncols(S) x 64) row-major panel
// and the rank-1 term carrying the implicit mean-centring: A_c X = A X - 1 (mu^T X),
// A_c^T Y = A^T Y - mu (1^T Y)  (single-svdlib randomized_svd; call sites pca/sparse/mod.rs:170-180,
// pca/sparse_masked/mod.rs:341-351).  The dense centred matrix is never formed.
//
// Work decomposition (merge-path flavour): the entry stream is cut into fixed chunks of SPMM_CHUNK stored
// entries, one warp per chunk, so skewed rows (gene rows of the transposed copy hold 10^2..10^5 entries)
// cost the same per warp.  A lane owns two adjacent panel columns; per entry the warp issues ONE coalesced
// 256 B (f32) / 512 B (f64) panel-row load and 2 FMAs per lane.  Rows completely inside a chunk are stored
// directly; rows cut by a chunk border are accumulated with atomics onto a pre-initialised output row.
#include <type_traits>

#include "common.cuh"

namespace salg {

template <typename T> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };

// chunk -> row containing its first entry (largest r with ptr[r] <= chunk*CH)
__global__ void chunk_row_kernel(const int64_t* __restrict__ ptr, int64_t nr, int64_t n_chunks, int ch,
                                 uint32_t* __restrict__ chunk_row) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    int64_t p = i * ch;
    int64_t lo = 0, hi = nr;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (ptr[mid] <= p) lo = mid; else hi = mid;
    }
    chunk_row[i] = (uint32_t)lo;
}

uint32_t* build_chunk_rows(salg_ctx* ctx, const int64_t* ptr, int64_t nr, int64_t nnz) {
    int64_t n_chunks = ceil_div(nnz, SPMM_CHUNK);
    uint32_t* out = (uint32_t*)dev_alloc(ctx, (size_t)(n_chunks + 1) * 4);
    if (n_chunks) {
        chunk_row_kernel<<<(unsigned)ceil_div(n_chunks, 256), 256, 0, ctx->stream>>>(ptr, nr, n_chunks, SPMM_CHUNK, out);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    return out;
}

// Rows that are empty or cut by a chunk border receive -alpha*corr (or 0) before the product kernel adds
// into them; rows owned by exactly one chunk are written by that chunk alone.
template <typename T>
__global__ void spmm_init_kernel(const int64_t* __restrict__ ptr, int64_t nr, T* __restrict__ out,
                                 const T* __restrict__ alpha, const double* __restrict__ corr) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double c0 = corr ? corr[2 * lane] : 0.0, c1 = corr ? corr[2 * lane + 1] : 0.0;
    using V = typename V2<T>::type;
    for (int64_t r = w; r < nr; r += nw) {
        int64_t s = ptr[r], e = ptr[r + 1];
        bool single = (e > s) && (s / SPMM_CHUNK == (e - 1) / SPMM_CHUNK);
        if (single) continue;
        double a = alpha ? (double)alpha[r] : 1.0;
        V v;
        v.x = (T)(-a * c0);
        v.y = (T)(-a * c1);
        reinterpret_cast<V*>(out + r * LP)[lane] = v;
    }
}

template <typename T, bool PATTERN>
__global__ void __launch_bounds__(256)
spmm_chunk_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ idx, const T* __restrict__ val,
                  const uint32_t* __restrict__ chunk_row, int64_t nr, int64_t nnz, int64_t n_chunks,
                  const T* __restrict__ X, T* __restrict__ out, const T* __restrict__ alpha,
                  const double* __restrict__ corr) {
    using V = typename V2<T>::type;
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const V* __restrict__ X2 = reinterpret_cast<const V*>(X);
    const double c0 = corr ? corr[2 * lane] : 0.0, c1 = corr ? corr[2 * lane + 1] : 0.0;

    for (int64_t chunk = w0; chunk < n_chunks; chunk += nw) {
        const int64_t s = chunk * SPMM_CHUNK;
        const int64_t e = (s + SPMM_CHUNK < nnz) ? s + SPMM_CHUNK : nnz;
        int64_t row = chunk_row[chunk];
        int64_t row_start = ptr[row];
        int64_t row_end = ptr[row + 1];
        T a0 = T(0), a1 = T(0);

        for (int64_t base = s; base < e; base += 32) {
            const int64_t p = base + lane;
            uint32_t c = 0;
            T v = T(0);
            if (p < e) {
                c = __ldcs(idx + p);
                if (!PATTERN) v = __ldcs(val + p);
                else v = T(1);
            }
            const int n = (e - base < 32) ? (int)(e - base) : 32;
            int k = 0;
            while (k < n) {
                int64_t lim = row_end - base;
                const int seg_end = lim < n ? (int)lim : n;
                for (; k + 4 <= seg_end; k += 4) {
                    uint32_t ck0 = __shfl_sync(0xFFFFFFFFu, c, k), ck1 = __shfl_sync(0xFFFFFFFFu, c, k + 1);
                    uint32_t ck2 = __shfl_sync(0xFFFFFFFFu, c, k + 2), ck3 = __shfl_sync(0xFFFFFFFFu, c, k + 3);
                    T v0 = __shfl_sync(0xFFFFFFFFu, v, k), v1 = __shfl_sync(0xFFFFFFFFu, v, k + 1);
                    T v2 = __shfl_sync(0xFFFFFFFFu, v, k + 2), v3 = __shfl_sync(0xFFFFFFFFu, v, k + 3);
                    V x0 = __ldg(X2 + (size_t)ck0 * (LP / 2) + lane);
                    V x1 = __ldg(X2 + (size_t)ck1 * (LP / 2) + lane);
                    V x2 = __ldg(X2 + (size_t)ck2 * (LP / 2) + lane);
                    V x3 = __ldg(X2 + (size_t)ck3 * (LP / 2) + lane);
                    a0 = fma(v0, x0.x, a0); a1 = fma(v0, x0.y, a1);
                    a0 = fma(v1, x1.x, a0); a1 = fma(v1, x1.y, a1);
                    a0 = fma(v2, x2.x, a0); a1 = fma(v2, x2.y, a1);
                    a0 = fma(v3, x3.x, a0); a1 = fma(v3, x3.y, a1);
                }
                for (; k < seg_end; k++) {
                    uint32_t ck = __shfl_sync(0xFFFFFFFFu, c, k);
                    T vk = __shfl_sync(0xFFFFFFFFu, v, k);
                    V x = __ldg(X2 + (size_t)ck * (LP / 2) + lane);
                    a0 = fma(vk, x.x, a0); a1 = fma(vk, x.y, a1);
                }
                if (base + k == row_end) {
                    // the row ends here: flush it
                    V* o = reinterpret_cast<V*>(out + row * LP) + lane;
                    if (row_start >= s) {   // whole row inside this chunk -> sole owner
                        double a = alpha ? (double)alpha[row] : 1.0;
                        V r;
                        r.x = (T)((double)a0 - a * c0);
                        r.y = (T)((double)a1 - a * c1);
                        *o = r;
                    } else {
                        atomicAdd(&o->x, a0);
                        atomicAdd(&o->y, a1);
                    }
                    a0 = T(0); a1 = T(0);
                    if (base + k < e) {
                        // next non-empty row (the entry at base+k belongs to it)
                        do {
                            row++;
                            row_start = row_end;
                            row_end = ptr[row + 1];
                        } while (row_end <= base + k);
                    }
                }
            }
        }
        if (row_end > e) {   // the chunk border cuts the last row: partial contribution
            V* o = reinterpret_cast<V*>(out + row * LP) + lane;
            atomicAdd(&o->x, a0);
            atomicAdd(&o->y, a1);
        }
    }
}

template <typename T>
void spmm_launch(salg_ctx* ctx, int prof_cls, const int64_t* ptr, const uint32_t* idx, const T* val,
                 const uint32_t* chunk_row, int64_t nr, int64_t nc, int64_t nnz, const T* X, T* out,
                 const T* alpha, const double* corr) {
    if (nr == 0) return;
    cudaStream_t st = ctx->stream;
    double bytes = (double)nnz * (sizeof(T) + 4) + (double)(nr + 1) * 8 + (double)nc * 60 * sizeof(T) +
                   (double)nr * 60 * sizeof(T);
    ProfScope ps(ctx, prof_cls, bytes);
    {
        int64_t want = ceil_div(nr * 32, 256);
        int64_t cap = (int64_t)ctx->sm_count * 16;
        spmm_init_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(ptr, nr, out, alpha, corr);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    int64_t n_chunks = ceil_div(nnz, SPMM_CHUNK);
    if (n_chunks == 0) return;
    int64_t want = ceil_div(n_chunks, 8);
    int64_t cap = (int64_t)ctx->sm_count * 8;   // 8 CTAs x 8 warps = 64 resident warps per SM
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (val)
        spmm_chunk_kernel<T, false><<<grid, 256, 0, st>>>(ptr, idx, val, chunk_row, nr, nnz, n_chunks, X, out, alpha, corr);
    else
        spmm_chunk_kernel<T, true><<<grid, 256, 0, st>>>(ptr, idx, val, chunk_row, nr, nnz, n_chunks, X, out, alpha, corr);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void spmm_launch<float>(salg_ctx*, int, const int64_t*, const uint32_t*, const float*, const uint32_t*,
                                 int64_t, int64_t, int64_t, const float*, float*, const float*, const double*);
template void spmm_launch<double>(salg_ctx*, int, const int64_t*, const uint32_t*, const double*, const uint32_t*,
                                  int64_t, int64_t, int64_t, const double*, double*, const double*, const double*);

template <typename T>
void spmm_A(salg_ctx* ctx, const salg_csr* c, const T* X, T* out, const double* corr, bool pattern) {
    if constexpr (std::is_same<T, float>::value) {
        // (shards of 2^31 or more stored entries have no tile format: A X falls back to the chunk kernel, whose offsets
        // are 64-bit; A^T Y needs the transposed copy or the tile format, both limited to < 2^31 entries per GPU)
        if (!pattern && tc_enabled(ctx) && (ctx->spmm_impl == 2 || c->nnz < ((int64_t)1 << 31))) {
            tc_spmm_A(ctx, c, X, out, corr);
            return;
        }
    }
    if (!c->chunk_row) c->chunk_row = build_chunk_rows(ctx, c->row_ptr, c->nrows, c->nnz);
    spmm_launch<T>(ctx, PROF_SPMM, c->row_ptr, c->col, pattern ? nullptr : (const T*)c->val, c->chunk_row, c->nrows,
                   c->ncols, c->nnz, X, out, nullptr, corr);
}
template void spmm_A<float>(salg_ctx*, const salg_csr*, const float*, float*, const double*, bool);
template void spmm_A<double>(salg_ctx*, const salg_csr*, const double*, double*, const double*, bool);

template <typename T>
void spmm_At(salg_ctx* ctx, const salg_csr* c, const T* Y, T* out, const T* mu, const double* corr) {
    if constexpr (std::is_same<T, float>::value) {
        if (tc_enabled(ctx)) {
            tc_spmm_At(ctx, c, Y, out, mu, corr);
            return;
        }
    }
    csr_ensure_transpose<T>(ctx, c);
    if (!c->t_chunk_row) c->t_chunk_row = build_chunk_rows(ctx, c->t_ptr, c->ncols, c->nnz);
    spmm_launch<T>(ctx, PROF_SPMMT, c->t_ptr, c->t_idx, (const T*)c->t_val, c->t_chunk_row, c->ncols, c->nrows, c->nnz,
                   Y, out, mu, corr);
}
template void spmm_At<float>(salg_ctx*, const salg_csr*, const float*, float*, const float*, const double*);
template void spmm_At<double>(salg_ctx*, const salg_csr*, const double*, double*, const double*, const double*);

}  // namespace salg
