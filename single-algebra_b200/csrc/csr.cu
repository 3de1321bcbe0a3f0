// csr.cu — device CSR container: upload (usize -> u32 narrowing, validation), download, column
// selection (MaskedCSRMatrix as a materialised compaction), transposed copy for the gather-form
// A^T products.  Replaces nalgebra_sparse::CsrMatrix<T> storage and
// single-svdlib's lanczos::masked::MaskedCSRMatrix (call site pca/sparse_masked/mod.rs:313).
#include <cub/cub.cuh>

#include "common.cuh"

namespace salg {

static size_t dsize(int dtype) { return dtype == SALG_F64 ? 8 : 4; }

salg_csr* csr_alloc(salg_ctx* ctx, int dtype, int64_t nrows, int64_t ncols, int64_t nnz) {
    salg_csr* c = new salg_csr();
    c->ctx = ctx;
    c->dtype = dtype;
    c->nrows = nrows;
    c->ncols = ncols;
    c->nnz = nnz;
    try {
        c->row_ptr = (int64_t*)dev_alloc(ctx, (size_t)(nrows + 1) * sizeof(int64_t));
        // +16 entries of slack so vectorised tail loads never leave the allocation
        c->col = (uint32_t*)dev_alloc(ctx, ((size_t)nnz + 16) * sizeof(uint32_t));
        c->val = dev_alloc(ctx, ((size_t)nnz + 16) * dsize(dtype));
    } catch (...) {
        csr_destroy(c);
        throw;
    }
    return c;
}

void csr_invalidate_transpose(const salg_csr* c) {
    dev_free(c->ctx, c->t_ptr);
    dev_free(c->ctx, c->t_idx);
    dev_free(c->ctx, c->t_val);
    dev_free(c->ctx, c->t_chunk_row);
    if (c->tc) tc_free(c->ctx, c->tc);
    c->tc = nullptr;
    c->t_chunk_row = nullptr;
    c->t_ptr = nullptr;
    c->t_idx = nullptr;
    c->t_val = nullptr;
    c->t_valid = false;
}

void csr_destroy(salg_csr* c) {
    if (!c) return;
    if (ctx_alive(c->ctx)) cudaSetDevice(c->ctx->device);
    csr_invalidate_transpose(c);
    dev_free(c->ctx, c->row_ptr);
    dev_free(c->ctx, c->col);
    dev_free(c->ctx, c->val);
    dev_free(c->ctx, c->chunk_row);
    delete c;
}

void exclusive_scan_i64(salg_ctx* ctx, const int64_t* in, int64_t* out, int64_t n) {
    size_t tmp_bytes = 0;
    SALG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, ctx->stream));
    DevBuf<uint8_t> tmp(tmp_bytes + 16, ctx->stream);
    SALG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tmp_bytes, in, out, n, ctx->stream));
}

// ---- upload -----------------------------------------------------------------------------------------
template <typename I>
__global__ void narrow_idx_kernel(const I* __restrict__ src, uint32_t* __restrict__ dst, int64_t n,
                                  uint64_t ncols, int* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) {
        uint64_t v = (uint64_t)src[i];
        bad |= (v >= ncols);
        dst[i] = (uint32_t)v;
    }
    if (bad) atomicOr(flag, 1);
}

template <typename I>
__global__ void copy_offsets_kernel(const I* __restrict__ src, int64_t* __restrict__ dst, int64_t n1,
                                    int64_t nnz, int* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    int64_t v = (int64_t)src[i];
    dst[i] = v;
    bool bad = v < 0 || v > nnz;
    if (i == 0) bad |= (v != 0);
    if (i == n1 - 1) bad |= (v != nnz);
    if (i + 1 < n1) bad |= ((int64_t)src[i + 1] < v);
    if (bad) atomicOr(flag, 2);
}

// one warp per row: column indices strictly increasing inside the row
__global__ void check_sorted_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                                    int64_t nrows, int64_t nnz, int* __restrict__ flag) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (int64_t r = w; r < nrows; r += nw) {
        int64_t s = ptr[r], e = ptr[r + 1];
        if (s < 0 || e > nnz || e < s) continue;   // malformed offsets are reported by copy_offsets_kernel
        for (int64_t p = s + lane; p + 1 < e; p += 32) bad |= (col[p] >= col[p + 1]);
    }
    if (bad) atomicOr(flag, 4);
}

static int grid_for(salg_ctx* ctx, int64_t n, int block, int per_sm = 8) {
    int64_t want = ceil_div(n, block);
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <typename T, typename OffT, typename IdxT>
static salg_csr* csr_upload(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const OffT* h_off,
                            const IdxT* h_idx, const T* h_val) {
    SALG_REQUIRE(ctx, SALG_ERR_BAD_ARG, "ctx is NULL");
    SALG_REQUIRE(nrows >= 0 && ncols >= 0 && nnz >= 0, SALG_ERR_BAD_ARG, "negative dimension");
    SALG_REQUIRE(ncols < (int64_t)0xFFFFFFFFLL, SALG_ERR_BAD_ARG, "ncols must fit 32 bits on the device");
    SALG_REQUIRE(h_off, SALG_ERR_BAD_ARG, "row_offsets is NULL");
    SALG_REQUIRE(nnz == 0 || (h_idx && h_val), SALG_ERR_BAD_ARG, "col_indices/values is NULL");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    salg_csr* c = csr_alloc(ctx, dtype_of<T>::value, nrows, ncols, nnz);
    try {
        ProfScope ps(ctx, PROF_H2D,
                     (double)nnz * (sizeof(T) + sizeof(IdxT)) + (double)(nrows + 1) * sizeof(OffT));
        DevBuf<int> flag(1, st);
        SALG_CUDA(cudaMemsetAsync(flag.get(), 0, sizeof(int), st));
        // values: straight copy (pinned source => DMA at link rate; pageable => driver-staged)
        if (nnz) SALG_CUDA(cudaMemcpyAsync(c->val, h_val, (size_t)nnz * sizeof(T), cudaMemcpyHostToDevice, st));
        // offsets
        {
            DevBuf<OffT> tmp((size_t)nrows + 1, st);
            SALG_CUDA(cudaMemcpyAsync(tmp.get(), h_off, (size_t)(nrows + 1) * sizeof(OffT),
                                      cudaMemcpyHostToDevice, st));
            int64_t n1 = nrows + 1;
            copy_offsets_kernel<OffT><<<(unsigned)ceil_div(n1, 256), 256, 0, st>>>(tmp.get(), c->row_ptr, n1, nnz,
                                                                                 flag.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        // indices: bounded device staging, narrowed to u32 chunk by chunk
        if (nnz) {
            const int64_t CH = (int64_t)1 << 26;  // 64 Mi entries per chunk
            int64_t ch = nnz < CH ? nnz : CH;
            DevBuf<IdxT> stage0((size_t)ch, st), stage1((size_t)(nnz > ch ? ch : 0), st);
            IdxT* stg[2] = {stage0.get(), stage1.get() ? stage1.get() : stage0.get()};
            int k = 0;
            for (int64_t o = 0; o < nnz; o += ch, k ^= 1) {
                int64_t n = nnz - o < ch ? nnz - o : ch;
                SALG_CUDA(cudaMemcpyAsync(stg[k], h_idx + o, (size_t)n * sizeof(IdxT), cudaMemcpyHostToDevice, st));
                narrow_idx_kernel<IdxT><<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(stg[k], c->col + o, n,
                                                                                  (uint64_t)ncols, flag.get());
                ctx->n_launch++;
                SALG_CUDA(cudaGetLastError());
            }
            SALG_CUDA(cudaMemsetAsync(c->col + nnz, 0, 16 * sizeof(uint32_t), st));
            SALG_CUDA(cudaMemsetAsync((char*)c->val + (size_t)nnz * sizeof(T), 0, 16 * sizeof(T), st));
            check_sorted_kernel<<<grid_for(ctx, nrows * 32, 256, 16), 256, 0, st>>>(c->row_ptr, c->col, nrows,
                                                                                  nnz, flag.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        int h_flag = 0;
        SALG_CUDA(cudaMemcpyAsync(&h_flag, flag.get(), sizeof(int), cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        if (h_flag & 2)
            throw Error(SALG_ERR_BAD_ARG, "invalid CSR: row offsets must start at 0, be non-decreasing and end at nnz");
        if (h_flag & 1) throw Error(SALG_ERR_BAD_ARG, "invalid CSR: column index out of range");
        if (h_flag & 4)
            throw Error(SALG_ERR_BAD_ARG, "invalid CSR: column indices must be strictly increasing within a row");
    } catch (...) {
        cudaStreamSynchronize(st);
        csr_destroy(c);
        throw;
    }
    return c;
}

// helpers for the streamed fit (pca.cu)
salg_csr* csr_upload_i32_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* off, const int32_t* idx,
                             const float* val) {
    return csr_upload<float, int64_t, uint32_t>(ctx, nrows, ncols, nnz, off, (const uint32_t*)idx, val);
}
// "shell" of a host-resident matrix: validated row offsets on the device, no entry arrays
salg_csr* csr_shell_from_host_offsets(salg_ctx* ctx, int dtype, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* h_off) {
    SALG_REQUIRE(ctx && h_off, SALG_ERR_BAD_ARG, "ctx/row_offsets is NULL");
    SALG_REQUIRE(nrows >= 0 && ncols >= 0 && nnz >= 0, SALG_ERR_BAD_ARG, "negative dimension");
    SALG_REQUIRE(ncols < (int64_t)0xFFFFFFFFLL, SALG_ERR_BAD_ARG, "ncols must fit 32 bits on the device");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    salg_csr* c = new salg_csr();
    c->ctx = ctx;
    c->dtype = dtype;
    c->nrows = nrows;
    c->ncols = ncols;
    c->nnz = nnz;
    try {
        c->row_ptr = (int64_t*)dev_alloc(ctx, (size_t)(nrows + 1) * 8);
        DevBuf<int64_t> tmp((size_t)nrows + 1, st);
        DevBuf<int> flag(1, st);
        SALG_CUDA(cudaMemsetAsync(flag.get(), 0, 4, st));
        SALG_CUDA(cudaMemcpyAsync(tmp.get(), h_off, (size_t)(nrows + 1) * 8, cudaMemcpyHostToDevice, st));
        copy_offsets_kernel<int64_t><<<(unsigned)ceil_div(nrows + 1, 256), 256, 0, st>>>(tmp.get(), c->row_ptr, nrows + 1, nnz,
                                                                                        flag.get());
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        int h = 0;
        SALG_CUDA(cudaMemcpyAsync(&h, flag.get(), 4, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        if (h) throw Error(SALG_ERR_BAD_ARG, "invalid CSR: row offsets must start at 0, be non-decreasing and end at nnz");
    } catch (...) {
        cudaStreamSynchronize(st);
        csr_destroy(c);
        throw;
    }
    return c;
}

__global__ void widen_idx_kernel(const uint32_t* __restrict__ src, uint64_t* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

template <typename T>
static void csr_download(salg_ctx* ctx, const salg_csr* c, uint64_t* off, uint64_t* idx, T* val) {
    SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
    SALG_REQUIRE(c->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "csr value type does not match the entry point");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (off) {
        static_assert(sizeof(int64_t) == sizeof(uint64_t), "");
        SALG_CUDA(cudaMemcpyAsync(off, c->row_ptr, (size_t)(c->nrows + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    if (val && c->nnz)
        SALG_CUDA(cudaMemcpyAsync(val, c->val, (size_t)c->nnz * sizeof(T), cudaMemcpyDeviceToHost, st));
    if (idx && c->nnz) {
        const int64_t CH = (int64_t)1 << 26;
        int64_t ch = c->nnz < CH ? c->nnz : CH;
        DevBuf<uint64_t> stage((size_t)ch, st);
        for (int64_t o = 0; o < c->nnz; o += ch) {
            int64_t n = c->nnz - o < ch ? c->nnz - o : ch;
            widen_idx_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(c->col + o, stage.get(), n);
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
            SALG_CUDA(cudaMemcpyAsync(idx + o, stage.get(), (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        }
    }
    SALG_CUDA(cudaStreamSynchronize(st));
}

// ---- column selection -----------------------------------------------------------------------------------
// map[c] = compact id of column c (rank among kept columns) or 0xFFFFFFFF when dropped
__global__ void build_colmap_kernel(const uint8_t* __restrict__ mask, const int64_t* __restrict__ rank,
                                    uint32_t* __restrict__ map, int64_t ncols) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncols) map[i] = mask[i] ? (uint32_t)rank[i] : 0xFFFFFFFFu;
}

__global__ void mask_to_i64_kernel(const uint8_t* __restrict__ mask, int64_t* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = mask[i] ? 1 : 0;
}

// pass 1: kept entries per row (one warp per row)
__global__ void count_kept_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                                  const uint32_t* __restrict__ map, int64_t nrows, int64_t* __restrict__ cnt) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < nrows; r += nw) {
        int64_t s = ptr[r], e = ptr[r + 1];
        int n = 0;
        for (int64_t p = s + lane; p < e; p += 32) n += (map[col[p]] != 0xFFFFFFFFu);
#pragma unroll
        for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
        if (lane == 0) cnt[r] = n;
    }
    if (w == 0 && lane == 0) cnt[nrows] = 0;
}

// pass 2: ordered per-row compaction (ballot prefix keeps the within-row order)
template <typename T>
__global__ void write_kept_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                                  const T* __restrict__ val, const uint32_t* __restrict__ map, int64_t nrows,
                                  const int64_t* __restrict__ new_ptr, uint32_t* __restrict__ new_col,
                                  T* __restrict__ new_val) {
    int lane = threadIdx.x & 31;
    unsigned lt = (1u << lane) - 1u;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < nrows; r += nw) {
        int64_t s = ptr[r], e = ptr[r + 1];
        int64_t o = new_ptr[r];
        for (int64_t base = s; base < e; base += 128) {      // 4 independent index/map loads in flight per lane
            uint32_t m[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int64_t p = base + 32 * u + lane;
                m[u] = (p < e) ? map[__ldcs(col + p)] : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int64_t p = base + 32 * u + lane;
                bool keep = (m[u] != 0xFFFFFFFFu);
                unsigned b = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) {
                    int64_t q = o + __popc(b & lt);
                    new_col[q] = m[u];
                    new_val[q] = val[p];
                }
                o += __popc(b);
            }
        }
    }
}

template <typename T>
salg_csr* csr_select_columns(salg_ctx* ctx, const salg_csr* c, const uint8_t* mask_host, int64_t* d_row_kept) {
    cudaStream_t st = ctx->stream;
    int64_t ncols = c->ncols, nrows = c->nrows;
    DevBuf<uint8_t> d_mask((size_t)ncols + 1, st);
    DevBuf<int64_t> d_m64((size_t)ncols + 1, st), d_rank((size_t)ncols + 1, st);
    DevBuf<uint32_t> d_map((size_t)ncols + 1, st);
    SALG_CUDA(cudaMemsetAsync(d_mask.get(), 0, (size_t)ncols + 1, st));
    if (ncols) SALG_CUDA(cudaMemcpyAsync(d_mask.get(), mask_host, (size_t)ncols, cudaMemcpyHostToDevice, st));
    mask_to_i64_kernel<<<(unsigned)ceil_div(ncols + 1, 256), 256, 0, st>>>(d_mask.get(), d_m64.get(), ncols + 1);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    exclusive_scan_i64(ctx, d_m64.get(), d_rank.get(), ncols + 1);
    int64_t n_keep = 0;
    SALG_CUDA(cudaMemcpyAsync(&n_keep, d_rank.get() + ncols, 8, cudaMemcpyDeviceToHost, st));
    if (ncols) {
        build_colmap_kernel<<<(unsigned)ceil_div(ncols, 256), 256, 0, st>>>(d_mask.get(), d_rank.get(), d_map.get(),
                                                                            ncols);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    DevBuf<int64_t> d_cnt((size_t)nrows + 1, st);
    int64_t* new_ptr = (int64_t*)dev_alloc(ctx, (size_t)(nrows + 1) * 8);
    salg_csr* out = nullptr;
    try {
        double bytes = (double)c->nnz * (sizeof(T) + 4) + 2.0 * (double)(nrows + 1) * 8;
        ProfScope ps(ctx, PROF_COMPACT, bytes);
        if (!d_row_kept) {
            count_kept_kernel<<<grid_for(ctx, (nrows + 1) * 32, 256, 8), 256, 0, st>>>(c->row_ptr, c->col, d_map.get(),
                                                                                       nrows, d_cnt.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        exclusive_scan_i64(ctx, d_row_kept ? d_row_kept : d_cnt.get(), new_ptr, nrows + 1);
        int64_t nnz_eff = 0;
        SALG_CUDA(cudaMemcpyAsync(&nnz_eff, new_ptr + nrows, 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        out = new salg_csr();
        out->ctx = ctx;
        out->dtype = c->dtype;
        out->nrows = nrows;
        out->ncols = n_keep;
        out->nnz = nnz_eff;
        out->row_ptr = new_ptr;
        new_ptr = nullptr;
        out->col = (uint32_t*)dev_alloc(ctx, ((size_t)nnz_eff + 16) * 4);
        out->val = dev_alloc(ctx, ((size_t)nnz_eff + 16) * sizeof(T));
        SALG_CUDA(cudaMemsetAsync(out->col + nnz_eff, 0, 16 * 4, st));
        SALG_CUDA(cudaMemsetAsync((T*)out->val + nnz_eff, 0, 16 * sizeof(T), st));
        if (nrows && c->nnz) {
            write_kept_kernel<T><<<grid_for(ctx, nrows * 32, 256, 8), 256, 0, st>>>(
                c->row_ptr, c->col, (const T*)c->val, d_map.get(), nrows, out->row_ptr, out->col, (T*)out->val);
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
    } catch (...) {
        dev_free(ctx, new_ptr);
        if (out) csr_destroy(out);
        throw;
    }
    return out;
}
template salg_csr* csr_select_columns<float>(salg_ctx*, const salg_csr*, const uint8_t*, int64_t*);
template salg_csr* csr_select_columns<double>(salg_ctx*, const salg_csr*, const uint8_t*, int64_t*);

// ---- transposed copy ------------------------------------------------------------------------------------
// CSR of A^T (== CSC of A): stable radix sort of the entries by column id keeps rows ascending inside
// each column, so the copy is deterministic and the A^T Y product needs no atomics.
__global__ void iota_kernel(uint32_t* __restrict__ a, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] = (uint32_t)i;
}

__global__ void col_hist_kernel(const uint32_t* __restrict__ col, int64_t nnz, unsigned long long* __restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) atomicAdd(&cnt[col[i]], 1ULL);
}

template <typename T>
__global__ void transpose_fill_kernel(const uint32_t* __restrict__ perm, const int64_t* __restrict__ ptr,
                                      int64_t nrows, const T* __restrict__ val, int64_t nnz,
                                      uint32_t* __restrict__ t_idx, T* __restrict__ t_val) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        int64_t p = perm[i];
        // row of entry p: largest r with ptr[r] <= p
        int64_t lo = 0, hi = nrows;
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (ptr[mid] <= p) lo = mid; else hi = mid;
        }
        t_idx[i] = (uint32_t)lo;
        t_val[i] = val[p];
    }
}

template <typename T>
void csr_ensure_transpose(salg_ctx* ctx, const salg_csr* c) {
    if (c->t_valid) return;
    cudaStream_t st = ctx->stream;
    SALG_REQUIRE(c->nnz < ((int64_t)1 << 31), SALG_ERR_UNSUPPORTED,
                 "transposed copy supports < 2^31 stored entries per GPU shard");
    SALG_REQUIRE(c->nrows < ((int64_t)1 << 32), SALG_ERR_UNSUPPORTED, "row count must fit 32 bits");
    int64_t nnz = c->nnz, ncols = c->ncols;
    ProfScope ps(ctx, PROF_TRANSPOSE, (double)nnz * 2.0 * (sizeof(T) + 4));
    c->t_ptr = (int64_t*)dev_alloc(ctx, (size_t)(ncols + 1) * 8);
    c->t_idx = (uint32_t*)dev_alloc(ctx, ((size_t)nnz + 16) * 4);
    c->t_val = dev_alloc(ctx, ((size_t)nnz + 16) * sizeof(T));
    SALG_CUDA(cudaMemsetAsync(c->t_idx + nnz, 0, 16 * 4, st));
    SALG_CUDA(cudaMemsetAsync((T*)c->t_val + nnz, 0, 16 * sizeof(T), st));
    // column histogram -> t_ptr
    {
        DevBuf<int64_t> cnt((size_t)ncols + 1, st);
        SALG_CUDA(cudaMemsetAsync(cnt.get(), 0, (size_t)(ncols + 1) * 8, st));
        if (nnz) {
            col_hist_kernel<<<grid_for(ctx, nnz, 256, 16), 256, 0, st>>>(c->col, nnz,
                                                                         (unsigned long long*)cnt.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        exclusive_scan_i64(ctx, cnt.get(), c->t_ptr, ncols + 1);
    }
    if (nnz) {
        DevBuf<uint32_t> keys_out((size_t)nnz, st), pos_in((size_t)nnz, st), pos_out((size_t)nnz, st);
        iota_kernel<<<grid_for(ctx, nnz, 256, 16), 256, 0, st>>>(pos_in.get(), nnz);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        int end_bit = 1;
        while (end_bit < 32 && ((uint64_t)1 << end_bit) < (uint64_t)(ncols > 1 ? ncols : 1)) end_bit++;
        size_t tmp_bytes = 0;
        SALG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t*)c->col, keys_out.get(),
                                                  (const uint32_t*)pos_in.get(), pos_out.get(), (int)nnz, 0, end_bit, st));
        DevBuf<uint8_t> tmp(tmp_bytes + 16, st);
        SALG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tmp_bytes, (const uint32_t*)c->col, keys_out.get(),
                                                  (const uint32_t*)pos_in.get(), pos_out.get(), (int)nnz, 0, end_bit, st));
        transpose_fill_kernel<T><<<grid_for(ctx, nnz, 256, 16), 256, 0, st>>>(
            pos_out.get(), c->row_ptr, c->nrows, (const T*)c->val, nnz, c->t_idx, (T*)c->t_val);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    c->t_valid = true;
}
template void csr_ensure_transpose<float>(salg_ctx*, const salg_csr*);
template void csr_ensure_transpose<double>(salg_ctx*, const salg_csr*);

// stored entries per row: row_offsets[r + 1] - row_offsets[r]  (MatrixNonZero::nonzero_row, src/sparse/csr.rs:79-122)
__global__ void row_nnz_kernel(const int64_t* __restrict__ ptr, int64_t nrows, uint64_t* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrows) out[r] = (uint64_t)(ptr[r + 1] - ptr[r]);
}
__global__ void f64_to_u64_kernel(const double* __restrict__ in, int64_t n, uint64_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint64_t)llrint(in[i]);
}

}  // namespace salg

using namespace salg;

extern "C" {

int salg_csr_upload_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const uint64_t* off,
                        const uint64_t* idx, const float* val, salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = csr_upload<float, uint64_t, uint64_t>(ctx, nrows, ncols, nnz, off, idx, val);
    });
}
/* CscMatrix<T> (col_offsets, row_indices, values; src/sparse/csc.rs) is stored as the CSR of A^T: the returned handle has
 * ncols stored rows.  Same validation as the CSR upload (offsets monotone, indices strictly ascending and in range). */
int salg_csc_upload_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const uint64_t* col_off,
                        const uint64_t* row_idx, const float* val, salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = csr_upload<float, uint64_t, uint64_t>(ctx, ncols, nrows, nnz, col_off, row_idx, val);
    });
}
int salg_csc_upload_f64(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const uint64_t* col_off,
                        const uint64_t* row_idx, const double* val, salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = csr_upload<double, uint64_t, uint64_t>(ctx, ncols, nrows, nnz, col_off, row_idx, val);
    });
}
int salg_csr_upload_f64(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const uint64_t* off,
                        const uint64_t* idx, const double* val, salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = csr_upload<double, uint64_t, uint64_t>(ctx, nrows, ncols, nnz, off, idx, val);
    });
}
int salg_csr_upload_i32_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* off,
                            const int32_t* idx, const float* val, salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = csr_upload<float, int64_t, uint32_t>(ctx, nrows, ncols, nnz, off, (const uint32_t*)idx, val);
    });
}
int salg_csr_upload_i32_f64(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* off,
                            const int32_t* idx, const double* val, salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = csr_upload<double, int64_t, uint32_t>(ctx, nrows, ncols, nnz, off, (const uint32_t*)idx, val);
    });
}

int salg_csr_free(salg_csr* c) {
    return guarded([&] { csr_destroy(c); });
}

int salg_csr_dims(const salg_csr* c, int64_t* nrows, int64_t* ncols, int64_t* nnz, int* dtype) {
    return guarded([&] {
        SALG_REQUIRE(c, SALG_ERR_BAD_ARG, "csr is NULL");
        if (nrows) *nrows = c->nrows;
        if (ncols) *ncols = c->ncols;
        if (nnz) *nnz = c->nnz;
        if (dtype) *dtype = c->dtype;
    });
}

int salg_csr_download_f32(salg_ctx* ctx, const salg_csr* c, uint64_t* off, uint64_t* idx, float* val) {
    return guarded([&] { csr_download<float>(ctx, c, off, idx, val); });
}
int salg_csr_download_f64(salg_ctx* ctx, const salg_csr* c, uint64_t* off, uint64_t* idx, double* val) {
    return guarded([&] { csr_download<double>(ctx, c, off, idx, val); });
}

int salg_csr_download_raw(salg_ctx* ctx, const salg_csr* c, int64_t* off, uint32_t* idx, void* val) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        size_t es = c->dtype == SALG_F64 ? 8 : 4;
        if (off) SALG_CUDA(cudaMemcpyAsync(off, c->row_ptr, (size_t)(c->nrows + 1) * 8, cudaMemcpyDeviceToHost, st));
        if (idx && c->nnz) SALG_CUDA(cudaMemcpyAsync(idx, c->col, (size_t)c->nnz * 4, cudaMemcpyDeviceToHost, st));
        if (val && c->nnz) SALG_CUDA(cudaMemcpyAsync(val, c->val, (size_t)c->nnz * es, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
    });
}

int salg_csr_transpose(salg_ctx* ctx, const salg_csr* c, salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c && out, SALG_ERR_BAD_ARG, "ctx/csr/out is NULL");
        SALG_REQUIRE(ctx->nranks == 1, SALG_ERR_UNSUPPORTED, "transpose of a row-sharded matrix is not a local operation");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        if (c->dtype == SALG_F64) csr_ensure_transpose<double>(ctx, c);
        else csr_ensure_transpose<float>(ctx, c);
        salg_csr* t = csr_alloc(ctx, c->dtype, c->ncols, c->nrows, c->nnz);
        const size_t es = dsize(c->dtype);
        SALG_CUDA(cudaMemcpyAsync(t->row_ptr, c->t_ptr, (size_t)(c->ncols + 1) * 8, cudaMemcpyDeviceToDevice, st));
        SALG_CUDA(cudaMemcpyAsync(t->col, c->t_idx, ((size_t)c->nnz + 16) * 4, cudaMemcpyDeviceToDevice, st));
        SALG_CUDA(cudaMemcpyAsync(t->val, c->t_val, ((size_t)c->nnz + 16) * es, cudaMemcpyDeviceToDevice, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        *out = t;
    });
}

int salg_nonzero_row(salg_ctx* ctx, const salg_csr* c, uint64_t* out) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
        if (c->nrows == 0) return;                                   // src/sparse/csr.rs:87-89
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        DevBuf<uint64_t> d((size_t)c->nrows, st);
        row_nnz_kernel<<<(unsigned)ceil_div(c->nrows, 256), 256, 0, st>>>(c->row_ptr, c->nrows, d.get());
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        SALG_CUDA(cudaMemcpyAsync(out, d.get(), (size_t)c->nrows * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
    });
}

int salg_nonzero_col(salg_ctx* ctx, const salg_csr* c, uint64_t* out) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
        if (c->ncols == 0) return;
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        const size_t n = (size_t)c->ncols;
        DevBuf<double> d_sum(n, st), d_cnt(n, st);
        if (c->dtype == SALG_F64) col_stats_device<double>(ctx, c, d_sum.get(), nullptr, d_cnt.get());
        else col_stats_device<float>(ctx, c, d_sum.get(), nullptr, d_cnt.get());
        DevBuf<uint64_t> d(n, st);
        f64_to_u64_kernel<<<(unsigned)ceil_div((int64_t)n, 256), 256, 0, st>>>(d_cnt.get(), (int64_t)n, d.get());
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        SALG_CUDA(cudaMemcpyAsync(out, d.get(), n * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
    });
}

int salg_csr_values_clone(salg_ctx* ctx, const salg_csr* c, void** out) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c && out, SALG_ERR_BAD_ARG, "ctx/csr/out is NULL");
        SALG_CUDA(cudaSetDevice(ctx->device));
        const size_t bytes = ((size_t)c->nnz + 16) * dsize(c->dtype);
        void* p = dev_alloc(ctx, bytes);
        SALG_CUDA(cudaMemcpyAsync(p, c->val, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        *out = p;
    });
}

int salg_csr_values_restore(salg_ctx* ctx, salg_csr* c, const void* clone) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c && clone, SALG_ERR_BAD_ARG, "ctx/csr/clone is NULL");
        SALG_CUDA(cudaSetDevice(ctx->device));
        if (c->t_valid || c->tc) {                       // values change: cached transposed copy / tile format are stale
            SALG_CUDA(cudaStreamSynchronize(ctx->stream));
            csr_invalidate_transpose(c);
        }
        SALG_CUDA(cudaMemcpyAsync(c->val, clone, ((size_t)c->nnz + 16) * dsize(c->dtype), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    });
}

int salg_dev_free(salg_ctx* ctx, void* p) {
    return guarded([&] {
        SALG_REQUIRE(ctx, SALG_ERR_BAD_ARG, "ctx is NULL");
        dev_free(ctx, p);
    });
}

int salg_csr_select_columns(salg_ctx* ctx, const salg_csr* c, const uint8_t* mask, int64_t mask_len,
                            salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c && out, SALG_ERR_BAD_ARG, "ctx/csr/out is NULL");
        SALG_REQUIRE(mask || c->ncols == 0, SALG_ERR_BAD_ARG, "mask is NULL");
        SALG_REQUIRE(mask_len == c->ncols, SALG_ERR_MASK_LEN,
                     "The mask vector length and the number of features (columns) have to be the same!");
        SALG_CUDA(cudaSetDevice(ctx->device));
        *out = c->dtype == SALG_F64 ? csr_select_columns<double>(ctx, c, mask)
                                    : csr_select_columns<float>(ctx, c, mask);
    });
}

}  // extern "C"
