// pca.cu — SparsePCA::fit / MaskedSparsePCA::fit / transform / fit_transform on the device
// (src/dimred/pca/sparse/mod.rs:102-285, src/dimred/pca/sparse_masked/mod.rs:255-546) with the SVD engine
// the reference delegates to single-svdlib (randomized_svd, svd_flip, svd_las2, MaskedCSRMatrix; SURVEY
// App. B) rebuilt from the kernels in spmm.cu / dense.cu / lanczos.cu.
//
// Randomized schedule (same Krylov space as the reference, App. B.1 steps 4-6):
//   Y = A_c Om;  q x { Y <- orth(Y); Z = A_c^T Y; Z <- orth(Z); Y = A_c Z };  Q = orth(Y);  B^T = A_c^T Q
//   B^T = Q_B R_B (CholeskyQR2),  R_B = U_R S V_R^T (one-CTA Jacobi)  =>  V = Q_B U_R,  U = Q V_R.
// Row-sharded contexts all-reduce the Gram matrices, the column statistics and every n_eff-sized panel.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace salg {

__global__ void omega_normal_kernel(uint64_t seed, int64_t n, int l, float* outf, double* outd) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * LP) return;
    int c = (int)(i & 63);
    double v = 0.0;
    if (c < l) {
        uint64_t z = seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i * 0xD1B54A32D192ED03ULL + 0x8CB92BA72F3D8DD7ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        uint64_t z2 = z * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL;
        z2 = (z2 ^ (z2 >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z2 = (z2 ^ (z2 >> 27)) * 0x94D049BB133111EBULL;
        z2 = z2 ^ (z2 >> 31);
        double u1 = ((double)(z >> 11) + 1.0) / 9007199254740993.0;
        double u2 = (double)(z2 >> 11) / 9007199254740992.0;
        v = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    if (outf) outf[i] = (float)v;
    if (outd) outd[i] = v;
}

template <typename T>
__global__ void gather_mean_kernel(const double* __restrict__ sum, const uint32_t* __restrict__ kept, int64_t n_eff,
                                   double inv_n, int center, T* __restrict__ mu) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_eff) mu[i] = center ? (T)(sum[kept ? kept[i] : i] * inv_n) : T(0);
}

template <typename T>
__global__ void rowscale_panel_kernel(const T* __restrict__ P, int64_t m, const T* __restrict__ wT,
                                      const double* __restrict__ wD, T* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m * LP) return;
    int64_t r = i >> 6;
    double w = wT ? (double)wT[r] : wD[r];
    out[i] = (T)((double)P[i] * w);
}

template <typename T>
__global__ void panel_sub_kernel(T* __restrict__ a, const T* __restrict__ b, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] = a[i] - b[i];
}

// Streamed fit: the matrix stays in (pinned) host memory and crosses PCIe ONCE, chunk by chunk, through the masked
// statistics + fused compaction pass; only the kept entries (scaled scratch) and the tile format stay on the device.
struct HostCsrSrc {
    const int64_t* off;
    const int32_t* idx;
    const void* val;
};
struct StreamFallback {};      // thrown when a streamed fit has to be redone from a resident upload

// one warp per row of a staged chunk: column indices in range and strictly increasing (what csr_upload checks)
__global__ void validate_rows_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col, int64_t nrows,
                                     uint32_t ncols, int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int bad = 0;
    for (int64_t r = w; r < nrows; r += nw) {
        const int64_t s = ptr[r], e = ptr[r + 1];
        for (int64_t p = s + lane; p < e; p += 32) {
            const uint32_t c = col[p];
            if (c >= ncols) bad |= 1;
            if (p + 1 < e && c >= col[p + 1]) bad |= 4;
        }
    }
    if (bad) atomicOr(flag, bad);
}

template <typename T>
static void pca_release(salg_pca* p) {
    if (!p) return;
    dev_free(p->ctx, p->d_V);
    dev_free(p->ctx, p->d_mean);
    dev_free(p->ctx, p->d_scores);
    dev_free(p->ctx, p->d_tscores);
    delete p;
}

static void pca_destroy(salg_pca* p) {
    if (!p) return;
    if (ctx_alive(p->ctx)) cudaSetDevice(p->ctx->device);
    pca_release<float>(p);
}

template <typename T>
static salg_pca* pca_fit(salg_ctx* ctx, const salg_csr* x, const salg_pca_params* prm, const uint8_t* mask,
                         int64_t mask_len, const T* omega, int64_t omega_rows, int64_t omega_cols,
                         const HostCsrSrc* host = nullptr) {
    SALG_REQUIRE(ctx && x && prm, SALG_ERR_BAD_ARG, "ctx/x/params is NULL");
    SALG_REQUIRE(x->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "csr value type does not match the entry point");
    SALG_REQUIRE(prm->n_components >= 1, SALG_ERR_BAD_ARG, "n_components must be >= 1");
    SALG_REQUIRE(prm->svd_method == SALG_SVD_LANCZOS || prm->svd_method == SALG_SVD_RANDOM, SALG_ERR_BAD_ARG,
                 "unknown svd_method");
    if (mask)
        SALG_REQUIRE(mask_len == x->ncols, SALG_ERR_MASK_LEN,
                     "The mask vector length and the number of features (columns) have to be the same!");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t ncols = x->ncols;
    const int64_t n_total = global_nrows(ctx, x->nrows);
    SALG_REQUIRE(n_total >= 2, SALG_ERR_BAD_ARG, "need at least 2 samples (explained variance divides by n - 1)");
    SALG_REQUIRE(ncols >= 1, SALG_ERR_BAD_ARG, "matrix has no columns");
    const bool center = prm->center != 0;

    salg_pca* P = new salg_pca();
    salg_csr* compact = nullptr;
    try {
        P->ctx = ctx;
        P->dtype = dtype_of<T>::value;
        P->ncols = ncols;
        P->center = center;
        P->masked = mask != nullptr;
        P->n_samples = n_total;
        P->n_fit_rows_local = x->nrows;

        // ---- column statistics over ALL columns: mean_ and total_var (pca/sparse/mod.rs:106-131,
        //      pca/sparse_masked/mod.rs:275-311) — one pass instead of the reference's three
        DevBuf<double> d_sum((size_t)ncols, st), d_sq((size_t)ncols, st);
        DevBuf<int64_t> d_row_kept;
        uint32_t* kept_col = nullptr;     // context scratch (fused compaction)
        T* kept_val = nullptr;
        int kept_shift = 0;
        bool fused_compact = false;
        if (mask) {
            // the compaction's count pass rides on the statistics pass (one read of the CSR instead of two)
            const size_t nw32 = (size_t)(ncols + 31) / 32;
            std::vector<uint32_t> bits(2 * nw32, 0u);     // mask words, then their exclusive prefix popcounts
            for (int64_t c = 0; c < ncols; c++)
                if (mask[c]) bits[c >> 5] |= 1u << (c & 31);
            uint32_t run = 0;
            for (size_t w = 0; w < nw32; w++) {
                bits[nw32 + w] = run;
                run += (uint32_t)__builtin_popcount(bits[w]);
            }
            DevBuf<uint32_t> d_bits(bits.size(), st);
            SALG_CUDA(cudaMemcpyAsync(d_bits.get(), bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, st));
            d_row_kept.alloc((size_t)x->nrows + 1, st);
            // f32 randomized fits on the tensor-core path only ever touch the operator through its tile format: the
            // statistics pass then also WRITES the kept entries (at their rows' original offsets in a scratch of the
            // operator's size) and the tile builder reads them from there — no second sweep over the full matrix
            if constexpr (std::is_same<T, float>::value) {
                fused_compact = tc_enabled(ctx) && prm->svd_method == SALG_SVD_RANDOM && x->nnz > 0 &&
                                col_stats_can_fuse_compaction(x, (int64_t)run) && !getenv("SALG_NO_FUSED_COMPACT");
            }
            DevBuf<int> d_ovf(2, st);         // [0] kept-slot overflow, [1] integer-accumulator violation (stats.cu)
            SALG_CUDA(cudaMemsetAsync(d_ovf.get(), 0, 8, st));
            if (fused_compact) {
                // row r's slot in the scratch starts at ptr[r] >> shift: 2^-shift >= 3 x the kept fraction of the columns
                kept_shift = 0;
                while (kept_shift < 6 && 3.0 * (double)run * (double)(2 << kept_shift) <= (double)ncols) kept_shift++;
                const size_t n_slots = (size_t)(x->nnz >> kept_shift) + 32;
                kept_col = (uint32_t*)ctx_scratch(ctx, n_slots * (4 + sizeof(T)));
                kept_val = (T*)(kept_col + n_slots);
                SALG_CUDA(cudaMemsetAsync(d_ovf.get(), 0, 4, st));
            }
            if (host) {
                if constexpr (std::is_same<T, float>::value) {
                    if (!fused_compact) throw StreamFallback{};
                    // ---- streamed statistics + compaction: double-buffered row chunks, copy stream || compute stream
                    ProfScope ps(ctx, PROF_H2D, (double)x->nnz * 8.0);
                    cudaStream_t cs = ctx->copy_stream;
                    SALG_CUDA(cudaMemsetAsync(d_sum.get(), 0, (size_t)ncols * 8, st));
                    SALG_CUDA(cudaMemsetAsync(d_sq.get(), 0, (size_t)ncols * 8, st));
                    SALG_CUDA(cudaMemsetAsync(d_row_kept.get(), 0, (size_t)(x->nrows + 1) * 8, st));
                    DevBuf<int> d_val_flag(1, st);
                    SALG_CUDA(cudaMemsetAsync(d_val_flag.get(), 0, 4, st));
                    // chunks of ~nnz / 16 entries (>= 16 Mi), cut at row boundaries
                    const int64_t target = getenv("SALG_STREAM_CHUNK") ? std::max<int64_t>(1, atoll(getenv("SALG_STREAM_CHUNK")))
                                                                       : std::max<int64_t>(x->nnz / 16, (int64_t)1 << 24);
                    std::vector<int64_t> cut{0};
                    while (cut.back() < x->nrows) {
                        const int64_t r0 = cut.back();
                        const int64_t* lo = std::upper_bound(host->off + r0 + 1, host->off + x->nrows + 1, host->off[r0] + target);
                        int64_t r1 = (int64_t)(lo - host->off) - 1;
                        if (r1 <= r0) r1 = r0 + 1;
                        if (r1 > x->nrows) r1 = x->nrows;
                        cut.push_back(r1);
                    }
                    int64_t cap = 0;
                    for (size_t k = 0; k + 1 < cut.size(); k++) cap = std::max(cap, host->off[cut[k + 1]] - host->off[cut[k]]);
                    uint32_t* s_idx[2];
                    float* s_val[2];
                    cudaEvent_t copied[2], done[2];
                    for (int b = 0; b < 2; b++) {
                        s_idx[b] = (uint32_t*)dev_alloc(ctx, ((size_t)cap + 32) * 4);
                        s_val[b] = (float*)dev_alloc(ctx, ((size_t)cap + 32) * 4);
                        SALG_CUDA(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
                        SALG_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
                    }
                    cudaEvent_t ready;
                    SALG_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
                    SALG_CUDA(cudaEventRecord(ready, st));                  // staging buffers / accumulators exist and are zeroed
                    SALG_CUDA(cudaStreamWaitEvent(cs, ready, 0));
                    bool intsum = !getenv("SALG_STATS_NO_INT");
                    try {
                        for (size_t k = 0; k + 1 < cut.size(); k++) {
                            const int b = (int)(k & 1);
                            const int64_t r0 = cut[k], r1 = cut[k + 1], e0 = host->off[r0], n = host->off[r1] - e0;
                            if (k >= 2) SALG_CUDA(cudaStreamWaitEvent(cs, done[b], 0));
                            if (n) {
                                SALG_CUDA(cudaMemcpyAsync(s_idx[b], host->idx + e0, (size_t)n * 4, cudaMemcpyHostToDevice, cs));
                                SALG_CUDA(cudaMemcpyAsync(s_val[b], (const float*)host->val + e0, (size_t)n * 4, cudaMemcpyHostToDevice, cs));
                            }
                            SALG_CUDA(cudaEventRecord(copied[b], cs));
                            SALG_CUDA(cudaStreamWaitEvent(st, copied[b], 0));
                            if (k == 0 && intsum) intsum = col_stats_probe_int_f32(ctx, st, s_val[b], n, d_ovf.get() + 1);
                            const int64_t nr = r1 - r0;
                            int64_t want = ceil_div(nr * 32, 256), capg = (int64_t)ctx->sm_count * 16;
                            validate_rows_kernel<<<(unsigned)std::max<int64_t>(1, std::min(want, capg)), 256, 0, st>>>(
                                x->row_ptr + r0, s_idx[b] - e0, nr, (uint32_t)ncols, d_val_flag.get());
                            ctx->n_launch++;
                            col_stats_masked_chunk_f32(ctx, st, x->row_ptr + r0, s_idx[b] - e0, s_val[b] - e0, nr, ncols, (int64_t)run,
                                                       d_sum.get(), d_sq.get(), d_bits.get(), d_row_kept.get() + r0, kept_col,
                                                       (float*)kept_val, kept_shift, d_ovf.get(), intsum);
                            SALG_CUDA(cudaEventRecord(done[b], st));
                        }
                        int h_flags[2] = {0, 0}, h_valid = 0;
                        SALG_CUDA(cudaMemcpyAsync(h_flags, d_ovf.get(), 8, cudaMemcpyDeviceToHost, st));
                        SALG_CUDA(cudaMemcpyAsync(&h_valid, d_val_flag.get(), 4, cudaMemcpyDeviceToHost, st));
                        SALG_CUDA(cudaStreamSynchronize(st));
                        SALG_CUDA(cudaStreamSynchronize(cs));
                        if (h_valid & 1) throw Error(SALG_ERR_BAD_ARG, "invalid CSR: column index out of range");
                        if (h_valid & 4)
                            throw Error(SALG_ERR_BAD_ARG, "invalid CSR: column indices must be strictly increasing within a row");
                        // a row overflowing its kept-entry slot, or values that are not raw counts after all: redo from a
                        // resident upload (the accumulated sums are not reusable)
                        if (h_flags[0] || (intsum && h_flags[1])) throw StreamFallback{};
                    } catch (...) {
                        cudaStreamSynchronize(cs);
                        cudaStreamSynchronize(st);
                        for (int b = 0; b < 2; b++) { dev_free(ctx, s_idx[b]); dev_free(ctx, s_val[b]); cudaEventDestroy(copied[b]); cudaEventDestroy(done[b]); }
                        cudaEventDestroy(ready);
                        throw;
                    }
                    for (int b = 0; b < 2; b++) { dev_free(ctx, s_idx[b]); dev_free(ctx, s_val[b]); cudaEventDestroy(copied[b]); cudaEventDestroy(done[b]); }
                    cudaEventDestroy(ready);
                    if (ctx->nranks > 1) {
                        allreduce_f64(ctx, d_sum.get(), (size_t)ncols);
                        allreduce_f64(ctx, d_sq.get(), (size_t)ncols);
                    }
                } else {
                    throw StreamFallback{};
                }
            } else
            col_stats_device<T>(ctx, x, d_sum.get(), d_sq.get(), nullptr, d_bits.get(), d_row_kept.get(), (int64_t)run,
                                kept_col, kept_val, kept_shift, d_ovf.get());
            if (fused_compact && !host) {
                int h_ovf = 0;
                SALG_CUDA(cudaMemcpyAsync(&h_ovf, d_ovf.get(), 4, cudaMemcpyDeviceToHost, st));
                SALG_CUDA(cudaStreamSynchronize(st));
                if (h_ovf) fused_compact = false;      // some row keeps more than its slot holds: separate compaction pass
            }
            SALG_CUDA(cudaStreamSynchronize(st));
        } else {
            col_stats_device<T>(ctx, x, d_sum.get(), d_sq.get(), nullptr);
        }
        std::vector<double> h_sum((size_t)ncols), h_sq((size_t)ncols);
        SALG_CUDA(cudaMemcpyAsync(h_sum.data(), d_sum.get(), (size_t)ncols * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaMemcpyAsync(h_sq.data(), d_sq.get(), (size_t)ncols * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        std::vector<uint32_t> kept;
        if (mask) {
            P->mask.assign(mask, mask + ncols);
            for (int64_t c = 0; c < ncols; c++)
                if (mask[c]) kept.push_back((uint32_t)c);
        }
        const int64_t n_eff = mask ? (int64_t)kept.size() : ncols;
        SALG_REQUIRE(n_eff >= 1, SALG_ERR_BAD_ARG, "the mask selects no column");
        P->n_eff = n_eff;
        P->mean_full.assign((size_t)ncols, 0.0);
        double tv = 0.0;
        const double n_d = (double)n_total;
        for (int64_t i = 0; i < n_eff; i++) {
            int64_t c = mask ? kept[i] : i;
            double mean = h_sum[c] / n_d;
            tv += (h_sq[c] - mean * h_sum[c]) / (n_d - 1.0);
        }
        if (center)
            for (int64_t c = 0; c < ncols; c++) P->mean_full[c] = h_sum[c] / n_d;
        P->total_var = tv;
        // squared Frobenius norm of the operator the SVD sees (centred or not): every correct factorisation has
        // sum sigma_i^2 below it; checked at the end (a mis-normalised basis from a rank-deficient sketch does not)
        double raw2 = 0.0;
        for (int64_t i = 0; i < n_eff; i++) raw2 += h_sq[mask ? kept[i] : i];
        const double fro2 = center ? tv * (n_d - 1.0) : raw2;

        // ---- operator: the (column-compacted) matrix
        const salg_csr* op = x;
        DevBuf<uint32_t> d_kept;
        if (mask && fused_compact) {
            if constexpr (std::is_same<T, float>::value) {
                // shell of the compacted operator: compact row offsets + tile format, no col / val arrays
                ProfScope ps(ctx, PROF_COMPACT, 0.0);
                compact = new salg_csr();
                compact->ctx = ctx;
                compact->dtype = x->dtype;
                compact->nrows = x->nrows;
                compact->ncols = n_eff;
                compact->row_ptr = (int64_t*)dev_alloc(ctx, (size_t)(x->nrows + 1) * 8);
                exclusive_scan_i64(ctx, d_row_kept.get(), compact->row_ptr, x->nrows + 1);
                int64_t nnz_eff = 0;
                SALG_CUDA(cudaMemcpyAsync(&nnz_eff, compact->row_ptr + x->nrows, 8, cudaMemcpyDeviceToHost, st));
                SALG_CUDA(cudaStreamSynchronize(st));
                compact->nnz = nnz_eff;
            }
            if constexpr (std::is_same<T, float>::value)
                tc_attach_tiles_f32(ctx, compact, x->row_ptr, kept_shift, kept_col, kept_val);
            op = compact;
            d_kept.alloc((size_t)n_eff, st);
            SALG_CUDA(cudaMemcpyAsync(d_kept.get(), kept.data(), (size_t)n_eff * 4, cudaMemcpyHostToDevice, st));
        } else if (mask) {
            compact = csr_select_columns<T>(ctx, x, mask, d_row_kept.get());
            op = compact;
            d_kept.alloc((size_t)n_eff, st);
            SALG_CUDA(cudaMemcpyAsync(d_kept.get(), kept.data(), (size_t)n_eff * 4, cudaMemcpyHostToDevice, st));
        }
        P->d_mean = dev_alloc(ctx, (size_t)n_eff * sizeof(T));
        T* d_mu = (T*)P->d_mean;
        gather_mean_kernel<T><<<(unsigned)ceil_div(n_eff, 256), 256, 0, st>>>(d_sum.get(), mask ? d_kept.get() : nullptr,
                                                                              n_eff, 1.0 / n_d, center ? 1 : 0, d_mu);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        SALG_CUDA(cudaStreamSynchronize(st));   // kept / h_* staging no longer needed by the device

        const int rank = (int)std::min<int64_t>(prm->n_components, std::min<int64_t>(n_total, n_eff));
        P->d_V = dev_alloc(ctx, (size_t)n_eff * LP * sizeof(T));
        T* d_V = (T*)P->d_V;
        DevBuf<int> d_flag(1, st);
        SALG_CUDA(cudaMemsetAsync(d_flag.get(), 0, 4, st));
        DevBuf<double> d_sign(LP, st);
        std::vector<double> s_host;
        int d_out = rank;
        bool scores_done = false;

        if (prm->svd_method == SALG_SVD_RANDOM) {
            // l = rank + n_oversamples, clamped to the dimensions of the operator: a wider sketch is rank-deficient by
            // construction (CholeskyQR would then run on its pivot floor), and min(n, n_eff) columns already span everything
            const int l_req = rank + std::max(0, prm->n_oversamples);
            const int l = (int)std::min<int64_t>(l_req, std::min<int64_t>(n_total, n_eff));
            SALG_REQUIRE(l <= LP, SALG_ERR_UNSUPPORTED,
                         "n_components + n_oversamples must be <= 64 (device panels are 64 columns wide)");
            const int q = std::max(0, prm->n_power_iterations);
            const bool tall_norm = prm->normalizer != SALG_NORM_NONE;
            const int64_t m_loc = op->nrows;
            DevBuf<T> Om((size_t)n_eff * LP, st), Y((size_t)std::max<int64_t>(m_loc, 1) * LP, st), Z((size_t)n_eff * LP, st);
            DevBuf<double> corr(LP, st), cs(LP, st), Rb(LP * LP, st), Ur(LP * LP, st), Sr(LP, st), Vr(LP * LP, st);
            DevBuf<T> M64(LP * LP, st);
            if (omega) {
                SALG_REQUIRE(omega_rows == n_eff && (omega_cols == l || omega_cols == l_req), SALG_ERR_BAD_ARG,
                             "omega must be (kept columns) x (rank + n_oversamples), row-major");
                DevBuf<T> raw((size_t)n_eff * l, st);
                std::vector<T> lead;                                // clamped sketch: the leading l columns of omega
                const T* src = omega;
                if (omega_cols != l) {
                    lead.resize((size_t)n_eff * l);
                    for (int64_t i = 0; i < n_eff; i++)
                        std::copy(omega + i * omega_cols, omega + i * omega_cols + l, lead.begin() + i * l);
                    src = lead.data();
                }
                SALG_CUDA(cudaMemcpyAsync(raw.get(), src, (size_t)n_eff * l * sizeof(T), cudaMemcpyHostToDevice, st));
                panel_pack<T>(ctx, raw.get(), n_eff, l, Om.get());
                SALG_CUDA(cudaStreamSynchronize(st));
            } else {
                omega_normal_kernel<<<(unsigned)ceil_div(n_eff * LP, 256), 256, 0, st>>>(
                    (uint64_t)prm->random_seed, n_eff, l, sizeof(T) == 4 ? (float*)Om.get() : nullptr,
                    sizeof(T) == 8 ? (double*)Om.get() : nullptr);
                ctx->n_launch++;
                SALG_CUDA(cudaGetLastError());
            }
            // f32 operators on the tensor-core path: the tall-side normaliser of the intermediate iterations is applied
            // implicitly.  One fused pass over Y yields its Gram matrix, its column sums and its pre-split fp16 operand;
            // R^{-1} (Cholesky of the Gram) then multiplies the SMALL side: A_c^T (Y R^{-1}) = (A_c^T Y) R^{-1}.
            bool fused = false;
            DevBuf<uint8_t> Yprep;
            DevBuf<float> yscales;
            DevBuf<unsigned> yamax;
            DevBuf<double> Gy;
            DevBuf<T> RiT;
            bool zfused = false;
            DevBuf<double> zpart;
            DevBuf<unsigned> zticket;
            DevBuf<float> zM, zscales;
            DevBuf<uint8_t> zXprep;
            if constexpr (std::is_same<T, float>::value) {
                fused = tc_enabled(ctx) && tall_norm && q > 0 && !getenv("SALG_NO_FUSED_NORM");
                if (fused) {
                    Yprep.alloc(tc_yprep_bytes(ctx, op) + 16, st);
                    yscales.alloc(2, st);
                    yamax.alloc(1, st);
                    Gy.alloc(GRAM_BUF, st);
                    RiT.alloc(LP * LP, st);
                    // one launch for the replicated Cholesky / Gram / Cholesky chain of a half step, one for its application
                    zfused = tc_zside_supported(ctx, op) && !getenv("SALG_NO_ZSIDE");
                    if (zfused) {
                        zpart.alloc(zside_part_elems(), st);
                        zticket.alloc(2, st);                          // {arrival counter, sticky two-step word}
                        SALG_CUDA(cudaMemsetAsync(zticket.get(), 0, 8, st));
                        zM.alloc(LP * LP, st);
                        zscales.alloc(2, st);
                        zXprep.alloc(tc_xprep_bytes(ctx, op) + 16, st);
                    }
                }
            }
            static const bool ax_single = getenv("SALG_AX_SINGLE") != nullptr;   // (experiment: 11-bit panel in the inner iterations)
            auto product_A = [&](const T* X, bool inner = false) {      // Y = A_c X (corr holds mu^T X)
                if constexpr (std::is_same<T, float>::value) {
                    if (fused) {
                        tc_spmm_A(ctx, op, X, Y.get(), center ? corr.get() : nullptr, yamax.get(), (inner && ax_single) ? 1 : 2);
                        return;
                    }
                }
                spmm_A<T>(ctx, op, X, Y.get(), center ? corr.get() : nullptr, false);
            };
            // Y = A_c Om
            if (center) panel_colsum<T>(ctx, Om.get(), n_eff, d_mu, corr.get());
            product_A(Om.get());
            for (int it = 0; it < q; it++) {
                if (fused) {
                    if constexpr (std::is_same<T, float>::value) {
                        // every rank centres its partial with ITS OWN column sums: sum_r (A_r^T Y_r - mu cs_r^T) is the
                        // centred product, so the Gram / column sums and the partial panel share one NCCL launch
                        tc_gram_prep(ctx, op, Y.get(), yamax.get(), Yprep.get(), yscales.get(), Gy.get(),
                                     zfused && zside_mode() == 2);       // (the pure one-step small side needs no Gram of Y)
                        tc_spmm_At_prepped(ctx, op, Yprep.get(), yscales.get(), Z.get(), d_mu,
                                           center ? Gy.get() + LP * LP : nullptr);
                        allreduce_gram_and_panel(ctx, Gy.get(), GRAM_BUF, Z.get(), (size_t)n_eff * LP);
                        if (zfused) {
                            // Z <- orth(Z R1^{-1}) and everything the next A X needs of it (mu^T Z, pre-split operand): 2 launches
                            zside_solve(ctx, Z.get(), n_eff, Gy.get(), l, tc_a_scale(ctx, op), zpart.get(), zticket.get(), zM.get(),
                                        zscales.get(), corr.get(), d_flag.get());
                            tc_zside_apply(ctx, op, Z.get(), zM.get(), center ? d_mu : nullptr, zscales.get(), zXprep.get(), corr.get());
                            tc_spmm_A_prepped(ctx, op, zXprep.get(), zscales.get(), Y.get(), center ? corr.get() : nullptr, yamax.get(),
                                              (it + 1 < q && ax_single) ? 1 : 2);
                            continue;
                        }
                        chol_inv<T>(ctx, Gy.get(), l, nullptr, nullptr, RiT.get(), d_flag.get());
                        panel_mul<T>(ctx, Z.get(), n_eff, RiT.get(), Z.get());
                    }
                    cholqr2<T>(ctx, Z.get(), n_eff, l, false, nullptr, nullptr, d_flag.get(), 1);
                    if (center) panel_colsum<T>(ctx, Z.get(), n_eff, d_mu, corr.get());
                    product_A(Z.get(), it + 1 < q);
                    continue;
                }
                if (tall_norm) {
                    // intermediate iterations only need a well-conditioned basis of span(Y) (the subspace does not
                    // depend on the normaliser): one CholeskyQR pass; the final Q gets two
                    cholqr2<T>(ctx, Y.get(), m_loc, l, true, cs.get(), nullptr, d_flag.get(), 1);
                } else if (center) {
                    panel_colsum<T>(ctx, Y.get(), m_loc, nullptr, cs.get());
                    allreduce_f64(ctx, cs.get(), LP);
                }
                spmm_At<T>(ctx, op, Y.get(), Z.get(), d_mu, center ? cs.get() : nullptr);
                allreduce_panel_T<T>(ctx, Z.get(), (size_t)n_eff * LP, center ? d_mu : nullptr, cs.get(), n_eff);
                cholqr2<T>(ctx, Z.get(), n_eff, l, false, nullptr, nullptr, d_flag.get(), 1);   // intermediate: one pass
                if (center) panel_colsum<T>(ctx, Z.get(), n_eff, d_mu, corr.get());
                spmm_A<T>(ctx, op, Z.get(), Y.get(), center ? corr.get() : nullptr, false);
            }
            // Q = orth(Y); B^T = A_c^T Q
            bool final_done = false;
            if constexpr (std::is_same<T, float>::value) {
                if (fused && !getenv("SALG_NO_FUSED_FINAL")) {
                    // CholeskyQR2 with the same fused pass: Q1 = Y R1^{-1} explicitly (one read + one write of the panel),
                    // the second Gram is taken of Q1 while it is pre-split, and R2^{-1} goes to the small side:
                    // A_c^T Q = (A_c^T Q1) R2^{-1}.  Q itself is never needed (the scores are A_c V).
                    tc_gram_prep(ctx, op, Y.get(), yamax.get(), Yprep.get(), yscales.get(), Gy.get());
                    allreduce_f64(ctx, Gy.get(), GRAM_BUF);
                    chol_inv<T>(ctx, Gy.get(), l, nullptr, nullptr, RiT.get(), d_flag.get());
                    // |max| of Q1 rides on the product: columns of Q1 have unit norm only while no pivot was floored (a
                    // rank-deficient sketch leaves amplified rounding noise of any size in the dependent columns; assuming
                    // |q| <= 1.06 there overflowed the fp16 split: non-finite singular values)
                    panel_mul<T>(ctx, Y.get(), m_loc, RiT.get(), Y.get(), yamax.get());
                    // (per-rank |max| and scale, as after every A X: each rank un-scales its own partial products)
                    tc_gram_prep(ctx, op, Y.get(), yamax.get(), Yprep.get(), yscales.get(), Gy.get());
                    tc_spmm_At_prepped(ctx, op, Yprep.get(), yscales.get(), Z.get(), d_mu,
                                       center ? Gy.get() + LP * LP : nullptr);       // local column sums (see above)
                    allreduce_gram_and_panel(ctx, Gy.get(), GRAM_BUF, Z.get(), (size_t)n_eff * LP);
                    chol_inv<T>(ctx, Gy.get(), l, nullptr, nullptr, RiT.get(), d_flag.get(), true);   // second pass: drop exact dependencies
                    panel_mul<T>(ctx, Z.get(), n_eff, RiT.get(), Z.get());
                    final_done = true;
                }
            }
            if (!final_done) {
                cholqr2<T>(ctx, Y.get(), m_loc, l, true, cs.get(), nullptr, d_flag.get(), 2);
                spmm_At<T>(ctx, op, Y.get(), Z.get(), d_mu, center ? cs.get() : nullptr);
                allreduce_panel_T<T>(ctx, Z.get(), (size_t)n_eff * LP, center ? d_mu : nullptr, cs.get(), n_eff);
            }
            // B^T = Q_B R_B; R_B = U_R S V_R^T
            cholqr2<T>(ctx, Z.get(), n_eff, l, false, nullptr, Rb.get(), d_flag.get(), 2);
            // fit_transform on the TMEM-operand path: the scores are A_c V = (A_c Q_B) U_R diag(sign), so the big product
            // A_c Q_B does not have to wait for the one-CTA Jacobi SVD — it runs beside it (the SVD on the side stream, the
            // persistent product kernel on one SM less) and only a rows x 64 by 64 x 64 product is left afterwards.
            bool tail_overlap = false;
            if constexpr (std::is_same<T, float>::value)
                tail_overlap = prm->keep_scores && fused && zfused && op->nrows > 0 && !getenv("SALG_NO_TAIL_OVERLAP");
            if (tail_overlap) {
                if (!ctx->ev_fork) {
                    SALG_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
                    SALG_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
                }
                SALG_CUDA(cudaEventRecord(ctx->ev_fork, st));
                SALG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
                ctx->stream = ctx->copy_stream;
                try {
                    jacobi_svd64(ctx, Rb.get(), l, Ur.get(), Sr.get(), Vr.get(), d_flag.get());
                } catch (...) {
                    ctx->stream = st;
                    throw;
                }
                ctx->stream = st;
                SALG_CUDA(cudaEventRecord(ctx->ev_join, ctx->copy_stream));
                try {
                    P->d_scores = dev_alloc(ctx, (size_t)op->nrows * LP * sizeof(T));
                    if constexpr (std::is_same<T, float>::value) {
                        if (center) panel_colsum<T>(ctx, Z.get(), n_eff, d_mu, corr.get());
                        ctx->sm_reserve = 1;
                        tc_spmm_A(ctx, op, Z.get(), (float*)P->d_scores, center ? corr.get() : nullptr);
                        ctx->sm_reserve = 0;
                    }
                    SALG_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
                } catch (...) {
                    ctx->sm_reserve = 0;
                    cudaStreamSynchronize(ctx->copy_stream);     // the side stream still reads this function's buffers
                    throw;
                }
            } else {
                jacobi_svd64(ctx, Rb.get(), l, Ur.get(), Sr.get(), Vr.get(), d_flag.get());
            }
            cast_mat64<T>(ctx, Ur.get(), M64.get(), nullptr);
            panel_mul<T>(ctx, Z.get(), n_eff, M64.get(), d_V);                 // V = Q_B U_R
            flip_find<T>(ctx, d_V, n_eff, d_sign.get());
            panel_colscale<T>(ctx, d_V, n_eff, d_sign.get());
            if (tail_overlap) {
                cast_mat64<T>(ctx, Ur.get(), M64.get(), d_sign.get());         // U_R diag(sign)
                panel_mul<T>(ctx, (const T*)P->d_scores, op->nrows, M64.get(), (T*)P->d_scores);
                scores_done = true;
            }
            s_host.resize(LP);
            SALG_CUDA(cudaMemcpyAsync(s_host.data(), Sr.get(), LP * 8, cudaMemcpyDeviceToHost, st));
            SALG_CUDA(cudaStreamSynchronize(st));
            s_host.resize(rank);
        } else {
            SALG_REQUIRE(rank <= LP, SALG_ERR_UNSUPPORTED, "n_components must be <= 64 (device panels are 64 columns wide)");
            int steps = 0;
            double tol = sizeof(T) == 4 ? 2e-6 : 1e-10;
            int got = lanczos_svd<T>(ctx, op, rank, prm->lanczos_max_steps, (uint64_t)prm->random_seed, tol, d_V, s_host,
                                     &steps);
            d_out = std::abs(got);
            if (got <= 0) P->numeric_flag |= 4;   // not all requested triplets met the acceptance bound
            SALG_REQUIRE(d_out >= 1, SALG_ERR_NUMERIC, "SVD computation failed: Lanczos found no singular triplet");
            if (prm->verbose) fprintf(stderr, "[salg] lanczos: %d steps, %d triplets\n", steps, d_out);
            flip_find<T>(ctx, d_V, n_eff, d_sign.get());
            panel_colscale<T>(ctx, d_V, n_eff, d_sign.get());
            SALG_CUDA(cudaStreamSynchronize(st));
        }
        if (prm->keep_scores && !scores_done) {
            // fit_transform = fit, then transform of the same rows (pca/sparse/mod.rs:355-358):
            // (X - 1 mu^T) V on the kept columns, computed while the compacted operator is still resident.
            // (U S from the factorisation is only the projection of this onto range(Q).)
            P->d_scores = dev_alloc(ctx, (size_t)std::max<int64_t>(op->nrows, 1) * LP * sizeof(T));
            DevBuf<double> corr(LP, st);
            if (center) panel_colsum<T>(ctx, d_V, n_eff, d_mu, corr.get());
            spmm_A<T>(ctx, op, d_V, (T*)P->d_scores, center ? corr.get() : nullptr, false);
            SALG_CUDA(cudaStreamSynchronize(st));
        }
        int h_flag = 0;
        SALG_CUDA(cudaMemcpy(&h_flag, d_flag.get(), 4, cudaMemcpyDeviceToHost));
        P->numeric_flag |= h_flag;
        P->d = d_out;
        P->singular_values.assign(s_host.begin(), s_host.begin() + d_out);
        P->explained_variance.resize(d_out);
        for (int i = 0; i < d_out; i++) {
            double s = P->singular_values[i];
            // the reference's two messages: pca/sparse/mod.rs:144 (Lanczos) and :180 (randomized)
            SALG_REQUIRE(std::isfinite(s), SALG_ERR_NUMERIC,
                         prm->svd_method == SALG_SVD_RANDOM ? "Randomized SVD computation failed: non-finite singular value"
                                                            : "SVD computation failed: non-finite singular value");
            P->explained_variance[i] = s * s / (n_d - 1.0);   // pca/sparse/mod.rs:210-216
        }
        if (prm->svd_method == SALG_SVD_RANDOM) {
            double s2 = 0.0;
            for (double sv : P->singular_values) s2 += sv * sv;
            // (f32 statistics and products: 1e-3 of slack; the failure this guards against overshoots by orders of magnitude)
            SALG_REQUIRE(s2 <= fro2 * (1.0 + 1e-3) + 1e-9 * raw2 + 1e-300, SALG_ERR_NUMERIC,
                         "Randomized SVD computation failed: the projected factor exceeds the operator's Frobenius norm "
                         "(rank-deficient sketch: fewer independent directions than n_components + n_oversamples)");
        }
        if (!center) {   // pca/sparse/mod.rs:218-223: without centring the total is the sum over components
            double t = 0.0;
            for (double e : P->explained_variance) t += e;
            P->total_var = t;
        }
        if (prm->verbose)
            fprintf(stderr, "[salg] fit: n=%lld n_eff=%lld d=%d total_var=%.6g flag=%d\n", (long long)n_total,
                    (long long)n_eff, d_out, P->total_var, P->numeric_flag);
        if (compact) {
            csr_destroy(compact);
            compact = nullptr;
        }
    } catch (...) {
        cudaStreamSynchronize(st);
        if (compact) csr_destroy(compact);
        pca_destroy(P);
        throw;
    }
    return P;
}

// scores (device, nrows x 64 panel) for the rows of x
template <typename T>
static void pca_transform_device(salg_ctx* ctx, const salg_pca* P, const salg_csr* x, int mode, T* d_scores) {
    cudaStream_t st = ctx->stream;
    const int64_t n_eff = P->n_eff;
    salg_csr* compact = nullptr;
    const salg_csr* op = x;
    try {
        if (P->masked) {
            SALG_REQUIRE((int64_t)P->mask.size() == x->ncols, SALG_ERR_MASK_LEN,
                         "The mask vector length and the number of features (columns) have to be the same!");
            compact = csr_select_columns<T>(ctx, x, P->mask.data());
            op = compact;
        } else {
            SALG_REQUIRE(x->ncols == P->ncols, SALG_ERR_BAD_ARG, "matrix has a different number of features than the fit");
        }
        const T* V = (const T*)P->d_V;
        const T* mu = (const T*)P->d_mean;
        DevBuf<double> corr(LP, st);
        if (mode == SALG_TRANSFORM_EXACT || !P->center) {
            if (P->center) panel_colsum<T>(ctx, V, n_eff, mu, corr.get());
            spmm_A<T>(ctx, op, V, d_scores, P->center ? corr.get() : nullptr, false);
        } else if (P->masked) {
            // pca/sparse_masked/mod.rs:488-529: only STORED kept entries contribute (x - mu_c) V[k, c]
            DevBuf<T> W((size_t)n_eff * LP, st), S2((size_t)std::max<int64_t>(op->nrows, 1) * LP, st);
            rowscale_panel_kernel<T><<<(unsigned)ceil_div(n_eff * LP, 256), 256, 0, st>>>(V, n_eff, mu, nullptr, W.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
            spmm_A<T>(ctx, op, V, d_scores, nullptr, false);
            spmm_A<T>(ctx, op, W.get(), S2.get(), nullptr, true);
            int64_t n = op->nrows * LP;
            if (n) { panel_sub_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 4096), 256, 0, st>>>(d_scores, S2.get(), n); ctx->n_launch++; }
            SALG_CUDA(cudaGetLastError());
        } else {
            // pca/sparse/mod.rs:268-282: the loop walks x.col_indices() of the WHOLE matrix, so column c is
            // visited cnt_c = nnz(column c) times
            DevBuf<double> d_sum((size_t)n_eff, st), d_cnt((size_t)n_eff, st);
            col_stats_device<T>(ctx, x, d_sum.get(), nullptr, d_cnt.get());
            DevBuf<T> W((size_t)n_eff * LP, st);
            rowscale_panel_kernel<T><<<(unsigned)ceil_div(n_eff * LP, 256), 256, 0, st>>>(V, n_eff, nullptr, d_cnt.get(), W.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
            panel_colsum<T>(ctx, W.get(), n_eff, mu, corr.get());
            spmm_A<T>(ctx, op, W.get(), d_scores, corr.get(), false);
        }
        SALG_CUDA(cudaStreamSynchronize(st));
        if (compact) csr_destroy(compact);
    } catch (...) {
        cudaStreamSynchronize(st);
        if (compact) csr_destroy(compact);
        throw;
    }
}

template <typename T>
static void pca_transform_api(salg_ctx* ctx, const salg_pca* P, const salg_csr* x, int mode, T* scores) {
    SALG_REQUIRE(ctx && x, SALG_ERR_BAD_ARG, "ctx/x is NULL");
    SALG_REQUIRE(P && P->d_V, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
    SALG_REQUIRE(P->dtype == dtype_of<T>::value && x->dtype == P->dtype, SALG_ERR_BAD_ARG, "value type mismatch");
    SALG_REQUIRE(mode == SALG_TRANSFORM_EXACT || mode == SALG_TRANSFORM_REFERENCE_COMPAT, SALG_ERR_BAD_ARG, "bad mode");
    SALG_REQUIRE(scores || x->nrows == 0, SALG_ERR_BAD_ARG, "scores is NULL");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (x->nrows == 0) return;
    DevBuf<T> S((size_t)x->nrows * LP, st), out((size_t)x->nrows * P->d, st);
    pca_transform_device<T>(ctx, P, x, mode, S.get());
    panel_unpack<T>(ctx, S.get(), x->nrows, (int)P->d, out.get());
    SALG_CUDA(cudaMemcpyAsync(scores, out.get(), (size_t)x->nrows * P->d * sizeof(T), cudaMemcpyDeviceToHost, st));
    SALG_CUDA(cudaStreamSynchronize(st));
}

template <typename T>
static void pca_fit_scores_api(salg_ctx* ctx, const salg_pca* P, T* scores) {
    SALG_REQUIRE(ctx, SALG_ERR_BAD_ARG, "ctx is NULL");
    SALG_REQUIRE(P && P->d_V, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
    SALG_REQUIRE(P->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "value type mismatch");
    SALG_REQUIRE(P->d_scores, SALG_ERR_BAD_ARG, "the model was fitted without keep_scores");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int64_t m = P->n_fit_rows_local;
    if (m == 0) return;
    SALG_REQUIRE(scores, SALG_ERR_BAD_ARG, "scores is NULL");
    DevBuf<T> out((size_t)m * P->d, st);
    panel_unpack<T>(ctx, (const T*)P->d_scores, m, (int)P->d, out.get());
    SALG_CUDA(cudaMemcpyAsync(scores, out.get(), (size_t)m * P->d * sizeof(T), cudaMemcpyDeviceToHost, st));
    SALG_CUDA(cudaStreamSynchronize(st));
}

template <typename T>
static void pca_components_api(const salg_pca* P, T* out) {
    SALG_REQUIRE(P && P->d_V, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
    SALG_REQUIRE(P->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "value type mismatch");
    SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
    salg_ctx* ctx = P->ctx;
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<T> t((size_t)P->d * P->n_eff, st);
    panel_to_rowmajor_t<T>(ctx, (const T*)P->d_V, P->n_eff, (int)P->d, t.get());
    SALG_CUDA(cudaMemcpyAsync(out, t.get(), (size_t)P->d * P->n_eff * sizeof(T), cudaMemcpyDeviceToHost, st));
    SALG_CUDA(cudaStreamSynchronize(st));
}

// ---- operator-level entry points -----------------------------------------------------------------------------------
template <typename T>
static void op_spmm_api(salg_ctx* ctx, const salg_csr* c, int transposed, const T* dense, int64_t k, const T* mu, T* out) {
    SALG_REQUIRE(ctx && c && dense && out, SALG_ERR_BAD_ARG, "NULL argument");
    SALG_REQUIRE(c->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "csr value type does not match the entry point");
    SALG_REQUIRE(k >= 1 && k <= LP, SALG_ERR_UNSUPPORTED, "k must be in 1..64");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int64_t n_in = transposed ? c->nrows : c->ncols, n_out = transposed ? c->ncols : c->nrows;
    DevBuf<T> raw((size_t)std::max<int64_t>(n_in * k, 1), st), X((size_t)std::max<int64_t>(n_in, 1) * LP, st),
        O((size_t)std::max<int64_t>(n_out, 1) * LP, st), oraw((size_t)std::max<int64_t>(n_out * k, 1), st), d_mu;
    DevBuf<double> corr(LP, st);
    if (n_in) SALG_CUDA(cudaMemcpyAsync(raw.get(), dense, (size_t)n_in * k * sizeof(T), cudaMemcpyHostToDevice, st));
    panel_pack<T>(ctx, raw.get(), n_in, (int)k, X.get());
    if (mu) {
        d_mu.alloc((size_t)std::max<int64_t>(c->ncols, 1), st);
        if (c->ncols) SALG_CUDA(cudaMemcpyAsync(d_mu.get(), mu, (size_t)c->ncols * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    if (!transposed) {
        if (mu) panel_colsum<T>(ctx, X.get(), n_in, d_mu.get(), corr.get());
        spmm_A<T>(ctx, c, X.get(), O.get(), mu ? corr.get() : nullptr, false);
    } else {
        if (mu) {
            panel_colsum<T>(ctx, X.get(), n_in, nullptr, corr.get());
            allreduce_f64(ctx, corr.get(), LP);
        }
        spmm_At<T>(ctx, c, X.get(), O.get(), mu ? d_mu.get() : nullptr, mu ? corr.get() : nullptr);
        allreduce_panel_T<T>(ctx, O.get(), (size_t)n_out * LP, mu ? d_mu.get() : nullptr, corr.get(), n_out);
    }
    panel_unpack<T>(ctx, O.get(), n_out, (int)k, oraw.get());
    if (n_out) SALG_CUDA(cudaMemcpyAsync(out, oraw.get(), (size_t)n_out * k * sizeof(T), cudaMemcpyDeviceToHost, st));
    SALG_CUDA(cudaStreamSynchronize(st));
}

template <typename T>
static void op_cholqr2_api(salg_ctx* ctx, const T* panel, int64_t m, int64_t k, T* q, double* r) {
    SALG_REQUIRE(ctx && panel && q, SALG_ERR_BAD_ARG, "NULL argument");
    SALG_REQUIRE(k >= 1 && k <= LP && m >= 1, SALG_ERR_UNSUPPORTED, "k must be in 1..64, m >= 1");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<T> raw((size_t)m * k, st), Y((size_t)m * LP, st);
    DevBuf<double> R(LP * LP, st);
    DevBuf<int> flag(1, st);
    SALG_CUDA(cudaMemsetAsync(flag.get(), 0, 4, st));
    SALG_CUDA(cudaMemcpyAsync(raw.get(), panel, (size_t)m * k * sizeof(T), cudaMemcpyHostToDevice, st));
    panel_pack<T>(ctx, raw.get(), m, (int)k, Y.get());
    cholqr2<T>(ctx, Y.get(), m, (int)k, true, nullptr, R.get(), flag.get(), 2);
    panel_unpack<T>(ctx, Y.get(), m, (int)k, raw.get());
    SALG_CUDA(cudaMemcpyAsync(q, raw.get(), (size_t)m * k * sizeof(T), cudaMemcpyDeviceToHost, st));
    if (r) {
        std::vector<double> h(LP * LP);
        SALG_CUDA(cudaMemcpyAsync(h.data(), R.get(), LP * LP * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < k; i++)
            for (int64_t j = 0; j < k; j++) r[i * k + j] = h[i * LP + j];
    }
    SALG_CUDA(cudaStreamSynchronize(st));
}

__global__ void mul64_kernel(const double* a, const double* b, double* o) {
    if (threadIdx.x < LP) o[threadIdx.x] = a[threadIdx.x] * b[threadIdx.x];
}
void mul64_kernel_launch(salg_ctx* ctx, const double* a, const double* b, double* o) {
    mul64_kernel<<<1, 64, 0, ctx->stream>>>(a, b, o);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

// Sum the per-rank partial panels of A^T Y.  Every rank subtracted mu*cs (cs = GLOBAL 1^T Y) from its
// partial, so after the sum the correction is present nranks times: add back (nranks-1)*mu*cs.
template <typename T>
__global__ void fix_corr_kernel(T* __restrict__ Z, int64_t n_eff, const T* __restrict__ mu,
                                const double* __restrict__ cs, double factor) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_eff * LP) return;
    Z[i] = (T)((double)Z[i] + factor * (double)mu[i >> 6] * cs[i & 63]);
}
template <typename T>
void allreduce_panel_T(salg_ctx* ctx, T* Z, size_t n, const T* mu, const double* cs, int64_t n_eff) {
    if (ctx->nranks <= 1) return;
    allreduce_T<T>(ctx, Z, n);
    if (mu) {
        fix_corr_kernel<T><<<(unsigned)ceil_div(n_eff * LP, 256), 256, 0, ctx->stream>>>(Z, n_eff, mu, cs,
                                                                                        (double)(ctx->nranks - 1));
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
}
template void allreduce_panel_T<float>(salg_ctx*, float*, size_t, const float*, const double*, int64_t);
template void allreduce_panel_T<double>(salg_ctx*, double*, size_t, const double*, const double*, int64_t);

}  // namespace salg

using namespace salg;

extern "C" {

int salg_pca_params_default(salg_pca_params* p) {
    return guarded([&] {
        SALG_REQUIRE(p, SALG_ERR_BAD_ARG, "params is NULL");
        memset(p, 0, sizeof(*p));
        p->n_components = 50;                 // pca/sparse/mod.rs:391
        p->svd_method = SALG_SVD_LANCZOS;     // pca/mod.rs:64-68
        p->n_oversamples = 10;
        p->n_power_iterations = 7;
        p->normalizer = SALG_NORM_QR;
        p->center = 1;                        // pca/sparse/mod.rs:398
        p->verbose = 0;
        p->random_seed = 42;                  // pca/sparse/mod.rs:397
        p->alpha = 1.0;                       // pca/sparse/mod.rs:392
        p->tolerance = 1e-6;                  // pca/sparse/mod.rs:393
        p->lanczos_max_steps = 0;
        p->keep_scores = 0;
    });
}

int salg_pca_fit_f32(salg_ctx* ctx, const salg_csr* x, const salg_pca_params* prm, const uint8_t* mask,
                     int64_t mask_len, const float* omega, int64_t orows, int64_t ocols, salg_pca** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = pca_fit<float>(ctx, x, prm, mask, mask_len, omega, orows, ocols);
    });
}
int salg_pca_fit_host_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* row_offsets,
                          const int32_t* col_indices, const float* values, const salg_pca_params* prm, const uint8_t* mask,
                          int64_t mask_len, const float* omega, int64_t orows, int64_t ocols, salg_pca** out) {
    return guarded([&] {
        SALG_REQUIRE(ctx && out && prm && row_offsets, SALG_ERR_BAD_ARG, "NULL argument");
        SALG_REQUIRE(nnz == 0 || (col_indices && values), SALG_ERR_BAD_ARG, "col_indices/values is NULL");
        if (mask)
            SALG_REQUIRE(mask_len == ncols, SALG_ERR_MASK_LEN,
                         "The mask vector length and the number of features (columns) have to be the same!");
        SALG_CUDA(cudaSetDevice(ctx->device));
        const bool can_stream = mask && prm->svd_method == SALG_SVD_RANDOM && tc_enabled(ctx) && nnz > 0 && nrows > 0 &&
                                !getenv("SALG_NO_STREAM_FIT");
        if (can_stream) {
            salg_csr* shell = csr_shell_from_host_offsets(ctx, SALG_F32, nrows, ncols, nnz, row_offsets);
            HostCsrSrc src{row_offsets, col_indices, values};
            try {
                *out = pca_fit<float>(ctx, shell, prm, mask, mask_len, omega, orows, ocols, &src);
                csr_destroy(shell);
                return;
            } catch (const StreamFallback&) {
                csr_destroy(shell);          // redo below from a resident upload
            } catch (...) {
                csr_destroy(shell);
                throw;
            }
        }
        salg_csr* c = csr_upload_i32_f32(ctx, nrows, ncols, nnz, row_offsets, col_indices, values);
        try {
            *out = pca_fit<float>(ctx, c, prm, mask, mask_len, omega, orows, ocols);
        } catch (...) {
            csr_destroy(c);
            throw;
        }
        csr_destroy(c);
    });
}
int salg_pca_fit_f64(salg_ctx* ctx, const salg_csr* x, const salg_pca_params* prm, const uint8_t* mask,
                     int64_t mask_len, const double* omega, int64_t orows, int64_t ocols, salg_pca** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = pca_fit<double>(ctx, x, prm, mask, mask_len, omega, orows, ocols);
    });
}

int salg_pca_free(salg_pca* p) {
    return guarded([&] { pca_destroy(p); });
}

int salg_pca_dims(const salg_pca* p, int64_t* d, int64_t* n_eff, int64_t* ncols, int* dtype) {
    return guarded([&] {
        SALG_REQUIRE(p, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
        if (d) *d = p->d;
        if (n_eff) *n_eff = p->n_eff;
        if (ncols) *ncols = p->ncols;
        if (dtype) *dtype = p->dtype;
    });
}

int salg_pca_components_f32(const salg_pca* p, float* out) {
    return guarded([&] { pca_components_api<float>(p, out); });
}
int salg_pca_components_f64(const salg_pca* p, double* out) {
    return guarded([&] { pca_components_api<double>(p, out); });
}
int salg_pca_singular_values_f64(const salg_pca* p, double* out) {
    return guarded([&] {
        SALG_REQUIRE(p, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        std::copy(p->singular_values.begin(), p->singular_values.end(), out);
    });
}
int salg_pca_explained_variance_f64(const salg_pca* p, double* out) {
    return guarded([&] {
        SALG_REQUIRE(p, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        std::copy(p->explained_variance.begin(), p->explained_variance.end(), out);
    });
}
int salg_pca_mean_f64(const salg_pca* p, double* out) {
    return guarded([&] {
        SALG_REQUIRE(p, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        std::copy(p->mean_full.begin(), p->mean_full.end(), out);
    });
}
int salg_pca_total_var(const salg_pca* p, double* out) {
    return guarded([&] {
        SALG_REQUIRE(p, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = p->total_var;
    });
}
int salg_pca_n_samples(const salg_pca* p, int64_t* out) {
    return guarded([&] {
        SALG_REQUIRE(p && out, SALG_ERR_BAD_ARG, "NULL argument");
        *out = p->n_samples;
    });
}
int salg_pca_numeric_flags(const salg_pca* p, int* out) {
    return guarded([&] {
        SALG_REQUIRE(p && out, SALG_ERR_BAD_ARG, "NULL argument");
        *out = p->numeric_flag;
    });
}

int salg_pca_transform_f32(salg_ctx* ctx, const salg_pca* p, const salg_csr* x, int mode, float* scores) {
    return guarded([&] { pca_transform_api<float>(ctx, p, x, mode, scores); });
}
int salg_pca_transform_f64(salg_ctx* ctx, const salg_pca* p, const salg_csr* x, int mode, double* scores) {
    return guarded([&] { pca_transform_api<double>(ctx, p, x, mode, scores); });
}
int salg_pca_fit_scores_f32(salg_ctx* ctx, const salg_pca* p, float* scores) {
    return guarded([&] { pca_fit_scores_api<float>(ctx, p, scores); });
}
int salg_pca_fit_scores_f64(salg_ctx* ctx, const salg_pca* p, double* scores) {
    return guarded([&] { pca_fit_scores_api<double>(ctx, p, scores); });
}

int salg_pca_transform_device(salg_ctx* ctx, const salg_pca* cp, const salg_csr* x, int mode) {
    return guarded([&] {
        salg_pca* p = const_cast<salg_pca*>(cp);
        SALG_REQUIRE(ctx && x, SALG_ERR_BAD_ARG, "ctx/x is NULL");
        SALG_REQUIRE(p && p->d_V, SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
        SALG_REQUIRE(x->dtype == p->dtype, SALG_ERR_BAD_ARG, "value type mismatch");
        SALG_CUDA(cudaSetDevice(ctx->device));
        size_t es = p->dtype == SALG_F64 ? 8 : 4;
        if (p->tscores_rows < x->nrows || !p->d_tscores) {
            SALG_CUDA(cudaStreamSynchronize(ctx->stream));
            dev_free(p->ctx, p->d_tscores);
            p->d_tscores = nullptr;
            p->d_tscores = dev_alloc(ctx, (size_t)std::max<int64_t>(x->nrows, 1) * LP * es);
            p->tscores_rows = x->nrows;
        }
        if (x->nrows == 0) return;
        if (p->dtype == SALG_F64) pca_transform_device<double>(ctx, p, x, mode, (double*)p->d_tscores);
        else pca_transform_device<float>(ctx, p, x, mode, (float*)p->d_tscores);
    });
}

int salg_op_spmm_f32(salg_ctx* ctx, const salg_csr* c, int transposed, const float* dense, int64_t k,
                     const float* mu, float* out) {
    return guarded([&] { op_spmm_api<float>(ctx, c, transposed, dense, k, mu, out); });
}
int salg_op_spmm_f64(salg_ctx* ctx, const salg_csr* c, int transposed, const double* dense, int64_t k,
                     const double* mu, double* out) {
    return guarded([&] { op_spmm_api<double>(ctx, c, transposed, dense, k, mu, out); });
}
int salg_op_cholqr2_f32(salg_ctx* ctx, const float* panel, int64_t m, int64_t k, float* q, double* r) {
    return guarded([&] { op_cholqr2_api<float>(ctx, panel, m, k, q, r); });
}
int salg_op_cholqr2_f64(salg_ctx* ctx, const double* panel, int64_t m, int64_t k, double* q, double* r) {
    return guarded([&] { op_cholqr2_api<double>(ctx, panel, m, k, q, r); });
}

int salg_op_small_svd(salg_ctx* ctx, const double* a, int64_t k, double* u, double* s, double* vt) {
    return guarded([&] {
        SALG_REQUIRE(ctx && a && s, SALG_ERR_BAD_ARG, "NULL argument");
        SALG_REQUIRE(k >= 1 && k <= LP, SALG_ERR_UNSUPPORTED, "k must be in 1..64");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        std::vector<double> h(LP * LP, 0.0);
        for (int64_t i = 0; i < k; i++)
            for (int64_t j = 0; j < k; j++) h[i * LP + j] = a[i * k + j];
        DevBuf<double> A(LP * LP, st), U(LP * LP, st), S(LP, st), V(LP * LP, st);
        DevBuf<int> flag(1, st);
        SALG_CUDA(cudaMemsetAsync(flag.get(), 0, 4, st));
        SALG_CUDA(cudaMemcpyAsync(A.get(), h.data(), LP * LP * 8, cudaMemcpyHostToDevice, st));
        jacobi_svd64(ctx, A.get(), (int)k, U.get(), S.get(), V.get(), flag.get());
        std::vector<double> hu(LP * LP), hv(LP * LP), hs(LP);
        SALG_CUDA(cudaMemcpyAsync(hu.data(), U.get(), LP * LP * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaMemcpyAsync(hv.data(), V.get(), LP * LP * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaMemcpyAsync(hs.data(), S.get(), LP * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < k; i++) {
            s[i] = hs[i];
            for (int64_t j = 0; j < k; j++) {
                if (u) u[i * k + j] = hu[i * LP + j];
                if (vt) vt[i * k + j] = hv[j * LP + i];   // vt = V^T
            }
        }
    });
}

int salg_op_tall_gram_f32(salg_ctx* ctx, const float* panel, int64_t m, int64_t k, int64_t device_rows, double* gram,
                          double* colsum, int iters, double* avg_ms) {
    return guarded([&] {
        SALG_REQUIRE(ctx && gram && colsum, SALG_ERR_BAD_ARG, "NULL argument");
        SALG_REQUIRE(k >= 1 && k <= LP, SALG_ERR_UNSUPPORTED, "panel width must be 1..64");
        SALG_REQUIRE(tc_enabled(ctx), SALG_ERR_UNSUPPORTED, "tensor-core path disabled (SALG_SPMM_IMPL)");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        // panel != NULL: host (m x k row-major); panel == NULL: a device-generated normal panel of device_rows rows (probe)
        const int64_t rows = panel ? m : device_rows;
        SALG_REQUIRE(rows >= 0, SALG_ERR_BAD_ARG, "negative row count");
        DevBuf<float> P((size_t)std::max<int64_t>(rows, 1) * LP, st);
        if (panel) {
            DevBuf<float> raw((size_t)std::max<int64_t>(rows * k, 1), st);
            if (rows) SALG_CUDA(cudaMemcpyAsync(raw.get(), panel, (size_t)rows * k * 4, cudaMemcpyHostToDevice, st));
            panel_pack<float>(ctx, raw.get(), rows, (int)k, P.get());
            SALG_CUDA(cudaStreamSynchronize(st));
        } else if (rows) {
            omega_normal_kernel<<<(unsigned)ceil_div(rows * LP, 256), 256, 0, st>>>(11, rows, (int)k, P.get(), nullptr);
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        DevBuf<double> G(GRAM_BUF, st);
        tc_gram_probe(ctx, P.get(), rows, G.get(), nullptr, iters, avg_ms);
        std::vector<double> h(GRAM_BUF);
        SALG_CUDA(cudaMemcpyAsync(h.data(), G.get(), GRAM_BUF * 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < k; i++)
            for (int64_t j = 0; j < k; j++) gram[i * k + j] = h[i * LP + j];
        for (int64_t j = 0; j < k; j++) colsum[j] = h[LP * LP + j];
    });
}

int salg_op_spmm_bench(salg_ctx* ctx, const salg_csr* c, int transposed, int64_t k, int iters, double* avg_ms) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c && avg_ms, SALG_ERR_BAD_ARG, "NULL argument");
        SALG_REQUIRE(iters >= 1, SALG_ERR_BAD_ARG, "iters must be >= 1");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        int64_t n_in = transposed ? c->nrows : c->ncols, n_out = transposed ? c->ncols : c->nrows;
        size_t es = c->dtype == SALG_F64 ? 8 : 4;
        DevBuf<uint8_t> X((size_t)std::max<int64_t>(n_in, 1) * LP * es, st), O((size_t)std::max<int64_t>(n_out, 1) * LP * es, st),
            mu((size_t)std::max<int64_t>(c->ncols, 1) * es, st);
        DevBuf<double> corr(LP, st);
        SALG_CUDA(cudaMemsetAsync(mu.get(), 0, (size_t)std::max<int64_t>(c->ncols, 1) * es, st));
        SALG_CUDA(cudaMemsetAsync(corr.get(), 0, LP * 8, st));
        omega_normal_kernel<<<(unsigned)ceil_div(std::max<int64_t>(n_in, 1) * LP, 256), 256, 0, st>>>(
            7, std::max<int64_t>(n_in, 1), (int)k, es == 4 ? (float*)X.get() : nullptr, es == 8 ? (double*)X.get() : nullptr);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        auto run = [&]() {
            if (c->dtype == SALG_F64) {
                if (transposed) spmm_At<double>(ctx, c, (double*)X.get(), (double*)O.get(), (double*)mu.get(), corr.get());
                else spmm_A<double>(ctx, c, (double*)X.get(), (double*)O.get(), corr.get(), false);
            } else {
                if (transposed) spmm_At<float>(ctx, c, (float*)X.get(), (float*)O.get(), (float*)mu.get(), corr.get());
                else spmm_A<float>(ctx, c, (float*)X.get(), (float*)O.get(), corr.get(), false);
            }
        };
        for (int i = 0; i < 3; i++) run();
        cudaEvent_t e0, e1;
        SALG_CUDA(cudaEventCreate(&e0));
        SALG_CUDA(cudaEventCreate(&e1));
        SALG_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < iters; i++) run();
        SALG_CUDA(cudaEventRecord(e1, st));
        SALG_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SALG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *avg_ms = (double)ms / iters;
    });
}

}  // extern "C"
