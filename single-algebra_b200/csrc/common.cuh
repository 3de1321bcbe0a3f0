// common.cuh — internal declarations shared by the translation units of libsalg_b200.so.
// Not part of the public ABI (that is include/salg.h).
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/salg.h"

namespace salg {

constexpr int LP = 64;  // padded panel width: every dense panel is (rows x 64), row-major, ld = 64
constexpr int SPMM_CHUNK = 256;          // stored entries per warp work item in the SpMM kernels
constexpr int GRAM_BUF = LP * LP + LP;   // Gram (64x64) followed by the panel's column sums (64), f64

// ---- errors ------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

#define SALG_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            throw ::salg::Error(_e == cudaErrorMemoryAllocation ? SALG_ERR_OOM : SALG_ERR_CUDA, \
                                std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                    __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")"); \
    } while (0)

#define SALG_NCCL(expr)                                                                          \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess)                                                                   \
            throw ::salg::Error(SALG_ERR_NCCL, std::string("NCCL error: ") +                     \
                                                   ncclGetErrorString(_r) + " at " + __FILE__ + \
                                                   ":" + std::to_string(__LINE__));              \
    } while (0)

#define SALG_REQUIRE(cond, code, msg)                      \
    do {                                                   \
        if (!(cond)) throw ::salg::Error((code), (msg));   \
    } while (0)

// Wraps the body of every extern "C" entry point: no exception crosses the ABI.
template <typename F>
int guarded(F&& f) {
    try {
        f();
        return SALG_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return SALG_ERR_OOM;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return SALG_ERR_BAD_ARG;
    } catch (...) {
        set_last_error("unknown error");
        return SALG_ERR_BAD_ARG;
    }
}

// ---- profiling classes ---------------------------------------------------------------------------
enum ProfClass {
    PROF_SPMM = 0,     // Y = A X   (gather over the CSR)
    PROF_SPMMT,        // Z = A^T Y (gather over the transposed copy)
    PROF_GRAM,
    PROF_CHOL,
    PROF_PANELMUL,
    PROF_JACOBI,
    PROF_STATS,
    PROF_TRANSPOSE,
    PROF_COMPACT,
    PROF_ELEMENTWISE,
    PROF_ALLREDUCE,
    PROF_H2D,
    PROF_SPMV,
    PROF_OTHER,
    PROF_NCLS
};

struct ProfRecord {
    int cls;
    cudaEvent_t e0, e1;
    double bytes;
};

}  // namespace salg

// ---- opaque handle definitions ---------------------------------------------------------------------
struct salg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    int spmm_impl = 0;             // 0 = tcgen05 tile-densified products for f32 operators, 1 = CUDA-core chunk kernels
    int64_t n_launch = 0;          // kernels of this library launched on `stream` (bench.py: gpu_launches)
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    // profiling
    bool prof_on = false;
    bool prof_products_only = false;
    std::vector<salg::ProfRecord> prof_pending;
    std::vector<cudaEvent_t> event_pool;
    double prof_ms[salg::PROF_NCLS] = {0};
    int64_t prof_launches[salg::PROF_NCLS] = {0};
    double prof_bytes[salg::PROF_NCLS] = {0};
    // pinned staging ring for uploads
    static constexpr int N_STAGE = 3;
    void* stage[N_STAGE] = {nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[N_STAGE] = {nullptr, nullptr, nullptr};
    size_t stage_bytes = 0;
    // grow-only device scratch kept across calls (multi-GB temporaries: the stream-ordered pool re-maps memory for an
    // allocation of that size on every fit, which was measured to cost more than the kernels using it)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* p2p = nullptr;           // peer-memory all-reduce state (p2p.cu), nullptr = NCCL only
    int sm_reserve = 0;            // SMs the persistent product kernels leave free (a one-CTA kernel running beside them on another stream)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // fork / join of the side stream (copy_stream) inside a fit
};

struct salg_csr {
    salg_ctx* ctx = nullptr;
    int dtype = SALG_F32;
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int64_t* row_ptr = nullptr;  // [nrows+1]
    uint32_t* col = nullptr;     // [nnz]
    void* val = nullptr;         // [nnz] of T
    // lazily built transposed copy (CSR of A^T == CSC of A); invalidated when values change
    mutable bool t_valid = false;
    mutable int64_t* t_ptr = nullptr;  // [ncols+1]
    mutable uint32_t* t_idx = nullptr; // [nnz] row ids, ascending within a column
    mutable void* t_val = nullptr;     // [nnz]
    // SpMM work decomposition: row holding the first entry of every SPMM_CHUNK-sized chunk
    mutable uint32_t* chunk_row = nullptr;
    mutable uint32_t* t_chunk_row = nullptr;
    // tile-densified format for the tcgen05 products (tc.cu); invalidated when values change
    mutable void* tc = nullptr;
};

struct salg_pca {
    salg_ctx* ctx = nullptr;
    int dtype = SALG_F32;
    int64_t d = 0, n_eff = 0, ncols = 0, n_fit_rows_local = 0, n_samples = 0;
    bool center = true;
    bool masked = false;
    std::vector<uint8_t> mask;          // full length (masked models)
    std::vector<double> singular_values, explained_variance, mean_full;
    double total_var = 0.0;
    int numeric_flag = 0;
    void* d_V = nullptr;      // device panel n_eff x 64 of T: column i = component i (sign-flipped)
    void* d_mean = nullptr;   // device T[n_eff] mean of the kept columns (zeros when !center)
    void* d_scores = nullptr; // device n_fit_rows_local x 64 of T (U*S) when keep_scores
    void* d_tscores = nullptr; int64_t tscores_rows = 0;   // last transform_device result
    std::vector<double> col_nnz_kept;   // per kept column stored-entry count (REFERENCE_COMPAT unmasked)
};

namespace salg {

// ---- stream-ordered temporary buffers ------------------------------------------------------------
// Temporaries up to 1 MB are recycled through a per-stream free list (size classes = powers of two) instead of going
// back to the CUDA pool: a fit makes ~250 such allocations, and on the launch-bound small-side chain the two runtime
// calls per buffer were idle time on the GPU.  Reuse is safe because every user of a block is ordered on the one stream
// the block belongs to.  Larger buffers use cudaMallocAsync / cudaFreeAsync directly.  (api.cu)
void* small_alloc(cudaStream_t s, size_t bytes);
void small_free(cudaStream_t s, void* p, size_t bytes);
void small_cache_release(cudaStream_t s);     // context teardown: hand the cached blocks back to the pool
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() = default;
    DevBuf(size_t n_, cudaStream_t s_) { alloc(n_, s_); }
    void alloc(size_t n_, cudaStream_t s_) {
        release();
        n = n_;
        s = s_;
        if (n) p = (T*)small_alloc(s, n * sizeof(T));
    }
    void release() {
        if (p) small_free(s, p, n * sizeof(T));
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
        return *this;
    }
    T* get() const { return p; }
};

// RAII profiling scope: records an event pair around the launches issued inside it.
struct ProfScope {
    salg_ctx* ctx;
    int idx = -1;
    ProfScope(salg_ctx* c, int cls, double bytes);
    ~ProfScope();
};
void prof_collect(salg_ctx* ctx);  // drains pending events into the per-class totals (syncs)

template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int value = SALG_F32; };
template <> struct dtype_of<double> { static constexpr int value = SALG_F64; };

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and device instead of before every launch (the
// power iteration launches ~140 kernels per fit, most of them shorter than the host calls around them)
void set_max_dyn_smem_impl(const void* kernel, int bytes);   // api.cu
template <typename K>
inline void set_max_dyn_smem(K kernel, int bytes) { set_max_dyn_smem_impl((const void*)kernel, bytes); }

// ---- device memory: stream-ordered pool (release threshold = max, see ctx_new) ---------------------------------
// cudaMalloc / cudaFree synchronise the device and were measured to stall a fit by 10-600 ms on multi-GB buffers;
// every buffer of the library therefore comes from the context's stream-ordered pool.  `owner` may already be
// destroyed when a handle is freed late (interpreter shutdown): then fall back to the synchronous call.
bool ctx_alive(const salg_ctx* ctx);
inline void* dev_alloc(salg_ctx* ctx, size_t bytes) {
    void* p = nullptr;
    SALG_CUDA(cudaMallocAsync(&p, bytes ? bytes : 1, ctx->stream));
    return p;
}
inline void dev_free(salg_ctx* owner, void* p) {
    if (!p) return;
    if (ctx_alive(owner)) cudaFreeAsync(p, owner->stream);
    else cudaFree(p);
}

inline void* ctx_scratch(salg_ctx* ctx, size_t bytes) {
    if (bytes > ctx->scratch_bytes) {
        SALG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        SALG_CUDA(cudaMalloc(&ctx->scratch, bytes));
        ctx->scratch_bytes = bytes;
    }
    return ctx->scratch;
}

// ---- csr.cu ---------------------------------------------------------------------------------------
salg_csr* csr_alloc(salg_ctx* ctx, int dtype, int64_t nrows, int64_t ncols, int64_t nnz);
void csr_destroy(salg_csr* c);
void csr_invalidate_transpose(const salg_csr* c);   // call with the stream idle (frees device memory)
template <typename T> void csr_ensure_transpose(salg_ctx* ctx, const salg_csr* c);
// d_row_kept (optional, [nrows+1]): per-row kept counts already computed by col_stats_device -> skips the count pass
template <typename T> salg_csr* csr_select_columns(salg_ctx* ctx, const salg_csr* c, const uint8_t* mask_host,
                                                   int64_t* d_row_kept = nullptr);
void exclusive_scan_i64(salg_ctx* ctx, const int64_t* in, int64_t* out, int64_t n);
salg_csr* csr_upload_i32_f32(salg_ctx* ctx, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* off, const int32_t* idx,
                             const float* val);
salg_csr* csr_shell_from_host_offsets(salg_ctx* ctx, int dtype, int64_t nrows, int64_t ncols, int64_t nnz, const int64_t* h_off);

// ---- stats.cu -------------------------------------------------------------------------------------
// column sums / sums of squares / stored-entry counts in f64 on the device (zeroed here; all-reduced
// over the ranks of a row-sharded context)
// keepbits (ceil(ncols/32) mask words followed by as many exclusive prefix popcounts; n_kept set bits) + row_kept
// [nrows+1]: also count, per row, the entries in kept columns (the compaction's count pass fused into the statistics pass)
template <typename T> void col_stats_device(salg_ctx* ctx, const salg_csr* c, double* d_sum, double* d_sumsq, double* d_cnt,
                                            const uint32_t* keepbits = nullptr, int64_t* row_kept = nullptr,
                                            int64_t n_kept = 0, uint32_t* kept_col = nullptr, void* kept_val = nullptr,
                                            int kept_shift = 0, int* kept_overflow = nullptr);
bool col_stats_can_fuse_compaction(const salg_csr* c, int64_t n_kept);
// streamed fits (pca.cu): the masked statistics + fused compaction pass over one staged row chunk, see stats.cu
bool col_stats_probe_int_f32(salg_ctx* ctx, cudaStream_t st, const float* d_val, int64_t n, int* d_flag1);
void col_stats_masked_chunk_f32(salg_ctx* ctx, cudaStream_t st, const int64_t* ptr, const uint32_t* col, const float* val,
                                int64_t nrows, int64_t ncols, int64_t n_kept, double* d_sum, double* d_sumsq,
                                const uint32_t* keepbits, int64_t* row_kept, uint32_t* kept_col, float* kept_val, int kept_shift,
                                int* flags, bool intsum);
template <typename T> void sum_row_device(salg_ctx* ctx, const salg_csr* c, T* d_out);
int64_t global_nrows(salg_ctx* ctx, int64_t local_rows);

// ---- spmm.cu --------------------------------------------------------------------------------------
// out(nr x 64) = S * X(nc x 64) - alpha * corr^T, S given as CSR arrays (nr rows); val == nullptr => all
// stored values are 1 (pattern product); alpha: per-row T (nullptr => 1), corr: 64 doubles (nullptr => 0).
template <typename T>
void spmm_launch(salg_ctx* ctx, int prof_cls, const int64_t* ptr, const uint32_t* idx, const T* val,
                 const uint32_t* chunk_row, int64_t nr, int64_t nc, int64_t nnz, const T* X, T* out,
                 const T* alpha, const double* corr);
uint32_t* build_chunk_rows(salg_ctx* ctx, const int64_t* ptr, int64_t nr, int64_t nnz);
// out(nrows x 64) = A X - 1 corr^T
template <typename T> void spmm_A(salg_ctx* ctx, const salg_csr* c, const T* X, T* out, const double* corr, bool pattern);
// out(ncols x 64) = A^T Y - mu corr^T  (gather over the transposed copy; local rows only)
template <typename T> void spmm_At(salg_ctx* ctx, const salg_csr* c, const T* Y, T* out, const T* mu, const double* corr);

// ---- tc.cu ----------------------------------------------------------------------------------------
// tile-densified tcgen05 products (f32 operators only).  SALG_SPMM_IMPL=chunk selects the CUDA-core kernels.
bool tc_enabled(const salg_ctx* ctx);
void tc_free(salg_ctx* owner, void* tiles);
// tile format of the operator whose row r lives at [in_ptr[r] >> in_shift, ... + c->row_ptr[r+1] - c->row_ptr[r]) of col / val
// (fused compaction: the kept entries sit at the rows' scaled ORIGINAL offsets); attaches it to c
void tc_attach_tiles_f32(salg_ctx* ctx, salg_csr* c, const int64_t* in_ptr, int in_shift, const uint32_t* col, const float* val);
void tc_spmm_A(salg_ctx* ctx, const salg_csr* c, const float* X, float* Y, const double* corr, unsigned* d_amax = nullptr,
               int b_terms = 2);
size_t tc_yprep_bytes(salg_ctx* ctx, const salg_csr* c);
// no_gram: only the column sums (G + 64*64 ..) and the pre-split operand are produced (the Gram part of G stays zero)
void tc_gram_prep(salg_ctx* ctx, const salg_csr* c, const float* Y, const unsigned* d_amax, uint8_t* Yprep, float* d_scales,
                  double* G /*GRAM_BUF*/, bool no_gram = false);
void tc_gram_probe(salg_ctx* ctx, const float* Y, int64_t m, double* G, uint8_t* yprep_out, int iters, double* avg_ms);
void tc_set_amax(salg_ctx* ctx, unsigned* d_amax, float bound);
void tc_spmm_At_prepped(salg_ctx* ctx, const salg_csr* c, const uint8_t* Yprep, const float* d_scales, float* Z,
                        const float* mu, const double* corr);
void tc_spmm_At(salg_ctx* ctx, const salg_csr* c, const float* Y, float* Z, const float* mu, const double* corr);
// fused small-side step of the power iteration (TMEM-operand products): zside_solve -> tc_zside_apply -> tc_spmm_A_prepped
bool tc_zside_supported(salg_ctx* ctx, const salg_csr* c);
size_t tc_xprep_bytes(salg_ctx* ctx, const salg_csr* c);
float tc_a_scale(salg_ctx* ctx, const salg_csr* c);
void tc_zside_apply(salg_ctx* ctx, const salg_csr* c, float* Z, const float* d_M, const float* mu, const float* d_scales,
                    uint8_t* Xprep, double* corr);
void tc_spmm_A_prepped(salg_ctx* ctx, const salg_csr* c, const uint8_t* Xprep, const float* d_scales, float* Y, const double* corr,
                       unsigned* d_amax, int b_terms);

// ---- dense.cu -------------------------------------------------------------------------------------
template <typename T> void panel_gram(salg_ctx* ctx, const T* P, int64_t m, double* d_out /*GRAM_BUF*/);
// drop: numerically dependent columns get a zero column in R^{-1} (zero column of Q) instead of a floored pivot (dense.cu)
template <typename T> void chol_inv(salg_ctx* ctx, const double* d_G, int k, double* d_R, double* d_Rinv, T* d_RinvT, int* d_flag,
                                    bool drop = false);
// d_amax (optional, zeroed here): receives the bits of max |out| as a float
template <typename T> void panel_mul(salg_ctx* ctx, const T* P, int64_t m, const T* d_M /*64x64 row-major*/, T* out,
                                     unsigned* d_amax = nullptr);
void mat64_mul(salg_ctx* ctx, const double* A, const double* B, double* C);  // C = A*B, 64x64 f64
void vec64_mat(salg_ctx* ctx, const double* v, const double* M, double* o);  // o = v*M
void jacobi_svd64(salg_ctx* ctx, const double* d_A, int k, double* d_U, double* d_S, double* d_V, int* d_flag);
template <typename T> void cast_mat64(salg_ctx* ctx, const double* src, T* dst, const double* colscale);
template <typename T> void panel_colsum(salg_ctx* ctx, const T* P, int64_t m, const T* w, double* d_out64);
template <typename T> void flip_find(salg_ctx* ctx, const T* V, int64_t n_eff, double* d_sign64);
template <typename T> void panel_colscale(salg_ctx* ctx, T* P, int64_t m, const double* d_scale64);
template <typename T> void panel_to_rowmajor_t(salg_ctx* ctx, const T* V, int64_t n, int d, T* out /* d x n */);
template <typename T> void panel_pack(salg_ctx* ctx, const T* src, int64_t m, int k, T* dst);      // (m x k) -> (m x 64) zero padded
template <typename T> void panel_unpack(salg_ctx* ctx, const T* src, int64_t m, int k, T* dst);    // (m x 64) -> (m x k)
// dense.cu: Cholesky of Gy, Gram of the raw Z, Cholesky of the transformed Gram -> M = R1^{-1} R2^{-1}, pre-split scale; zeroes corr
void zside_solve(salg_ctx* ctx, const float* Z, int64_t n, const double* d_Gy, int k, float a_scale, double* d_part,
                 unsigned* d_ticket, float* d_M, float* d_scales, double* d_corr, int* d_flag);
size_t zside_part_elems();
int zside_mode();
template <typename T> void cholqr2(salg_ctx* ctx, T* Y, int64_t m_local, int k, bool sharded, double* d_colsum64,
                                   double* d_Rtot, int* d_flag, int passes);

// p2p.cu: one-shot all-reduce over NVLink peer memory (small operands); collective init right after the NCCL communicator
void p2p_init(salg_ctx* ctx);
void p2p_destroy(salg_ctx* ctx);
bool p2p_allreduce(salg_ctx* ctx, double* buf64, size_t n64, float* buf32, size_t n32);

void allreduce_f64(salg_ctx* ctx, double* buf, size_t n);
void allreduce_gram_and_panel(salg_ctx* ctx, double* gram, size_t n_gram, float* panel, size_t n_panel);
template <typename T> void allreduce_T(salg_ctx* ctx, T* buf, size_t n);

// ---- pca.cu ---------------------------------------------------------------------------------------
void mul64_kernel_launch(salg_ctx* ctx, const double* a, const double* b, double* o);   // o = a .* b (64)
// all-reduce of the per-rank partials of A^T Y - mu cs^T (restores the single rank-1 correction)
template <typename T> void allreduce_panel_T(salg_ctx* ctx, T* Z, size_t n, const T* mu, const double* cs, int64_t n_eff);

// ---- lanczos.cu -----------------------------------------------------------------------------------
// Golub-Kahan-Lanczos with full reorthogonalisation on the UNCENTRED operator.
// Returns d (<= k) converged-or-best triplets: d_V panel (n_eff x 64, column i = right vector i), s.
template <typename T>
int lanczos_svd(salg_ctx* ctx, const salg_csr* op, int k, int max_steps, uint64_t seed, double tol, T* d_Vpanel,
                std::vector<double>& s_out, int* steps_out);

}  // namespace salg
