// elementwise.cu — Normalize::normalize (src/sparse/csr.rs:1013-1068), Log1P::log1p_normalize
// (src/sparse/csr.rs:1070-1079) and the fused preprocessing pass sum_row -> normalize(ROW) -> log1p
// (SURVEY K10-K12).  All in place on the device-resident values; bandwidth-bound streams.
#include "common.cuh"

namespace salg {

// scale[i] = sums[i] > 0 ? target / sums[i] : 0     (csr.rs:1021-1030), arithmetic in U
template <typename U>
__global__ void make_scale_kernel(const U* __restrict__ sums, U target, U* __restrict__ scale, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        U s = sums[i];
        scale[i] = s > U(0) ? target / s : U(0);
    }
}

// ROW: one warp per row; rows whose scale is <= 0 are left untouched (csr.rs:1054-1055)
template <typename T, typename U>
__global__ void normalize_row_kernel(const int64_t* __restrict__ ptr, T* __restrict__ val, int64_t nrows,
                                     const U* __restrict__ scale) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < nrows; r += nw) {
        U sc = scale[r];
        if (!(sc > U(0))) continue;
        int64_t s = ptr[r], e = ptr[r + 1];
        for (int64_t p = s + lane; p < e; p += 32) val[p] = (T)((U)val[p] * sc);  // csr.rs:1060
    }
}

// COLUMN: flat stream, scale gathered by column id (csr.rs:1036-1044)
template <typename T, typename U>
__global__ void normalize_col_kernel(const uint32_t* __restrict__ col, T* __restrict__ val, int64_t nnz,
                                     const U* __restrict__ scale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        U sc = scale[col[i]];
        if (sc > U(0)) val[i] = (T)((U)val[i] * sc);
    }
}

// v <- ln(fl(1 + v)): two roundings in T, not log1p (csr.rs:1074-1075)
template <typename T>
__global__ void log1p_kernel(T* __restrict__ val, int64_t nnz) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        T x = T(1) + val[i];
        val[i] = (T)log(x);
    }
}

// fused row pass: s = sum(row) (f64 accumulate, rounded to T as sum_row returns it),
// scale = target / s if s > 0 (U = T), v <- ln(1 + T(v * scale)); rows with s <= 0 only get ln(1 + v).
template <typename T>
__global__ void preprocess_row_kernel(const int64_t* __restrict__ ptr, T* __restrict__ val, int64_t nrows,
                                      T target) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < nrows; r += nw) {
        int64_t s = ptr[r], e = ptr[r + 1];
        double a = 0.0;
        for (int64_t p = s + lane; p < e; p += 32) a += (double)val[p];
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        T sum = (T)a;
        T sc = sum > T(0) ? target / sum : T(0);
        bool scaled = sc > T(0);
        for (int64_t p = s + lane; p < e; p += 32) {   // second touch hits L1/L2: the row was just read
            T v = val[p];
            if (scaled) v = (T)(v * sc);
            val[p] = (T)log(T(1) + v);
        }
    }
}

static int flat_grid(salg_ctx* ctx, int64_t n, int block) {
    int64_t want = ceil_div(n > 0 ? n : 1, block);
    int64_t cap = (int64_t)ctx->sm_count * 16;
    return (int)(want < cap ? want : cap);
}

template <typename T, typename U>
static void normalize_api(salg_ctx* ctx, salg_csr* c, const U* sums, int64_t n_sums, U target, int direction) {
    SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
    SALG_REQUIRE(c->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "csr value type does not match the entry point");
    SALG_REQUIRE(direction == SALG_ROW || direction == SALG_COLUMN, SALG_ERR_BAD_ARG, "direction must be ROW or COLUMN");
    int64_t need = direction == SALG_ROW ? c->nrows : c->ncols;
    // the reference indexes `sums` unchecked and panics when it is short -> error code here
    SALG_REQUIRE(n_sums >= need, SALG_ERR_BAD_ARG, "sums is shorter than the normalised dimension");
    SALG_REQUIRE(sums || need == 0, SALG_ERR_BAD_ARG, "sums is NULL");
    SALG_CUDA(cudaSetDevice(ctx->device));
    if (need == 0 || c->nnz == 0) return;
    cudaStream_t st = ctx->stream;
    DevBuf<U> d_sums((size_t)need, st), d_scale((size_t)need, st);
    SALG_CUDA(cudaMemcpyAsync(d_sums.get(), sums, (size_t)need * sizeof(U), cudaMemcpyHostToDevice, st));
    make_scale_kernel<U><<<(unsigned)ceil_div(need, 256), 256, 0, st>>>(d_sums.get(), target, d_scale.get(), need);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    if (direction == SALG_ROW) {
        ProfScope ps(ctx, PROF_ELEMENTWISE,
                     2.0 * (double)c->nnz * sizeof(T) + (double)(c->nrows + 1) * 8 + (double)c->nrows * sizeof(U));
        normalize_row_kernel<T, U><<<flat_grid(ctx, c->nrows * 32, 256), 256, 0, st>>>(c->row_ptr, (T*)c->val,
                                                                                       c->nrows, d_scale.get());
        ctx->n_launch++;
    } else {
        ProfScope ps(ctx, PROF_ELEMENTWISE,
                     2.0 * (double)c->nnz * sizeof(T) + (double)c->nnz * 4 + (double)c->ncols * sizeof(U));
        normalize_col_kernel<T, U><<<flat_grid(ctx, c->nnz, 256), 256, 0, st>>>(c->col, (T*)c->val, c->nnz,
                                                                                d_scale.get());
        ctx->n_launch++;
    }
    SALG_CUDA(cudaGetLastError());
    SALG_CUDA(cudaStreamSynchronize(st));
    csr_invalidate_transpose(c);
}

template <typename T>
static void log1p_api(salg_ctx* ctx, salg_csr* c) {
    if (c->nnz == 0) return;
    {
        ProfScope ps(ctx, PROF_ELEMENTWISE, 2.0 * (double)c->nnz * sizeof(T));
        log1p_kernel<T><<<flat_grid(ctx, c->nnz, 256), 256, 0, ctx->stream>>>((T*)c->val, c->nnz);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    SALG_CUDA(cudaStreamSynchronize(ctx->stream));
    csr_invalidate_transpose(c);
}

template <typename T>
__global__ void cast_f64_to_T_kernel(const double* __restrict__ src, T* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (T)src[i];
}

template <typename T>
static void preprocess_api(salg_ctx* ctx, salg_csr* c, T target, T* col_sum, T* col_sumsq) {
    SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
    SALG_REQUIRE(c->dtype == dtype_of<T>::value, SALG_ERR_BAD_ARG, "csr value type does not match the entry point");
    SALG_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (c->nrows && c->nnz) {
        ProfScope ps(ctx, PROF_ELEMENTWISE, 2.0 * (double)c->nnz * sizeof(T) + (double)(c->nrows + 1) * 8);
        preprocess_row_kernel<T><<<flat_grid(ctx, c->nrows * 32, 256), 256, 0, st>>>(c->row_ptr, (T*)c->val,
                                                                                     c->nrows, target);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    }
    SALG_CUDA(cudaStreamSynchronize(st));
    csr_invalidate_transpose(c);
    int64_t n = c->ncols;
    if ((col_sum || col_sumsq) && n) {
        DevBuf<double> d_sum((size_t)n, st), d_sq((size_t)n, st);
        col_stats_device<T>(ctx, c, d_sum.get(), d_sq.get(), nullptr);
        DevBuf<T> o1((size_t)n, st), o2((size_t)n, st);
        cast_f64_to_T_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_sum.get(), o1.get(), n);
        ctx->n_launch++;
        cast_f64_to_T_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_sq.get(), o2.get(), n);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        if (col_sum) SALG_CUDA(cudaMemcpyAsync(col_sum, o1.get(), (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
        if (col_sumsq) SALG_CUDA(cudaMemcpyAsync(col_sumsq, o2.get(), (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
    }
    SALG_CUDA(cudaStreamSynchronize(st));
}

}  // namespace salg

using namespace salg;

extern "C" {

int salg_normalize_f32(salg_ctx* ctx, salg_csr* c, const float* sums, int64_t n, float target, int dir) {
    return guarded([&] { normalize_api<float, float>(ctx, c, sums, n, target, dir); });
}
int salg_normalize_f64(salg_ctx* ctx, salg_csr* c, const double* sums, int64_t n, double target, int dir) {
    return guarded([&] { normalize_api<double, double>(ctx, c, sums, n, target, dir); });
}
int salg_normalize_f32_u64(salg_ctx* ctx, salg_csr* c, const double* sums, int64_t n, double target, int dir) {
    return guarded([&] { normalize_api<float, double>(ctx, c, sums, n, target, dir); });
}

/* CscMatrix::normalize (src/sparse/csc.rs:680-735): same arithmetic; on the stored CSR of A^T the roles of ROW and
 * COLUMN are exchanged. */
static int csc_dir(int dir) { return dir == SALG_ROW ? SALG_COLUMN : (dir == SALG_COLUMN ? SALG_ROW : dir); }
int salg_csc_normalize_f32(salg_ctx* ctx, salg_csr* c, const float* sums, int64_t n, float target, int dir) {
    return guarded([&] { normalize_api<float, float>(ctx, c, sums, n, target, csc_dir(dir)); });
}
int salg_csc_normalize_f64(salg_ctx* ctx, salg_csr* c, const double* sums, int64_t n, double target, int dir) {
    return guarded([&] { normalize_api<double, double>(ctx, c, sums, n, target, csc_dir(dir)); });
}
int salg_csc_normalize_f32_u64(salg_ctx* ctx, salg_csr* c, const double* sums, int64_t n, double target, int dir) {
    return guarded([&] { normalize_api<float, double>(ctx, c, sums, n, target, csc_dir(dir)); });
}

int salg_log1p(salg_ctx* ctx, salg_csr* c) {
    return guarded([&] {
        SALG_REQUIRE(ctx && c, SALG_ERR_BAD_ARG, "ctx/csr is NULL");
        SALG_CUDA(cudaSetDevice(ctx->device));
        if (c->dtype == SALG_F64) log1p_api<double>(ctx, c); else log1p_api<float>(ctx, c);
    });
}

int salg_preprocess_f32(salg_ctx* ctx, salg_csr* c, float target, float* col_sum, float* col_sumsq) {
    return guarded([&] { preprocess_api<float>(ctx, c, target, col_sum, col_sumsq); });
}
int salg_preprocess_f64(salg_ctx* ctx, salg_csr* c, double target, double* col_sum, double* col_sumsq) {
    return guarded([&] { preprocess_api<double>(ctx, c, target, col_sum, col_sumsq); });
}

}  // extern "C"
