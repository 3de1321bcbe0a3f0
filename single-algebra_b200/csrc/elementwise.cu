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

// Row streams: one warp per row, kRowU independent 128 B loads in flight per lane array (a B200 SM needs ~40 KB in
// flight to keep HBM busy; one load per lane and trip left these passes latency bound at ~55 % of the roofline).
constexpr int kRowU = 16;

// ROW: rows whose scale is <= 0 are left untouched (csr.rs:1054-1055)
template <typename T, typename U>
__global__ void normalize_row_kernel(const int64_t* __restrict__ ptr, T* __restrict__ val, int64_t nrows,
                                     const U* __restrict__ scale) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int RU = sizeof(T) == 4 ? kRowU : kRowU / 2;
    for (int64_t r = w; r < nrows; r += nw) {
        U sc = scale[r];
        if (!(sc > U(0))) continue;
        const int64_t s = ptr[r];
        const uint32_t len = (uint32_t)(ptr[r + 1] - s);
        T* __restrict__ vr = val + s;
        // whole batches carry no per-load predicate (seven predicate registers would cap the loads in flight at six);
        // the tail batch clamps its index to the row's last entry and guards the stores
        uint32_t p0 = 0;
        for (; p0 + 32 * RU <= len; p0 += 32 * RU) {
            T v[RU];
#pragma unroll
            for (int u = 0; u < RU; u++) v[u] = vr[p0 + lane + 32 * u];
#pragma unroll
            for (int u = 0; u < RU; u++) vr[p0 + lane + 32 * u] = (T)((U)v[u] * sc);  // csr.rs:1060
        }
        if (p0 < len) {
            T v[RU];
            const uint32_t last = len - 1;
#pragma unroll
            for (int u = 0; u < RU; u++) {
                const uint32_t q = p0 + lane + 32 * u;
                v[u] = vr[q < last ? q : last];
            }
            __syncwarp();                                  // every clamped read of the last entry precedes its update
#pragma unroll
            for (int u = 0; u < RU; u++) {
                const uint32_t q = p0 + lane + 32 * u;
                if (q < len) vr[q] = (T)((U)v[u] * sc);
            }
        }
    }
}

// COLUMN: flat stream, 4 entries per thread and trip, scale gathered by column id (csr.rs:1036-1044).
// The arrays carry 16 entries of slack and start 256 B aligned, so whole quads are loaded; only stores are guarded.
template <typename T, typename U>
__global__ void normalize_col_kernel(const uint32_t* __restrict__ col, T* __restrict__ val, int64_t nnz,
                                     const U* __restrict__ scale) {
    const int64_t n4 = (nnz + 3) >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const uint4 c4 = __ldcs(reinterpret_cast<const uint4*>(col) + i);
        const uint32_t cc[4] = {c4.x, c4.y, c4.z, c4.w};
        T* v = val + (i << 2);
        const int64_t left = nnz - (i << 2);
        if (left >= 4) {
            T x[4];
            if (sizeof(T) == 4) {
                const float4 f = *reinterpret_cast<const float4*>(v);
                x[0] = (T)f.x; x[1] = (T)f.y; x[2] = (T)f.z; x[3] = (T)f.w;
            } else {
                const double2 d0 = *reinterpret_cast<const double2*>(v), d1 = *(reinterpret_cast<const double2*>(v) + 1);
                x[0] = (T)d0.x; x[1] = (T)d0.y; x[2] = (T)d1.x; x[3] = (T)d1.y;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const U sc = scale[cc[k]];
                if (sc > U(0)) x[k] = (T)((U)x[k] * sc);
            }
            if (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(v) = make_float4((float)x[0], (float)x[1], (float)x[2], (float)x[3]);
            } else {
                *reinterpret_cast<double2*>(v) = make_double2((double)x[0], (double)x[1]);
                *(reinterpret_cast<double2*>(v) + 1) = make_double2((double)x[2], (double)x[3]);
            }
        } else {
            for (int k = 0; k < (int)left; k++) {
                const U sc = scale[cc[k]];
                if (sc > U(0)) v[k] = (T)((U)v[k] * sc);
            }
        }
    }
}

// v <- ln(fl(1 + v)): two roundings in T, not log1p (csr.rs:1074-1075).  128-bit accesses, two quads per
// thread and trip.
template <typename T>
__device__ __forceinline__ T ln_1p_two_step(T v) {
    const T x = T(1) + v;
    return (T)log(x);
}

template <typename T>
__global__ void log1p_kernel(T* __restrict__ val, int64_t nnz) {
    constexpr int Q = 16 / sizeof(T);                // entries per 128-bit access
    const int64_t nq = nnz / Q;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sizeof(T) == 4) {
        float4* v4 = reinterpret_cast<float4*>(val);
        for (; i + stride < nq; i += 2 * stride) {
            float4 a = v4[i], b = v4[i + stride];
            a.x = ln_1p_two_step(a.x); a.y = ln_1p_two_step(a.y); a.z = ln_1p_two_step(a.z); a.w = ln_1p_two_step(a.w);
            b.x = ln_1p_two_step(b.x); b.y = ln_1p_two_step(b.y); b.z = ln_1p_two_step(b.z); b.w = ln_1p_two_step(b.w);
            v4[i] = a;
            v4[i + stride] = b;
        }
        if (i < nq) {
            float4 a = v4[i];
            a.x = ln_1p_two_step(a.x); a.y = ln_1p_two_step(a.y); a.z = ln_1p_two_step(a.z); a.w = ln_1p_two_step(a.w);
            v4[i] = a;
        }
    } else {
        double2* v2 = reinterpret_cast<double2*>(val);
        for (; i + stride < nq; i += 2 * stride) {
            double2 a = v2[i], b = v2[i + stride];
            a.x = ln_1p_two_step(a.x); a.y = ln_1p_two_step(a.y);
            b.x = ln_1p_two_step(b.x); b.y = ln_1p_two_step(b.y);
            v2[i] = a;
            v2[i + stride] = b;
        }
        if (i < nq) {
            double2 a = v2[i];
            a.x = ln_1p_two_step(a.x); a.y = ln_1p_two_step(a.y);
            v2[i] = a;
        }
    }
    // tail (< Q entries)
    const int64_t t = nq * Q + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nnz) val[t] = ln_1p_two_step(val[t]);
}

// fused row pass: s = sum(row) (f64 accumulate, rounded to T as sum_row returns it),
// scale = target / s if s > 0 (U = T), v <- ln(1 + T(v * scale)); rows with s <= 0 only get ln(1 + v).
// Persistent grid of 4 CTAs per SM: the second touch of a row is served by L2 as long as the rows in flight across the
// chip (one per warp) stay well below its capacity.
template <typename T>
__global__ void __launch_bounds__(256) preprocess_row_kernel(const int64_t* __restrict__ ptr, T* __restrict__ val,
                                                             int64_t nrows, T target) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int RU = sizeof(T) == 4 ? kRowU : kRowU / 2;
    for (int64_t r = w; r < nrows; r += nw) {
        const int64_t s = ptr[r];
        const uint32_t len = (uint32_t)(ptr[r + 1] - s);
        T* __restrict__ vr = val + s;
        double a = 0.0;
        T v[RU];
        const uint32_t last = len ? len - 1 : 0;
        uint32_t p0 = 0;
        for (; p0 + 32 * RU <= len; p0 += 32 * RU) {       // whole batches: no per-load predicate
#pragma unroll
            for (int u = 0; u < RU; u++) v[u] = vr[p0 + lane + 32 * u];
#pragma unroll
            for (int u = 0; u < RU; u++) a += (double)v[u];
        }
        if (p0 < len) {                                    // tail batch: clamped index, value zeroed afterwards
#pragma unroll
            for (int u = 0; u < RU; u++) {
                const uint32_t q = p0 + lane + 32 * u;
                v[u] = vr[q < last ? q : last];
            }
#pragma unroll
            for (int u = 0; u < RU; u++) {
                if (!(p0 + lane + 32 * u < len)) v[u] = T(0);
                a += (double)v[u];
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        const T sum = (T)a;
        const T sc = sum > T(0) ? target / sum : T(0);
        const bool scaled = sc > T(0);
        if (len <= 32 * RU) {                              // the whole row is still in registers (v of the only batch)
#pragma unroll
            for (int u = 0; u < RU; u++) {
                const uint32_t q = lane + 32 * u;
                T x = v[u];
                if (scaled) x = (T)(x * sc);
                if (q < len) __stcs(vr + q, ln_1p_two_step(x));
            }
            continue;
        }
        // second touch: served by L2 (the row was just read).  Last-use loads and streaming stores keep the rows that
        // are between their two touches resident instead of the lines nobody will read again.
        for (p0 = 0; p0 + 32 * RU <= len; p0 += 32 * RU) {
#pragma unroll
            for (int u = 0; u < RU; u++) v[u] = __ldlu(vr + p0 + lane + 32 * u);
#pragma unroll
            for (int u = 0; u < RU; u++) {
                T x = v[u];
                if (scaled) x = (T)(x * sc);
                __stcs(vr + p0 + lane + 32 * u, ln_1p_two_step(x));
            }
        }
        if (p0 < len) {
#pragma unroll
            for (int u = 0; u < RU; u++) {
                const uint32_t q = p0 + lane + 32 * u;
                v[u] = q < len ? __ldlu(vr + q) : T(0);    // (no clamped re-read of an entry another lane is updating)
            }
#pragma unroll
            for (int u = 0; u < RU; u++) {
                const uint32_t q = p0 + lane + 32 * u;
                T x = v[u];
                if (scaled) x = (T)(x * sc);
                if (q < len) __stcs(vr + q, ln_1p_two_step(x));
            }
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
        normalize_col_kernel<T, U><<<flat_grid(ctx, (c->nnz + 3) / 4, 256), 256, 0, st>>>(c->col, (T*)c->val, c->nnz,
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
        log1p_kernel<T><<<flat_grid(ctx, c->nnz / (16 / sizeof(T)) + 16, 256), 256, 0, ctx->stream>>>((T*)c->val, c->nnz);
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
        static const int ctas_per_sm = getenv("SALG_PRE_CTAS") ? atoi(getenv("SALG_PRE_CTAS")) : 4;
        int64_t want = ceil_div(c->nrows * 32, 256), cap = (int64_t)ctx->sm_count * ctas_per_sm;
        preprocess_row_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(c->row_ptr, (T*)c->val, c->nrows,
                                                                                      target);
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
