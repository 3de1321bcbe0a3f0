// lanczos.cu — SVDMethod::Lanczos: replaces single-svdlib's lanczos::svd_las2 as called at
// pca/sparse/mod.rs:136-144 and pca/sparse_masked/mod.rs:322-330.
//
// las2 runs single-vector Lanczos on A^T A with selective reorthogonalisation; here the same Krylov space
// is built by Golub-Kahan-Lanczos bidiagonalisation (A V = U B, A^T U = V B^T + beta v e^T) with FULL
// reorthogonalisation (two classical Gram-Schmidt passes against the whole basis, both sides), the Ritz
// values come from the tridiagonal T = B^T B by implicit QL on the host (the analogue of las2's imtql2),
// and the pair (theta, y) is accepted on the classical bound alpha_j beta_j |y_j| <= tol * theta_1.
// The operator is the UNCENTRED matrix even when center == true — that is what the reference passes
// (SURVEY §0.6).  The two SpMVs per step are the HBM-bound part: nnz*(S+I) bytes each.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace salg {

// ---- SpMV: y = S x, S as CSR arrays --------------------------------------------------------------------------
// partial dot product of one row for thread `tid` of `NT` cooperating threads: U independent index / value loads in
// flight per thread, then U gathers of x.  Whole batches carry no per-load predicate (ptxas has seven predicate
// registers: predicated loads go out six at a time); the tail batch clamps its index to the row's last entry.
template <typename T, int NT>
__device__ __forceinline__ double row_dot_partial(const uint32_t* __restrict__ ir, const T* __restrict__ vr, uint32_t len,
                                                  const T* __restrict__ x, int tid) {
    constexpr int U = 8;
    double a = 0.0;
    uint32_t p0 = 0;
    for (; p0 + NT * U <= len; p0 += NT * U) {
        uint32_t c[U];
        T v[U], xv[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            c[u] = __ldcs(ir + p0 + tid + NT * u);
            v[u] = __ldcs(vr + p0 + tid + NT * u);
        }
#pragma unroll
        for (int u = 0; u < U; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
        for (int u = 0; u < U; u++) a = fma((double)v[u], (double)xv[u], a);
    }
    if (p0 < len) {
        uint32_t c[U];
        T v[U], xv[U];
        const uint32_t last = len - 1;
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint32_t q = p0 + tid + NT * u;
            q = q < last ? q : last;
            c[u] = __ldcs(ir + q);
            v[u] = __ldcs(vr + q);
        }
#pragma unroll
        for (int u = 0; u < U; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
        for (int u = 0; u < U; u++)
            if (p0 + tid + NT * u < len) a = fma((double)v[u], (double)xv[u], a);
    }
    return a;
}

template <typename T>
__global__ void __launch_bounds__(256, 3)
spmv_warp_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ idx, const T* __restrict__ val,
                 int64_t nr, const T* __restrict__ x, T* __restrict__ y) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < nr; r += nw) {
        const int64_t s = ptr[r];
        double a = row_dot_partial<T, 32>(idx + s, val + s, (uint32_t)(ptr[r + 1] - s), x, lane);
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        if (lane == 0) y[r] = (T)a;
    }
}

// one CTA per row, for operands whose rows are long and skewed (gene rows of the transposed copy)
template <typename T>
__global__ void __launch_bounds__(256, 3)
spmv_block_kernel(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ idx, const T* __restrict__ val,
                  int64_t nr, const T* __restrict__ x, T* __restrict__ y) {
    __shared__ double part[8];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = blockIdx.x; r < nr; r += gridDim.x) {
        const int64_t s = ptr[r];
        double a = row_dot_partial<T, 256>(idx + s, val + s, (uint32_t)(ptr[r + 1] - s), x, (int)threadIdx.x);
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        __syncthreads();
        if (lane == 0) part[warp] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < 8; i++) t += part[i];
            y[r] = (T)t;
        }
    }
}

template <typename T>
static void spmv(salg_ctx* ctx, const int64_t* ptr, const uint32_t* idx, const T* val, int64_t nr, int64_t nc,
                 int64_t nnz, const T* x, T* y) {
    if (nr == 0) return;
    ProfScope ps(ctx, PROF_SPMV, (double)nnz * (sizeof(T) + 4) + (double)(nr + 1) * 8 + (double)(nr + nc) * sizeof(T));
    bool long_rows = nnz / (nr > 0 ? nr : 1) >= 2048;
    if (long_rows) {
        int64_t cap = (int64_t)ctx->sm_count * 8;
        spmv_block_kernel<T><<<(unsigned)(nr < cap ? nr : cap), 256, 0, ctx->stream>>>(ptr, idx, val, nr, x, y);
        ctx->n_launch++;
    } else {
        int64_t want = ceil_div(nr * 32, 256);
        int64_t cap = (int64_t)ctx->sm_count * 8;
        spmv_warp_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(ptr, idx, val, nr, x, y);
        ctx->n_launch++;
    }
    SALG_CUDA(cudaGetLastError());
}

// ---- dense basis kernels -------------------------------------------------------------------------------------------
// c[i] = <B_i, w>, i < j  (basis rows are contiguous vectors of length n); c zeroed by the caller
template <typename T>
__global__ void __launch_bounds__(256)
basis_dots_kernel(const T* __restrict__ B, int64_t n, int j, const T* __restrict__ w, double* __restrict__ c) {
    __shared__ double part[8];
    int i = blockIdx.y;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* b = B + (size_t)i * n;
    double a = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        a = fma((double)b[p], (double)w[p], a);
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
    if (lane == 0) part[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += part[k];
        atomicAdd(&c[i], t);
    }
}

// w[p] -= sum_i c[i] B_i[p]
template <typename T>
__global__ void basis_update_kernel(const T* __restrict__ B, int64_t n, int j, const double* __restrict__ c,
                                    T* __restrict__ w) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double a = (double)w[p];
    for (int i = 0; i < j; i++) a = fma(-c[i], (double)B[(size_t)i * n + p], a);
    w[p] = (T)a;
}

template <typename T>
__global__ void __launch_bounds__(256) sqnorm_kernel(const T* __restrict__ w, int64_t n, double* __restrict__ out) {
    __shared__ double part[8];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256) {
        double x = (double)w[p];
        a = fma(x, x, a);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
    if (lane == 0) part[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += part[k];
        atomicAdd(out, t);
    }
}

// dst = w / sqrt(sq) and coef = sqrt(sq); a numerically vanished vector (invariant subspace reached) is
// stored as 0 with coef = 0 so the host can truncate the factorisation there.
template <typename T>
__global__ void normalize_store_kernel(const T* __restrict__ w, int64_t n, const double* __restrict__ sq,
                                       const double* __restrict__ scale_ref, double rel_tiny, T* __restrict__ dst,
                                       double* __restrict__ coef) {
    double nrm = sqrt(*sq);
    double ref = scale_ref ? *scale_ref : 0.0;
    bool dead = !(nrm > rel_tiny * ref) || !(nrm > 0.0);
    double inv = dead ? 0.0 : 1.0 / nrm;
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[p] = (T)((double)w[p] * inv);
    if (p == 0) *coef = dead ? 0.0 : nrm;
}

// out panel (n x 64): out[p][i] = sum_t B_t[p] * Q[t][i], Q (j x 64) row-major f64
template <typename T>
__global__ void ritz_vectors_kernel(const T* __restrict__ B, int64_t n, int j, const double* __restrict__ Q,
                                    T* __restrict__ out) {
    int64_t p = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 6);
    int i = threadIdx.x & 63;
    if (p >= n) return;
    double a = 0.0;
    for (int t = 0; t < j; t++) a = fma((double)B[(size_t)t * n + p], Q[(size_t)t * LP + i], a);
    out[p * LP + i] = (T)a;
}

// ---- host: eigen-decomposition of a symmetric tridiagonal matrix by implicit QL --------------------------------------
// d[0..n): diagonal (overwritten by eigenvalues), e[0..n-1): sub-diagonal (destroyed; e[n-1] unused).
// Z: zrows x n row-major, receives Z <- Z * (eigenvector matrix); pass identity for the vectors or a single
// row e_n^T for the last components only.  Returns false when an eigenvalue fails to converge.
static bool tridiag_ql(std::vector<double>& d, std::vector<double>& e, int n, std::vector<double>& Z, int zrows) {
    if (n <= 1) return true;
    e[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 80) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--) {
                    double f = s * e[i], b = c * e[i];
                    r = std::hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    for (int k = 0; k < zrows; k++) {
                        double* z = &Z[(size_t)k * n];
                        double f2 = z[i + 1];
                        z[i + 1] = s * z[i] + c * f2;
                        z[i] = c * z[i] - s * f2;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    return true;
}

static void host_normal(std::vector<double>& out, uint64_t seed) {
    // splitmix64 + Box-Muller: the start vector only has to be generic (las2 also seeds it randomly)
    uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    auto next = [&]() {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    };
    for (size_t i = 0; i < out.size(); i += 2) {
        double u1 = ((next() >> 11) + 1.0) / 9007199254740993.0;
        double u2 = (next() >> 11) / 9007199254740992.0;
        double r = std::sqrt(-2.0 * std::log(u1));
        out[i] = r * std::cos(6.283185307179586 * u2);
        if (i + 1 < out.size()) out[i + 1] = r * std::sin(6.283185307179586 * u2);
    }
}

template <typename T>
int lanczos_svd(salg_ctx* ctx, const salg_csr* op, int k, int max_steps, uint64_t seed, double tol, T* d_Vpanel,
                std::vector<double>& s_out, int* steps_out) {
    cudaStream_t st = ctx->stream;
    const int64_t nr = op->nrows, n = op->ncols;
    const int64_t nr_total = global_nrows(ctx, nr);
    // step limit: the reference passes iterations = max(n_samples, n_features) (pca/sparse/mod.rs:134-142), which las2
    // clamps to min(nrows, ncols); lanczos_max_steps > 0 lowers it
    int64_t mmax = std::min<int64_t>(n, nr_total);
    if (max_steps > 0) mmax = std::min<int64_t>(mmax, max_steps);
    SALG_REQUIRE(mmax >= 1, SALG_ERR_BAD_ARG, "empty operator");
    const int m = (int)mmax;
    csr_ensure_transpose<T>(ctx, op);
    const double eps_t = sizeof(T) == 4 ? 1.2e-7 : 2.3e-16;

    // The two Krylov bases grow with the iteration (capacity doubles from max(128, 2k) steps): allocating them for the
    // step LIMIT would need (n + nrows) * min(n, nrows) values — 120 GB for the reference's own 10M x 2500 test shape —
    // when convergence takes ~100-300 steps.
    const size_t nr1 = (size_t)(nr > 0 ? nr : 1);
    int cap = (int)std::min<int64_t>(m, std::max(128, 2 * k + 20));
    DevBuf<T> Vb((size_t)(cap + 1) * n, st), Ub((size_t)cap * nr1, st), wu(nr1, st), wv((size_t)n, st);
    auto grow = [&](int used) {            // `used` steps (and used + 1 right vectors) are live
        const int ncap = (int)std::min<int64_t>(m, (int64_t)cap * 2);
        DevBuf<T> nV((size_t)(ncap + 1) * n, st), nU((size_t)ncap * nr1, st);
        SALG_CUDA(cudaMemcpyAsync(nV.get(), Vb.get(), (size_t)(used + 1) * n * sizeof(T), cudaMemcpyDeviceToDevice, st));
        SALG_CUDA(cudaMemcpyAsync(nU.get(), Ub.get(), (size_t)used * nr1 * sizeof(T), cudaMemcpyDeviceToDevice, st));
        Vb = std::move(nV);
        Ub = std::move(nU);
        cap = ncap;
    };
    DevBuf<double> coef((size_t)m + 1, st), alpha((size_t)m, st), beta((size_t)m, st), sq(1, st);
    SALG_CUDA(cudaMemsetAsync(alpha.get(), 0, (size_t)m * 8, st));
    SALG_CUDA(cudaMemsetAsync(beta.get(), 0, (size_t)m * 8, st));

    auto reorth = [&](const T* basis, int64_t len, int j, T* w, bool sharded) {
        if (j == 0 || len == 0) {
            if (sharded && j > 0) {   // keep the collective sequence identical on every rank
                for (int pass = 0; pass < 2; pass++) {
                    SALG_CUDA(cudaMemsetAsync(coef.get(), 0, (size_t)j * 8, st));
                    allreduce_f64(ctx, coef.get(), (size_t)j);
                }
            }
            return;
        }
        ProfScope ps(ctx, PROF_OTHER, 4.0 * (double)j * len * sizeof(T));
        for (int pass = 0; pass < 2; pass++) {
            SALG_CUDA(cudaMemsetAsync(coef.get(), 0, (size_t)j * 8, st));
            int gx = (int)std::min<int64_t>(ceil_div(len, 256 * 8), 64);
            basis_dots_kernel<T><<<dim3(gx, j), 256, 0, st>>>(basis, len, j, w, coef.get());
            ctx->n_launch++;
            if (sharded) allreduce_f64(ctx, coef.get(), (size_t)j);
            basis_update_kernel<T><<<(unsigned)ceil_div(len, 256), 256, 0, st>>>(basis, len, j, coef.get(), w);
            ctx->n_launch++;
        }
        SALG_CUDA(cudaGetLastError());
    };
    auto norm_store = [&](const T* w, int64_t len, bool sharded, const double* ref, T* dst, double* out_coef) {
        SALG_CUDA(cudaMemsetAsync(sq.get(), 0, 8, st));
        // (one f64 atomic per CTA on ONE address: same-address atomics serialise in L2 at ~25 clk each, so the grid stays at one CTA per SM)
        if (len) { sqnorm_kernel<T><<<(unsigned)std::min<int64_t>(ceil_div(len, 256 * 8), ctx->sm_count), 256, 0, st>>>(w, len, sq.get()); ctx->n_launch++; }
        if (sharded) allreduce_f64(ctx, sq.get(), 1);
        normalize_store_kernel<T><<<(unsigned)ceil_div(len > 0 ? len : 1, 256), 256, 0, st>>>(
            w, len, sq.get(), ref, 100.0 * eps_t, dst, out_coef);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
    };

    // start vector (replicated on every rank: same seed)
    {
        std::vector<double> h((size_t)n);
        host_normal(h, seed);
        std::vector<T> ht((size_t)n);
        for (int64_t i = 0; i < n; i++) ht[i] = (T)h[i];
        SALG_CUDA(cudaMemcpyAsync(wv.get(), ht.data(), (size_t)n * sizeof(T), cudaMemcpyHostToDevice, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        norm_store(wv.get(), n, false, nullptr, Vb.get(), coef.get() + m);
    }

    std::vector<double> h_alpha(m), h_beta(m), theta, resid;
    std::vector<double> Qfull;
    int j_done = 0, j_used = 0;
    bool converged = false;
    const int check_every = 20;
    for (int j = 0; j < m; j++) {
        if (j == cap) grow(j);
        // u_j
        spmv<T>(ctx, op->row_ptr, op->col, (const T*)op->val, nr, n, op->nnz, Vb.get() + (size_t)j * n, wu.get());
        reorth(Ub.get(), nr, j, wu.get(), ctx->nranks > 1);
        norm_store(wu.get(), nr, ctx->nranks > 1, j > 0 ? alpha.get() : nullptr, Ub.get() + (size_t)j * nr, alpha.get() + j);
        // v_{j+1}
        spmv<T>(ctx, op->t_ptr, op->t_idx, (const T*)op->t_val, n, nr, op->nnz, Ub.get() + (size_t)j * nr, wv.get());
        allreduce_T<T>(ctx, wv.get(), (size_t)n);
        reorth(Vb.get(), n, j + 1, wv.get(), false);
        norm_store(wv.get(), n, false, alpha.get(), Vb.get() + (size_t)(j + 1) * n, beta.get() + j);
        j_done = j + 1;

        bool last = (j_done == m);
        if (last || (j_done >= k + 10 && (j_done - (k + 10)) % check_every == 0)) {
            SALG_CUDA(cudaMemcpyAsync(h_alpha.data(), alpha.get(), (size_t)j_done * 8, cudaMemcpyDeviceToHost, st));
            SALG_CUDA(cudaMemcpyAsync(h_beta.data(), beta.get(), (size_t)j_done * 8, cudaMemcpyDeviceToHost, st));
            SALG_CUDA(cudaStreamSynchronize(st));
            int jj = j_done;
            bool broke = false;
            for (int i = 0; i < j_done; i++) {
                if (h_alpha[i] == 0.0) { jj = i; broke = true; break; }
                if (h_beta[i] == 0.0) { jj = i + 1; broke = true; break; }
            }
            if (jj == 0) { j_used = 0; break; }
            std::vector<double> d(jj), e(jj), Z(jj, 0.0);
            for (int i = 0; i < jj; i++) {
                d[i] = h_alpha[i] * h_alpha[i] + (i > 0 ? h_beta[i - 1] * h_beta[i - 1] : 0.0);
                e[i] = i + 1 < jj ? h_alpha[i] * h_beta[i] : 0.0;
            }
            Z[jj - 1] = 1.0;
            bool ok = tridiag_ql(d, e, jj, Z, 1);
            std::vector<int> ord(jj);
            for (int i = 0; i < jj; i++) ord[i] = i;
            std::sort(ord.begin(), ord.end(), [&](int a, int b) { return d[a] > d[b]; });
            int kk = std::min(k, jj);
            double coupling = h_alpha[jj - 1] * h_beta[jj - 1];
            double worst = 0.0;
            for (int i = 0; i < kk; i++) worst = std::max(worst, coupling * std::fabs(Z[ord[i]]));
            bool conv = ok && jj >= kk && worst <= tol * std::fabs(d[ord[0]]);
            if (conv || broke || last) {
                converged = conv || broke || (jj == std::min<int64_t>(n, nr_total));
                j_used = jj;
                break;
            }
        }
    }
    if (steps_out) *steps_out = j_used;
    s_out.clear();
    SALG_CUDA(cudaMemsetAsync(d_Vpanel, 0, (size_t)n * LP * sizeof(T), st));
    if (j_used == 0) return 0;
    // final Ritz extraction with eigenvectors
    {
        int jj = j_used;
        std::vector<double> d(jj), e(jj), Z((size_t)jj * jj, 0.0);
        for (int i = 0; i < jj; i++) {
            d[i] = h_alpha[i] * h_alpha[i] + (i > 0 ? h_beta[i - 1] * h_beta[i - 1] : 0.0);
            e[i] = i + 1 < jj ? h_alpha[i] * h_beta[i] : 0.0;
            Z[(size_t)i * jj + i] = 1.0;
        }
        bool ok = tridiag_ql(d, e, jj, Z, jj);
        SALG_REQUIRE(ok, SALG_ERR_NUMERIC, "SVD computation failed: tridiagonal QL did not converge");
        std::vector<int> ord(jj);
        for (int i = 0; i < jj; i++) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](int a, int b) { return d[a] > d[b]; });
        int kk = std::min(k, jj);
        std::vector<double> Q((size_t)jj * LP, 0.0);
        for (int i = 0; i < kk; i++) {
            double th = d[ord[i]];
            s_out.push_back(std::sqrt(th > 0.0 ? th : 0.0));
            for (int t = 0; t < jj; t++) Q[(size_t)t * LP + i] = Z[(size_t)t * jj + ord[i]];
        }
        DevBuf<double> dQ((size_t)jj * LP, st);
        SALG_CUDA(cudaMemcpyAsync(dQ.get(), Q.data(), Q.size() * 8, cudaMemcpyHostToDevice, st));
        ritz_vectors_kernel<T><<<(unsigned)ceil_div(n, 4), 256, 0, st>>>(Vb.get(), n, jj, dQ.get(), d_Vpanel);
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        SALG_CUDA(cudaStreamSynchronize(st));
    }
    return converged ? (int)s_out.size() : -(int)s_out.size();
}
template int lanczos_svd<float>(salg_ctx*, const salg_csr*, int, int, uint64_t, double, float*, std::vector<double>&, int*);
template int lanczos_svd<double>(salg_ctx*, const salg_csr*, int, int, uint64_t, double, double*, std::vector<double>&, int*);

}  // namespace salg
