// dense.cu — tall-skinny panel kernels around the sparse products: Gram + Cholesky + triangular apply
// (CholeskyQR2, replacing nalgebra's qr()/lu() power-iteration normalisers inside single-svdlib's
// randomized_svd), the one-CTA Jacobi SVD of the projected l x l factor (K7), svd_flip (K8,
// single-svdlib randomized::svd_flip called at pca/sparse/mod.rs:203) and layout helpers.
// Every panel is (rows x 64) row-major (LP = 64 >= l = n_components + n_oversamples, zero padded);
// every small matrix is 64 x 64 row-major f64 with the leading k x k block meaningful.
#include "common.cuh"

namespace salg {

// ---- Gram: G = P^T P (f64 accumulation), cs = 1^T P ------------------------------------------------------
// out[0..4096) = G row-major, out[4096..4160) = column sums.  out must be zeroed by the caller.
template <typename T>
__global__ void __launch_bounds__(256)
panel_gram_kernel(const T* __restrict__ P, int64_t m, double* __restrict__ out) {
    constexpr int TR = 32;                       // rows per tile
    __shared__ __align__(16) T tile[TR][LP + 4];   // row stride 68 elements: 16 B aligned rows, conflict-free float4 reads
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tj = tid & 15;      // 16 x 16 threads, 4 x 4 outputs each
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
    double cs = 0.0;
    const int64_t n_tiles = (m + TR - 1) / TR;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t r0 = t * TR;
        __syncthreads();
        for (int i = tid; i < TR * LP; i += 256) {
            int r = i >> 6, c = i & 63;
            tile[r][c] = (r0 + r < m) ? P[(r0 + r) * LP + c] : T(0);
        }
        __syncthreads();
        if (sizeof(T) == 4) {
            // f32 panels: 32-row partial products in f32 (relative error <= 32 * 2^-24 per partial), summed in f64
            float p[4][4];
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) p[x][y] = 0.f;
#pragma unroll 8
            for (int r = 0; r < TR; r++) {
                const float4 av = *reinterpret_cast<const float4*>(&tile[r][4 * ti]);
                const float4 bv = *reinterpret_cast<const float4*>(&tile[r][4 * tj]);
                const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) p[x][y] = fmaf(a[x], b[y], p[x][y]);
            }
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) acc[x][y] += (double)p[x][y];
        } else {
#pragma unroll 4
            for (int r = 0; r < TR; r++) {
                double a[4], b[4];
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    a[x] = (double)tile[r][4 * ti + x];
                    b[x] = (double)tile[r][4 * tj + x];
                }
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) acc[x][y] = fma(a[x], b[y], acc[x][y]);
            }
        }
        if (tid < LP) {
            for (int r = 0; r < TR; r++) cs += (double)tile[r][tid];
        }
    }
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) atomicAdd(&out[(4 * ti + x) * LP + 4 * tj + y], acc[x][y]);
    if (tid < LP) atomicAdd(&out[LP * LP + tid], cs);
}

template <typename T>
void panel_gram(salg_ctx* ctx, const T* P, int64_t m, double* d_out) {
    SALG_CUDA(cudaMemsetAsync(d_out, 0, GRAM_BUF * sizeof(double), ctx->stream));
    if (m == 0) return;
    ProfScope ps(ctx, PROF_GRAM, (double)m * 60 * sizeof(T));
    int64_t want = ceil_div(m, 32);
    int64_t cap = (int64_t)ctx->sm_count * 4;
    panel_gram_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(P, m, d_out);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void panel_gram<float>(salg_ctx*, const float*, int64_t, double*);
template void panel_gram<double>(salg_ctx*, const double*, int64_t, double*);

// ---- Cholesky G = L L^T of the leading k x k block, R = L^T and R^{-1}; one CTA of 512 threads ----------------------
// This kernel is replicated on every GPU of a row-sharded run and runs ~25 times per fit, so its latency is serial
// time at any GPU count (the first version, one thread per row with the row in registers, took 92 us per call).
// Here the matrix lives in shared memory and every elimination step is ONE block-wide rank-1 update behind ONE
// barrier: step j updates the trailing block of A (columns > j) and, in the same sweep, the forward substitution of
// the identity (columns <= j of B), so L^{-1} falls out of the same 64 steps.  Column j of A is left UNSCALED
// (readers multiply by 1/l_jj), which is what removes the second barrier of the textbook loop.
// Thread (g, c): column c, rows g, g + 8, ..., g + 56.  Pivots that are not safely positive are floored
// (rank-deficient panels: l > rank(A)); flag bit 1 is raised, the caller's second CholeskyQR pass re-orthonormalises.
constexpr int CHOL_THREADS = 512;
constexpr int CHOL_LD = LP + 1;
constexpr size_t CHOL_SMEM = (size_t)(2 * LP * CHOL_LD + 2 * LP + 2) * sizeof(double);

template <typename T>
__global__ void __launch_bounds__(CHOL_THREADS)
chol_inv_kernel(const double* __restrict__ G, int k, double* __restrict__ R, double* __restrict__ Rinv,
                T* __restrict__ RinvT, int* __restrict__ flag) {
    extern __shared__ double chol_sm[];
    double* A = chol_sm;                       // [64][65] lower triangle: trailing matrix, later unscaled columns of L
    double* B = A + LP * CHOL_LD;              // [64][65] forward-substituted identity, later unscaled rows of L^{-1}
    double* dinv = B + LP * CHOL_LD;           // 1 / l_jj
    double* ldiag = dinv + LP;                 // l_jj
    double* s_md = ldiag + LP;                 // [2]
    const int tid = threadIdx.x;
    const int c = tid & 63, g = tid >> 6;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int t = g + 8 * m;
        A[t * CHOL_LD + c] = (t < k && c < k) ? ((c <= t) ? G[t * LP + c] : 0.0) : ((t == c) ? 1.0 : 0.0);
        B[t * CHOL_LD + c] = (t == c) ? 1.0 : 0.0;
    }
    if (tid < 64) {
        double md = (tid < k) ? G[tid * LP + tid] : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) md = fmax(md, __shfl_xor_sync(0xFFFFFFFFu, md, o));
        if ((tid & 31) == 0) s_md[tid >> 5] = md;
    }
    __syncthreads();
    const double mdiag = fmax(s_md[0], s_md[1]);
    const double floor_piv = (mdiag > 0.0 ? mdiag : 1.0) * 1e-13;
    bool bad = false;
    for (int j = 0; j < k; j++) {
        double p = A[j * CHOL_LD + j];
        if (!(p > floor_piv)) {
            p = floor_piv;
            bad = true;
        }
        const double di = rsqrt(p);
        if (tid == 0) {
            dinv[j] = di;
            ldiag[j] = p * di;
        }
        // the row-j factors this thread needs: l_cj (trailing update) or (L^{-1})_jc (substitution), both scaled once
        const double rowj = ((c > j) ? A[c * CHOL_LD + j] : B[j * CHOL_LD + c]) * di;
        double* M = (c > j) ? A : B;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int t = g + 8 * m;
            if (t > j && (c <= j || c <= t)) {
                const double ltj = A[t * CHOL_LD + j] * di;
                M[t * CHOL_LD + c] = fma(-ltj, rowj, M[t * CHOL_LD + c]);
            }
        }
        __syncthreads();
    }
    if (bad && tid == 0) atomicOr(flag, 1);
    if (tid < 64 && tid >= k) {
        dinv[tid] = 1.0;
        ldiag[tid] = 1.0;
    }
    __syncthreads();
    // outputs (row-major 64 x 64): R[r][i] = L[i][r], Rinv[r][i] = (L^{-1})[i][r]; thread (g, c) writes rows r = g + 8m,
    // column i = c (consecutive threads -> consecutive addresses)
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int r = g + 8 * m, i = c;
        const double x = (r <= i) ? B[i * CHOL_LD + r] * dinv[i] : 0.0;
        if (Rinv) Rinv[r * LP + i] = x;
        if (RinvT) RinvT[r * LP + i] = (T)x;
        if (R) R[r * LP + i] = (r < i) ? A[i * CHOL_LD + r] * dinv[r] : ((r == i) ? ldiag[r] : 0.0);
    }
}

template <typename T>
void chol_inv(salg_ctx* ctx, const double* d_G, int k, double* d_R, double* d_Rinv, T* d_RinvT, int* d_flag) {
    ProfScope ps(ctx, PROF_CHOL, 0.0);
    set_max_dyn_smem(chol_inv_kernel<T>, (int)((int)CHOL_SMEM));
    chol_inv_kernel<T><<<1, CHOL_THREADS, CHOL_SMEM, ctx->stream>>>(d_G, k, d_R, d_Rinv, d_RinvT, d_flag);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void chol_inv<float>(salg_ctx*, const double*, int, double*, double*, float*, int*);
template void chol_inv<double>(salg_ctx*, const double*, int, double*, double*, double*, int*);

// ---- out = P * M (M 64 x 64 row-major, T); in place allowed ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
panel_mul_kernel(const T* P, int64_t m, const T* __restrict__ M, T* out) {
    constexpr int TR = 64;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    T (*Ms)[LP] = reinterpret_cast<T (*)[LP]>(dyn_smem);
    T (*Pt)[TR + 4] = reinterpret_cast<T (*)[TR + 4]>(dyn_smem + sizeof(T) * LP * LP);   // transposed tile: Pt[col][row]
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;   // rows 4ty..4ty+3, cols 4tx..4tx+3
    for (int i = tid; i < LP * LP; i += 256) Ms[i >> 6][i & 63] = M[i];
    const int64_t n_tiles = (m + TR - 1) / TR;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t r0 = t * TR;
        __syncthreads();
        for (int i = tid; i < TR * LP; i += 256) {
            int r = i >> 6, c = i & 63;
            Pt[c][r] = (r0 + r < m) ? P[(r0 + r) * LP + c] : T(0);
        }
        __syncthreads();
        T acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = T(0);
#pragma unroll 8
        for (int kk = 0; kk < LP; kk++) {
            T a[4], b[4];
            if (sizeof(T) == 4) {
                const float4 av = *reinterpret_cast<const float4*>(&Pt[kk][4 * ty]);
                const float4 bv = *reinterpret_cast<const float4*>(&Ms[kk][4 * tx]);
                a[0] = (T)av.x; a[1] = (T)av.y; a[2] = (T)av.z; a[3] = (T)av.w;
                b[0] = (T)bv.x; b[1] = (T)bv.y; b[2] = (T)bv.z; b[3] = (T)bv.w;
            } else {
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    a[x] = Pt[kk][4 * ty + x];
                    b[x] = Ms[kk][4 * tx + x];
                }
            }
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
#pragma unroll
        for (int x = 0; x < 4; x++) {
            int64_t r = r0 + 4 * ty + x;
            if (r < m) {
#pragma unroll
                for (int y = 0; y < 4; y++) out[r * LP + 4 * tx + y] = acc[x][y];
            }
        }
    }
}

template <typename T>
void panel_mul(salg_ctx* ctx, const T* P, int64_t m, const T* d_M, T* out) {
    if (m == 0) return;
    ProfScope ps(ctx, PROF_PANELMUL, 2.0 * (double)m * 60 * sizeof(T));
    int64_t want = ceil_div(m, 64);
    int64_t cap = (int64_t)ctx->sm_count * 4;
    constexpr int kSmem = (int)sizeof(T) * (LP * LP + LP * (64 + 4));
    set_max_dyn_smem(panel_mul_kernel<T>, (int)(kSmem));
    panel_mul_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, kSmem, ctx->stream>>>(P, m, d_M, out);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void panel_mul<float>(salg_ctx*, const float*, int64_t, const float*, float*);
template void panel_mul<double>(salg_ctx*, const double*, int64_t, const double*, double*);

// ---- 64 x 64 f64 helpers ------------------------------------------------------------------------------------
__global__ void mat64_mul_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double (*As)[LP + 1] = reinterpret_cast<double (*)[LP + 1]>(dyn_smem);
    double (*Bs)[LP + 1] = As + LP;
    for (int i = threadIdx.x; i < LP * LP; i += blockDim.x) {
        As[i >> 6][i & 63] = A[i];
        Bs[i >> 6][i & 63] = B[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LP * LP; i += blockDim.x) {
        int r = i >> 6, c = i & 63;
        double s = 0.0;
        for (int t = 0; t < LP; t++) s = fma(As[r][t], Bs[t][c], s);
        C[i] = s;
    }
}
void mat64_mul(salg_ctx* ctx, const double* A, const double* B, double* C) {
    constexpr int kSmem = 2 * LP * (LP + 1) * 8;
    set_max_dyn_smem(mat64_mul_kernel, (int)(kSmem));
    mat64_mul_kernel<<<1, 256, kSmem, ctx->stream>>>(A, B, C);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

// v_out(64) = v_in(64) * M(64x64)
__global__ void vec64_mat_kernel(const double* __restrict__ v, const double* __restrict__ M, double* __restrict__ o) {
    __shared__ double vs[LP];
    if (threadIdx.x < LP) vs[threadIdx.x] = v[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < LP) {
        double s = 0.0;
        for (int t = 0; t < LP; t++) s = fma(vs[t], M[t * LP + threadIdx.x], s);
        o[threadIdx.x] = s;
    }
}
void vec64_mat(salg_ctx* ctx, const double* v, const double* M, double* o) {
    vec64_mat_kernel<<<1, 64, 0, ctx->stream>>>(v, M, o);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

template <typename T>
__global__ void cast_mat64_kernel(const double* __restrict__ s, T* __restrict__ d, const double* __restrict__ colscale) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < LP * LP) d[i] = (T)(colscale ? s[i] * colscale[i & 63] : s[i]);
}
template <typename T>
void cast_mat64(salg_ctx* ctx, const double* src, T* dst, const double* colscale) {
    cast_mat64_kernel<T><<<16, 256, 0, ctx->stream>>>(src, dst, colscale);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void cast_mat64<float>(salg_ctx*, const double*, float*, const double*);
template void cast_mat64<double>(salg_ctx*, const double*, double*, const double*);

// ---- one-sided Jacobi SVD of a k x k matrix, one CTA of 32 warps ---------------------------------------------
// A = U diag(S) V^T, S descending.  Rows of W hold the columns of A (so a column rotation touches two
// contiguous shared-memory rows); one warp per pair, round-robin tournament ordering.
__global__ void __launch_bounds__(1024)
jacobi_svd64_kernel(const double* __restrict__ A, int k, double* __restrict__ U, double* __restrict__ S,
                    double* __restrict__ V, int* __restrict__ flag) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double (*W)[LP + 1] = reinterpret_cast<double (*)[LP + 1]>(dyn_smem);
    double (*Vt)[LP + 1] = W + LP;
    __shared__ double sig[LP];
    __shared__ int order[LP];
    __shared__ int s_rot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < LP * LP; i += 1024) {
        int r = i >> 6, c = i & 63;
        W[c][r] = (r < k && c < k) ? A[r * LP + c] : 0.0;
        Vt[r][c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int ke = (k + 1) & ~1;        // even number of players (a zero column pads odd k)
    const int n_pairs = ke / 2;
    const double tol = 1.5e-14;   // ~ k * eps: below the rounding noise of the k-term dot products
    int sweep = 0;
    for (; sweep < 30; sweep++) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int round = 0; round < ke - 1; round++) {
            if (warp < n_pairs) {
                int p, q;
                if (warp == 0) {
                    p = round % (ke - 1);
                    q = ke - 1;
                } else {
                    p = (round + warp) % (ke - 1);
                    q = (round - warp + (ke - 1)) % (ke - 1);
                }
                if (p > q) { int t = p; p = q; q = t; }
                double wp0 = W[p][lane], wp1 = W[p][lane + 32];
                double wq0 = W[q][lane], wq1 = W[q][lane + 32];
                double alpha = wp0 * wp0 + wp1 * wp1;
                double beta = wq0 * wq0 + wq1 * wq1;
                double gamma = wp0 * wq0 + wp1 * wq1;
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    alpha += __shfl_xor_sync(0xFFFFFFFFu, alpha, o);
                    beta += __shfl_xor_sync(0xFFFFFFFFu, beta, o);
                    gamma += __shfl_xor_sync(0xFFFFFFFFu, gamma, o);
                }
                if (fabs(gamma) > tol * sqrt(alpha * beta) && gamma != 0.0) {
                    // rotation from fast reciprocal / reciprocal-square-root intrinsics: the 59 rounds x ~8 sweeps are a
                    // serial chain, divisions and square roots were most of its latency
                    double zeta = (beta - alpha) * __drcp_rn(2.0 * gamma);
                    double h2 = fma(zeta, zeta, 1.0);
                    double w = h2 * rsqrt(h2);                                   // sqrt(1 + zeta^2)
                    double t = (zeta >= 0.0 ? 1.0 : -1.0) * __drcp_rn(fabs(zeta) + w);
                    double c = rsqrt(fma(t, t, 1.0)), s = c * t;
                    W[p][lane] = c * wp0 - s * wq0;
                    W[p][lane + 32] = c * wp1 - s * wq1;
                    W[q][lane] = s * wp0 + c * wq0;
                    W[q][lane + 32] = s * wp1 + c * wq1;
                    double vp0 = Vt[p][lane], vp1 = Vt[p][lane + 32];
                    double vq0 = Vt[q][lane], vq1 = Vt[q][lane + 32];
                    Vt[p][lane] = c * vp0 - s * vq0;
                    Vt[p][lane + 32] = c * vp1 - s * vq1;
                    Vt[q][lane] = s * vp0 + c * vq0;
                    Vt[q][lane + 32] = s * vp1 + c * vq1;
                    if (lane == 0) s_rot = 1;
                }
            }
            __syncthreads();
        }
        int any = s_rot;
        __syncthreads();
        if (!any) break;
    }
    if (tid == 0 && sweep >= 30) atomicOr(flag, 2);
    // singular values = row norms of W
    if (warp < 2) {
        int p = tid;   // 0..63
        double s = 0.0;
        for (int i = 0; i < LP; i++) s += W[p][i] * W[p][i];
        sig[p] = (p < k) ? sqrt(s) : -1.0;
    }
    __syncthreads();
    if (tid < LP) {
        // rank by descending sigma (stable on index)
        int rank = 0;
        double me = sig[tid];
        for (int j = 0; j < LP; j++) {
            double o = sig[j];
            rank += (o > me) || (o == me && j < tid);
        }
        order[rank] = tid;
    }
    __syncthreads();
    for (int i = tid; i < LP * LP; i += 1024) {
        int r = i >> 6, c = i & 63;    // output element [r][c]; column c is the c-th largest triplet
        int src = order[c];
        double sg = sig[src];
        double u = 0.0, v = 0.0;
        if (c < k && r < k) {
            u = sg > 0.0 ? W[src][r] / sg : (r == c ? 1.0 : 0.0);
            v = Vt[src][r];
        } else if (r == c) {
            u = 1.0;
            v = 1.0;
        }
        U[i] = u;
        V[i] = v;
    }
    if (tid < LP) S[tid] = tid < k ? sig[order[tid]] : 0.0;
}

void jacobi_svd64(salg_ctx* ctx, const double* d_A, int k, double* d_U, double* d_S, double* d_V, int* d_flag) {
    ProfScope ps(ctx, PROF_JACOBI, 0.0);
    constexpr int kSmem = 2 * LP * (LP + 1) * 8;
    set_max_dyn_smem(jacobi_svd64_kernel, (int)(kSmem));
    jacobi_svd64_kernel<<<1, 1024, kSmem, ctx->stream>>>(d_A, k, d_U, d_S, d_V, d_flag);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

// ---- out[j] = sum_r w[r] * P[r][j]  (w == nullptr -> 1); f64, out zeroed here ---------------------------------
template <typename T>
__global__ void panel_colsum_kernel(const T* __restrict__ P, int64_t m, const T* __restrict__ w,
                                    double* __restrict__ out) {
    __shared__ double part[4][LP];
    const int c = threadIdx.x & 63, g = threadIdx.x >> 6;   // 256 threads: 4 row groups x 64 columns
    double s = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * 4 + g; r < m; r += (int64_t)gridDim.x * 4) {
        double x = (double)P[r * LP + c];
        s += w ? (double)w[r] * x : x;
    }
    part[g][c] = s;
    __syncthreads();
    if (g == 0) atomicAdd(&out[c], part[0][c] + part[1][c] + part[2][c] + part[3][c]);
}
template <typename T>
void panel_colsum(salg_ctx* ctx, const T* P, int64_t m, const T* w, double* d_out64) {
    SALG_CUDA(cudaMemsetAsync(d_out64, 0, LP * sizeof(double), ctx->stream));
    if (m == 0) return;
    int64_t want = ceil_div(m, 64);
    int64_t cap = (int64_t)ctx->sm_count * 4;
    panel_colsum_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(P, m, w, d_out64);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void panel_colsum<float>(salg_ctx*, const float*, int64_t, const float*, double*);
template void panel_colsum<double>(salg_ctx*, const double*, int64_t, const double*, double*);

// ---- svd_flip(u, vt, u_based_decision = false): sign[i] = sgn(V[argmax_j |V[j][i]|][i]) -----------------------
template <typename T>
__global__ void flip_find_kernel(const T* __restrict__ V, int64_t n, double* __restrict__ sign) {
    // one CTA per component; first maximum wins ties (argmax semantics)
    __shared__ double s_abs[256];
    __shared__ long long s_idx[256];
    const int comp = blockIdx.x;
    double best = -1.0;
    long long bi = 0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        double a = fabs((double)V[j * LP + comp]);
        if (a > best) { best = a; bi = j; }
    }
    s_abs[threadIdx.x] = best;
    s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            double a = s_abs[threadIdx.x + o];
            long long i2 = s_idx[threadIdx.x + o];
            if (a > s_abs[threadIdx.x] || (a == s_abs[threadIdx.x] && i2 < s_idx[threadIdx.x])) {
                s_abs[threadIdx.x] = a;
                s_idx[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double v = n > 0 ? (double)V[s_idx[0] * LP + comp] : 1.0;
        sign[comp] = v < 0.0 ? -1.0 : 1.0;
    }
}

template <typename T>
__global__ void panel_colscale_kernel(T* __restrict__ P, int64_t m, const double* __restrict__ scale) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m * LP; i += stride) P[i] = (T)((double)P[i] * scale[i & 63]);
}

template <typename T>
void flip_find(salg_ctx* ctx, const T* V, int64_t n_eff, double* d_sign64) {
    flip_find_kernel<T><<<LP, 256, 0, ctx->stream>>>(V, n_eff, d_sign64);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void flip_find<float>(salg_ctx*, const float*, int64_t, double*);
template void flip_find<double>(salg_ctx*, const double*, int64_t, double*);

template <typename T>
void panel_colscale(salg_ctx* ctx, T* P, int64_t m, const double* d_scale64) {
    if (m == 0) return;
    int64_t want = ceil_div(m * LP, 256);
    int64_t cap = (int64_t)ctx->sm_count * 16;
    panel_colscale_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(P, m, d_scale64);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void panel_colscale<float>(salg_ctx*, float*, int64_t, const double*);
template void panel_colscale<double>(salg_ctx*, double*, int64_t, const double*);

// ---- layout helpers ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void panel_to_rowmajor_t_kernel(const T* __restrict__ V, int64_t n, int d, T* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over n * d, component fastest on the read side
    if (i >= n * d) return;
    int64_t j = i / d;
    int c = (int)(i % d);
    out[(int64_t)c * n + j] = V[j * LP + c];
}
template <typename T>
void panel_to_rowmajor_t(salg_ctx* ctx, const T* V, int64_t n, int d, T* out) {
    if (n * d == 0) return;
    panel_to_rowmajor_t_kernel<T><<<(unsigned)ceil_div(n * d, 256), 256, 0, ctx->stream>>>(V, n, d, out);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void panel_to_rowmajor_t<float>(salg_ctx*, const float*, int64_t, int, float*);
template void panel_to_rowmajor_t<double>(salg_ctx*, const double*, int64_t, int, double*);

template <typename T>
__global__ void panel_pack_kernel(const T* __restrict__ src, int64_t m, int k, T* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m * LP; i += stride) {
        int64_t r = i >> 6;
        int c = (int)(i & 63);
        dst[i] = c < k ? src[r * k + c] : T(0);
    }
}
template <typename T>
void panel_pack(salg_ctx* ctx, const T* src, int64_t m, int k, T* dst) {
    if (m == 0) return;
    int64_t want = ceil_div(m * LP, 256);
    int64_t cap = (int64_t)ctx->sm_count * 16;
    panel_pack_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(src, m, k, dst);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void panel_pack<float>(salg_ctx*, const float*, int64_t, int, float*);
template void panel_pack<double>(salg_ctx*, const double*, int64_t, int, double*);

template <typename T>
__global__ void panel_unpack_kernel(const T* __restrict__ src, int64_t m, int k, T* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m * k; i += stride) {
        int64_t r = i / k;
        int c = (int)(i % k);
        dst[i] = src[r * LP + c];
    }
}
template <typename T>
void panel_unpack(salg_ctx* ctx, const T* src, int64_t m, int k, T* dst) {
    if (m * k == 0) return;
    int64_t want = ceil_div(m * k, 256);
    int64_t cap = (int64_t)ctx->sm_count * 16;
    panel_unpack_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, ctx->stream>>>(src, m, k, dst);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void panel_unpack<float>(salg_ctx*, const float*, int64_t, int, float*);
template void panel_unpack<double>(salg_ctx*, const double*, int64_t, int, double*);

// ---- CholeskyQR2 --------------------------------------------------------------------------------------------------
// Y (m_local x 64, leading k columns meaningful) <- Q with Q^T Q = I over ALL ranks' rows when `sharded`
// (Gram all-reduced), else over the local rows (replicated small-side panels).  Optional outputs:
// d_colsum64 = 1^T Q (global), d_Rtot = R2 R1 (64 x 64 f64) with Y_in = Q Rtot.
template <typename T>
void cholqr2(salg_ctx* ctx, T* Y, int64_t m_local, int k, bool sharded, double* d_colsum64, double* d_Rtot,
             int* d_flag, int passes) {
    cudaStream_t st = ctx->stream;
    DevBuf<double> G(GRAM_BUF, st), R1(LP * LP, st), Ri1(LP * LP, st), R2(LP * LP, st), Ri2(LP * LP, st), cs(LP, st);
    DevBuf<T> RiT(LP * LP, st);
    for (int pass = 0; pass < passes; pass++) {
        panel_gram<T>(ctx, Y, m_local, G.get());
        if (sharded) allreduce_f64(ctx, G.get(), GRAM_BUF);
        double* R = pass == 0 ? R1.get() : R2.get();
        double* Ri = pass == 0 ? Ri1.get() : Ri2.get();
        chol_inv<T>(ctx, G.get(), k, R, Ri, RiT.get(), d_flag);
        panel_mul<T>(ctx, Y, m_local, RiT.get(), Y);
        if (d_colsum64) {
            if (pass == 0) vec64_mat(ctx, G.get() + LP * LP, Ri, passes == 1 ? d_colsum64 : cs.get());
            else vec64_mat(ctx, cs.get(), Ri, d_colsum64);
        }
    }
    if (d_Rtot) {
        if (passes == 1) SALG_CUDA(cudaMemcpyAsync(d_Rtot, R1.get(), LP * LP * 8, cudaMemcpyDeviceToDevice, st));
        else mat64_mul(ctx, R2.get(), R1.get(), d_Rtot);
    }
}
template void cholqr2<float>(salg_ctx*, float*, int64_t, int, bool, double*, double*, int*, int);
template void cholqr2<double>(salg_ctx*, double*, int64_t, int, bool, double*, double*, int*, int);

}  // namespace salg
