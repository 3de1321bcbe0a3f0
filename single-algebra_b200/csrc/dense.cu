// dense.cu — tall-skinny panel kernels around the sparse products: Gram + Cholesky + triangular apply
// (CholeskyQR2, replacing nalgebra's qr()/lu() power-iteration normalisers inside single-svdlib's
// randomized_svd), the one-CTA Jacobi SVD of the projected l x l factor (K7), svd_flip (K8,
// single-svdlib randomized::svd_flip called at pca/sparse/mod.rs:203) and layout helpers.
// Every panel is (rows x 64) row-major (LP = 64 >= l = n_components + n_oversamples, zero padded);
// every small matrix is 64 x 64 row-major f64 with the leading k x k block meaningful.
#include <algorithm>

#include "common.cuh"

namespace salg {

// power-of-two scale that puts `amax` just below 2^14 (same rule as tcgen05.cuh: tc_pow2_scale)
__device__ __forceinline__ float tc_pow2_scale_dev(float amax) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int e;
    frexpf(amax, &e);
    return ldexpf(1.f, 14 - e);
}

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
// This kernel is replicated on every GPU of a row-sharded run and runs ~18 times per fit, so its latency is serial
// time at any GPU count (first version, one thread per row: 92 us per call; second, one block-wide rank-1 update and one
// barrier per column: 35 us).  This version is BLOCKED by 8 columns: per block (1) one warp factors the 8 x 8 diagonal
// block in registers (the pivot chain — shuffle, rsqrt, 7 fused multiply-adds per column — is the only serial part left)
// and inverts it, (2) 8-term triangular products give the panel below and the block row of L^{-1} (forward substitution of
// the identity, carried along), (3) ONE block-wide rank-8 update of the trailing matrix and of the rows of L^{-1} below.
// 8 x 3 barriers instead of 64.  The matrix is padded with the identity to 64 x 64, so every block is full.
// Pivots that are not safely positive (below 1e-13 of the largest diagonal entry; rank-deficient panels: l > rank(A)) raise flag
// bit 1 and are floored or dropped, see below.
constexpr int CHOL_THREADS = 512;
constexpr int CHOL_LD = LP + 1;
constexpr int CHOL_NB = 8;
constexpr size_t CHOL_SMEM = (size_t)(2 * LP * CHOL_LD + 2 * CHOL_NB * (CHOL_NB + 1) + 4 + LP) * sizeof(double);

// The factorisation proper, for CHOL_THREADS threads of one CTA: on return (after its last barrier) chol_sm holds L in the
// lower triangle of A = chol_sm[0 .. 64*65) and L^{-1} in B = chol_sm + 64*65 (both ld CHOL_LD).  G: k x k, leading
// dimension ldg, global or shared.  Returns THIS thread's flags (warp 0 only): bit 0 a pivot was floored, bit 1 a pivot below
// warn_rel x the largest diagonal entry.  Not inlined: the fused small-side kernel calls it from three places, and three copies
// of this unrolled body measured +10 us per launch (instruction-cache misses on a one-CTA latency chain).
__device__ __noinline__ int chol_inv_block(const double* G, int ldg, int k, double* chol_sm, double warn_rel = 0.0, int drop = 0) {
    double* A = chol_sm;                       // [64][65] lower triangle: trailing matrix, finished columns hold L
    double* B = A + LP * CHOL_LD;              // [64][65] forward-substituted identity, finished rows hold L^{-1}
    double* L8 = B + LP * CHOL_LD;             // [8][9] diagonal block of L
    double* Li8 = L8 + CHOL_NB * (CHOL_NB + 1);   // [8][9] its inverse
    double* s_md = Li8 + CHOL_NB * (CHOL_NB + 1); // [2]
    int* s_fmask = reinterpret_cast<int*>(s_md + 2);   // dependent columns of the current block
    double* s_diag = s_md + 4;                         // [64] the matrix's own diagonal (squared column norms of the panel)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = tid & 63, g = tid >> 6;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int t = g + 8 * m;
        A[t * CHOL_LD + c] = (t < k && c < k) ? ((c <= t) ? G[t * ldg + c] : 0.0) : ((t == c) ? 1.0 : 0.0);
        B[t * CHOL_LD + c] = (t == c) ? 1.0 : 0.0;
    }
    if (tid < 64) {
        double md = (tid < k) ? G[tid * ldg + tid] : 0.0;
        s_diag[tid] = (tid < k) ? md : 1.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) md = fmax(md, __shfl_xor_sync(0xFFFFFFFFu, md, o));
        if ((tid & 31) == 0) s_md[tid >> 5] = md;
    }
    __syncthreads();
    const double mdiag = fmax(s_md[0], s_md[1]);
    const double floor_piv = (mdiag > 0.0 ? mdiag : 1.0) * 1e-13;
    int bad = 0;                               // bit 0: a pivot was floored, bit 1: a pivot below warn_rel * (largest diagonal entry)
    for (int j0 = 0; j0 < LP; j0 += CHOL_NB) {
        // ---- (1) diagonal block: lane r (mod 8) holds row r of the block
        if (warp == 0) {
            const int r = lane & 7;
            int fmask = 0;
            double a[CHOL_NB], di[CHOL_NB];
#pragma unroll
            for (int cc = 0; cc < CHOL_NB; cc++) a[cc] = A[(j0 + r) * CHOL_LD + j0 + cc];
#pragma unroll
            for (int j = 0; j < CHOL_NB; j++) {
                double p = __shfl_sync(0xFFFFFFFFu, a[j], j);
                if (warn_rel > 0.0 && !(p > warn_rel * s_diag[j0 + j])) bad |= 2;   // (scale-free: sin^2 of the column's angle to the span before it)
                // A pivot that is not safely positive marks a column that (numerically) depends on the ones before it.  Its
                // sub-diagonal column of L is set to ZERO (no trailing update from it) and
                //   drop = 0: the pivot is floored: Q = Y R^{-1} gets the residual divided by sqrt(floor) there — amplified
                //             rounding noise that a following pass re-orthonormalises when it is independent (ill-conditioned
                //             f32 panels: their true pivots sit at the same 1e-14 as the noise);
                //   drop = 1: the column is dropped altogether (zero column of L, zero row of L^{-1}: zero column of Q, zero
                //             row of R).  The LAST pass of a CholeskyQR uses this: what is still dependent after a pass that
                //             amplified every independent direction to O(1) is an exact dependency (rank(A) < l).
                // Rounds 1-2 floored the pivot and kept the sub-diagonal column: with several dependent columns the trailing
                // updates are of the size of the floor itself and grow geometrically (18 dependent columns of a rank-11 f32
                // sketch: inf / NaN), and never-orthogonal noise columns of Q inflate the singular values of Q^T A.
                // drop pass: scale-free test against the column's own squared norm (sin^2 of its angle to the span of the
                // columns before it below 1e-5): the pass before may have left a noise column of any size, and the Gram may
                // come from the tensor-core pass (fp16 two-term operands: entries good to ~1e-6 of the norms)
                const bool dep = drop ? !(p > 1e-5 * s_diag[j0 + j]) : !(p > floor_piv);
                if (dep) {
                    bad |= 1;
                    fmask |= 1 << j;
                    p = drop ? 0.0 : floor_piv;
                }
                const double d = (dep && drop) ? 0.0 : rsqrt(p);
                di[j] = d;
                a[j] = (r == j) ? p * d : (dep ? 0.0 : a[j] * d);          // column j of L (rows >= j)
#pragma unroll
                for (int cc = j + 1; cc < CHOL_NB; cc++) {
                    const double lc = __shfl_sync(0xFFFFFFFFu, a[j], cc);
                    if (r >= cc) a[cc] = fma(-a[j], lc, a[cc]);
                }
            }
            if (lane == 0) *s_fmask = fmask;
            if (lane < CHOL_NB) {
#pragma unroll
                for (int cc = 0; cc < CHOL_NB; cc++) L8[r * (CHOL_NB + 1) + cc] = (cc <= r) ? a[cc] : 0.0;
            }
            __syncwarp();
            if (lane < CHOL_NB) {           // column `lane` of the inverse by forward substitution
                const int cc = lane;
                double x[CHOL_NB];
#pragma unroll
                for (int t = 0; t < CHOL_NB; t++) {
                    double sacc = 0.0;
#pragma unroll
                    for (int m = 0; m < t; m++) sacc = fma(L8[t * (CHOL_NB + 1) + m], x[m], sacc);
                    x[t] = (t < cc) ? 0.0 : ((t == cc) ? di[t] : -sacc * di[t]);
                    Li8[t * (CHOL_NB + 1) + cc] = x[t];
                }
            }
        }
        __syncthreads();
        // ---- (2) panel below the block (threads 0-63: one row each), the block itself (64-127), block row of L^{-1} (128-191)
        if (tid < 64) {
            const int t = tid;
            if (t >= j0 + CHOL_NB) {
                double a[CHOL_NB];
#pragma unroll
                for (int b = 0; b < CHOL_NB; b++) a[b] = A[t * CHOL_LD + j0 + b];
#pragma unroll
                for (int aa = 0; aa < CHOL_NB; aa++) {
                    double sacc = 0.0;
#pragma unroll
                    for (int b = 0; b <= aa; b++) sacc = fma(a[b], Li8[aa * (CHOL_NB + 1) + b], sacc);
                    A[t * CHOL_LD + j0 + aa] = ((*s_fmask >> aa) & 1) ? 0.0 : sacc;
                }
            }
        } else if (tid < 128) {
            const int r = (tid - 64) >> 3, cc = (tid - 64) & 7;
            A[(j0 + r) * CHOL_LD + j0 + cc] = L8[r * (CHOL_NB + 1) + cc];
        } else if (tid < 192) {
            const int cc = tid - 128;
            if (cc < j0 + CHOL_NB) {
                double bcol[CHOL_NB];
#pragma unroll
                for (int b = 0; b < CHOL_NB; b++) bcol[b] = B[(j0 + b) * CHOL_LD + cc];
#pragma unroll
                for (int aa = 0; aa < CHOL_NB; aa++) {
                    double sacc = 0.0;
#pragma unroll
                    for (int b = 0; b <= aa; b++) sacc = fma(Li8[aa * (CHOL_NB + 1) + b], bcol[b], sacc);
                    B[(j0 + aa) * CHOL_LD + cc] = sacc;
                }
            }
        }
        __syncthreads();
        // ---- (3) rank-8 update: trailing matrix (columns > block, lower triangle) and rows of L^{-1} below the block
        if (j0 + CHOL_NB < LP) {
            const bool trail = c >= j0 + CHOL_NB;
            double mine[CHOL_NB];              // l_{c, j0+a} (trailing) or (L^{-1})_{j0+a, c} (substitution)
#pragma unroll
            for (int aa = 0; aa < CHOL_NB; aa++)
                mine[aa] = trail ? A[c * CHOL_LD + j0 + aa] : B[(j0 + aa) * CHOL_LD + c];
            double* M = trail ? A : B;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int t = g + 8 * m;
                if (t >= j0 + CHOL_NB && (!trail || c <= t)) {
                    double acc = M[t * CHOL_LD + c];
#pragma unroll
                    for (int aa = 0; aa < CHOL_NB; aa++) acc = fma(-A[t * CHOL_LD + j0 + aa], mine[aa], acc);
                    M[t * CHOL_LD + c] = acc;
                }
            }
        }
        __syncthreads();
    }
    return bad;
}

template <typename T>
__global__ void __launch_bounds__(CHOL_THREADS)
chol_inv_kernel(const double* __restrict__ G, int k, double* __restrict__ R, double* __restrict__ Rinv,
                T* __restrict__ RinvT, int* __restrict__ flag, int drop) {
    extern __shared__ double chol_sm[];
    const double* A = chol_sm;
    const double* B = A + LP * CHOL_LD;
    const int tid = threadIdx.x, lane = tid & 31;
    const int c = tid & 63, g = tid >> 6;
    const int bad = chol_inv_block(G, LP, k, chol_sm, 0.0, drop);
    if ((bad & 1) && lane == 0) atomicOr(flag, 1);
    // outputs (row-major 64 x 64): R[r][i] = L[i][r], Rinv[r][i] = (L^{-1})[i][r]; thread (g, c) writes rows r = g + 8m,
    // column i = c (consecutive threads -> consecutive addresses)
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int r = g + 8 * m, i = c;
        const double x = (r <= i) ? B[i * CHOL_LD + r] : 0.0;
        if (Rinv) Rinv[r * LP + i] = x;
        if (RinvT) RinvT[r * LP + i] = (T)x;
        if (R) R[r * LP + i] = (r <= i) ? A[i * CHOL_LD + r] : 0.0;
    }
}

template <typename T>
void chol_inv(salg_ctx* ctx, const double* d_G, int k, double* d_R, double* d_Rinv, T* d_RinvT, int* d_flag, bool drop) {
    ProfScope ps(ctx, PROF_CHOL, 0.0);
    set_max_dyn_smem(chol_inv_kernel<T>, (int)((int)CHOL_SMEM));
    chol_inv_kernel<T><<<1, CHOL_THREADS, CHOL_SMEM, ctx->stream>>>(d_G, k, d_R, d_Rinv, d_RinvT, d_flag, drop ? 1 : 0);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void chol_inv<float>(salg_ctx*, const double*, int, double*, double*, float*, int*, bool);
template void chol_inv<double>(salg_ctx*, const double*, int, double*, double*, double*, int*, bool);

// ---- the replicated small side of one power-iteration half step in ONE launch ---------------------------------------------
// Given the all-reduced raw product Z0 = A_c^T Y (n x 64) the step needs Z2 = an orthonormal basis of span(Z0).  It used to be
//   Z1 = Z0 R1^{-1} (R1 = chol(Y^T Y): the tall-side normaliser applied on the small side),  Z2 = Z1 R2^{-1} (R2 = chol(Z1^T Z1)),
// i.e. chol_inv, panel_mul, panel_gram, chol_inv, panel_mul (+ column sums, |max|, pre-split) = 11 launches that every rank of a
// row-sharded fit repeats.  Here: CTAs 1..G take the Gram of the RAW panel in f64, GZ = Z0^T Z0 (per-CTA partials in fixed
// slots, summed in slot order: bit-identical on every rank and from run to run), CTA 0 factors and publishes M with Z2 = Z0 M
// together with the fp16 pre-split scale for Z2.  Two variants:
//   one step   M = chol(GZ)^{-1}: one 28 us Cholesky.  Z0 is already rounded to f32 when R1^{-1} would be applied, so the detour
//              through Z1 recovers nothing; it only keeps each Cholesky at cond kappa^2 instead of kappa^4 (kappa = sigma_1 /
//              sigma_l of the sketch).  Measured: config 2's raw-count sketch (kappa > 1000) floors a pivot this way.
//   two steps  CTA 0 factors Gy (while the others take the Gram), then G1 = R1^{-T} GZ R1^{-1} (= Z1^T Z1 without touching the
//              panel), M = R1^{-1} chol(G1)^{-1}: every Cholesky sees cond kappa^2, as in the explicit chain.
//   adaptive   one step; a pivot below 1e-9 of the largest diagonal entry switches this launch and the rest of the fit to two
//              steps (the decision is taken on bit-identical sums: all ranks agree).  Measured: configs 2 AND 3 both switch in
//              their first half step, so the default is two steps outright (SALG_ZSIDE_MODE = 1; 0 adaptive, 2 one step).
// Columns of Z2 have unit norm unless a pivot was floored; then the bound for the pre-split comes from diag(M^T GZ M).
// A second launch (tm.cu: tm_zside_apply_kernel) applies M, takes mu^T Z2 and writes the pre-split operand of the next A X.
constexpr int ZS_MAX_G = 64;
constexpr size_t ZS_SMEM = CHOL_SMEM + (size_t)4 * LP * CHOL_LD * sizeof(double);

// C (64 x 64, ld 65) = X Y (TX: X^T Y).  The first 128 threads hold 8 x 4 outputs each (rows ib + 8 u, columns jb + 16 y: every
// load instruction of a warp reads 16 consecutive doubles or two broadcast words): 12 shared-memory loads per 32 fused
// multiply-adds.  With all 512 threads at 8 outputs each, every warp re-read the same row of Y and the product was bound by
// the shared-memory pipe (~9 us per product, three per step).
template <bool TX>
__device__ __forceinline__ void zs_mat64(const double* X, const double* Y, double* Cm) {
    const int tid = threadIdx.x;
    if (tid >= 128) return;
    const int jb = tid & 15, ib = tid >> 4;
    double acc[8][4];
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
        for (int y = 0; y < 4; y++) acc[u][y] = 0.0;
#pragma unroll 2
    for (int m = 0; m < LP; m++) {
        double x[8], yv[4];
#pragma unroll
        for (int u = 0; u < 8; u++) x[u] = TX ? X[m * CHOL_LD + ib + 8 * u] : X[(ib + 8 * u) * CHOL_LD + m];
#pragma unroll
        for (int y = 0; y < 4; y++) yv[y] = Y[m * CHOL_LD + jb + 16 * y];
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int y = 0; y < 4; y++) acc[u][y] = fma(x[u], yv[y], acc[u][y]);
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
        for (int y = 0; y < 4; y++) Cm[(ib + 8 * u) * CHOL_LD + jb + 16 * y] = acc[u][y];
}

__global__ void __launch_bounds__(CHOL_THREADS)
zside_solve_kernel(const float* __restrict__ Z, int64_t n, const double* __restrict__ Gy, int k, float a_scale, int mode,
                   double* __restrict__ part /* [gridDim.x - 1][4096] */, unsigned* __restrict__ ticket,
                   float* __restrict__ M_out, float* __restrict__ scales, double* __restrict__ corr, int* __restrict__ flag) {
    extern __shared__ double chol_sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int G = (int)gridDim.x - 1;
    if (blockIdx.x > 0) {
        // ---- partial Gram of the raw panel in f64: two 32-row sub-tiles per trip (one per half of the CTA), 4 x 4 outputs per thread
        float (*tile)[LP + 4] = reinterpret_cast<float (*)[LP + 4]>(chol_sm);            // [64][68]
        double* red = chol_sm + (64 * (LP + 4) * 4) / 8;                                 // [256][16]
        const int half = tid >> 8, t8 = tid & 255, ti = t8 >> 4, tj = t8 & 15;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
        const int64_t n_tiles = (n + 63) / 64;
        for (int64_t t = blockIdx.x - 1; t < n_tiles; t += G) {
            const int64_t r0 = t * 64;
            __syncthreads();
            for (int i = tid; i < 64 * LP; i += CHOL_THREADS) {
                const int r = i >> 6, c = i & 63;
                tile[r][c] = (r0 + r < n) ? Z[(r0 + r) * LP + c] : 0.f;
            }
            __syncthreads();
#pragma unroll 4
            for (int r = 0; r < 32; r++) {
                const float4 av = *reinterpret_cast<const float4*>(&tile[32 * half + r][4 * ti]);
                const float4 bv = *reinterpret_cast<const float4*>(&tile[32 * half + r][4 * tj]);
                const double a[4] = {(double)av.x, (double)av.y, (double)av.z, (double)av.w};
                const double b[4] = {(double)bv.x, (double)bv.y, (double)bv.z, (double)bv.w};
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) acc[x][y] = fma(a[x], b[y], acc[x][y]);
            }
        }
        __syncthreads();
        if (half == 1) {
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) red[t8 * 16 + x * 4 + y] = acc[x][y];
        }
        __syncthreads();
        if (half == 0) {
            double* slot = part + (size_t)(blockIdx.x - 1) * (LP * LP);
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) slot[(4 * ti + x) * LP + 4 * tj + y] = acc[x][y] + red[t8 * 16 + x * 4 + y];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(ticket, 1u);
        return;
    }
    // ---- CTA 0
    double* B = chol_sm + LP * CHOL_LD;
    double* RiS = chol_sm + CHOL_SMEM / sizeof(double);
    double* S = RiS + LP * CHOL_LD;
    double* T = S + LP * CHOL_LD;
    double* U = T + LP * CHOL_LD;
    const int c = tid & 63, g = tid >> 6;
    // mode 0: adaptive — the first half step of a fit (random mixtures: cond kappa^4) takes two steps; later ones try the
    // one-step factorisation of GZ and fall back, in the same launch, when a column stands at less than 1e-3 rad to the span
    // of the columns before it (pivot below 1e-6 of its own squared norm).  ticket[1] = a half step was taken (host: 0 per fit).
    // 1: two steps always, 2: one step always.
    const bool sticky = mode == 1 || (mode == 0 && __ldcg(ticket + 1) == 0u);   // adaptive: the FIRST half step of a fit goes two steps outright
    int bad = 0;
    auto factor_gy = [&]() {
        bad |= chol_inv_block(Gy, LP, k, chol_sm) & 1;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int r = g + 8 * m;
            RiS[r * CHOL_LD + c] = (r <= c) ? B[c * CHOL_LD + r] : 0.0;        // R1^{-1}[r][c] = (L^{-1})[c][r]
        }
    };
    if (sticky) factor_gy();                                                   // (while the other CTAs take the Gram)
    if (tid < LP) corr[tid] = 0.0;
    if (tid == 0) {
        while (atomicAdd(ticket, 0u) < (unsigned)G) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
    {
        // slot sums in slot order; four slots' loads in flight per thread (the sum is latency-bound on one SM)
        double sacc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const double* p0 = part + (size_t)g * LP + c;                          // element (row g + 8 m, column c) = p0[m * 8 * LP]
        int q = 0;
        for (; q + 4 <= G; q += 4) {
            double v[4][8];
#pragma unroll
            for (int qq = 0; qq < 4; qq++)
#pragma unroll
                for (int m = 0; m < 8; m++) v[qq][m] = __ldcg(p0 + (size_t)(q + qq) * (LP * LP) + m * 8 * LP);
#pragma unroll
            for (int qq = 0; qq < 4; qq++)
#pragma unroll
                for (int m = 0; m < 8; m++) sacc[m] += v[qq][m];
        }
        for (; q < G; q++)
#pragma unroll
            for (int m = 0; m < 8; m++) sacc[m] += __ldcg(p0 + (size_t)q * (LP * LP) + m * 8 * LP);
#pragma unroll
        for (int m = 0; m < 8; m++) S[(g + 8 * m) * CHOL_LD + c] = sacc[m];
    }
    if (tid == 0) {
        *ticket = 0u;                                                         // ready for the next launch on this stream
        ticket[1] = 1u;                                                       // (a half step of this fit has been taken)
    }
    __syncthreads();
    bool two = sticky;
    if (!two) {
        const int f = chol_inv_block(S, CHOL_LD, k, chol_sm, mode == 2 ? 0.0 : 1e-6);
        const int weak = __syncthreads_or(f != 0 ? 1 : 0);
        if (weak && mode == 0) {
            __syncthreads();
            factor_gy();
            two = true;
        } else {
            bad |= f & 1;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int r = g + 8 * m;
                T[r * CHOL_LD + c] = (r <= c) ? B[c * CHOL_LD + r] : 0.0;      // M = R^{-1}
            }
        }
    }
    if (two) {
        __syncthreads();
        zs_mat64<false>(S, RiS, T);                                           // T = GZ R1^{-1}
        __syncthreads();
        zs_mat64<true>(RiS, T, U);                                            // U = R1^{-T} GZ R1^{-1} = Z1^T Z1
        __syncthreads();
        bad |= chol_inv_block(U, CHOL_LD, k, chol_sm) & 1;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int r = g + 8 * m;
            U[r * CHOL_LD + c] = (r <= c) ? B[c * CHOL_LD + r] : 0.0;          // R2^{-1}
        }
        __syncthreads();
        zs_mat64<false>(RiS, U, T);                                           // M = R1^{-1} R2^{-1}
    }
    const int any_bad = __syncthreads_or(bad != 0 ? 1 : 0);
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int r = g + 8 * m;
        M_out[r * LP + c] = (float)T[r * CHOL_LD + c];
    }
    if (any_bad) {
        // floored pivots: columns of Z2 are no longer unit vectors and M^T GZ M cancels too heavily to bound them, so the
        // largest |entry| of Z0 M is taken from the panel itself (one CTA, rare: rank-deficient sketches)
        if (tid == 0) atomicOr(flag, 1);
        __syncthreads();
        double mx = 0.0;
        for (int64_t r = g; r < n; r += 8) {
            double acc = 0.0;
            for (int kk = 0; kk < LP; kk++) acc = fma((double)Z[r * LP + kk], T[kk * CHOL_LD + c], acc);
            mx = fmax(mx, fabs(acc));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        if (lane == 0) RiS[tid >> 5] = mx;
        __syncthreads();
        if (tid == 0) {
            double m2 = 0.0;
            for (int w = 0; w < CHOL_THREADS / 32; w++) m2 = fmax(m2, RiS[w]);
            U[0] = m2;
        }
        __syncthreads();
    }
    if (tid == 0) {
        // |z| <= its column's norm = 1 up to rounding (no floored pivot), else the measured maximum; 1/16 of slack for the f32 apply
        const double d = any_bad ? U[0] : 1.0;
        const float bound = (isfinite(d) && d > 0.0) ? (float)(d * 1.0625) : 1.0625f;
        const float sc = tc_pow2_scale_dev(bound);
        scales[0] = sc;
        scales[1] = 1.f / (sc * a_scale);
    }
}

void zside_solve(salg_ctx* ctx, const float* Z, int64_t n, const double* d_Gy, int k, float a_scale, double* d_part,
                 unsigned* d_ticket, float* d_M, float* d_scales, double* d_corr, int* d_flag) {
    ProfScope ps(ctx, PROF_CHOL, 0.0);
    const int64_t tiles = ceil_div(n, 64);
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>(ZS_MAX_G, ceil_div(tiles, 2)));
    set_max_dyn_smem(zside_solve_kernel, (int)ZS_SMEM);
    zside_solve_kernel<<<1 + G, CHOL_THREADS, ZS_SMEM, ctx->stream>>>(Z, n, d_Gy, k, a_scale, zside_mode(), d_part, d_ticket, d_M,
                                                                      d_scales, d_corr, d_flag);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
// SALG_ZSIDE_MODE: 1 two steps (default), 0 adaptive, 2 one step always (experiments)
int zside_mode() {
    static const int v = getenv("SALG_ZSIDE_MODE") ? atoi(getenv("SALG_ZSIDE_MODE")) : 1;
    return v;
}
size_t zside_part_elems() { return (size_t)ZS_MAX_G * LP * LP; }

// ---- out = P * M (M 64 x 64 row-major, T); in place allowed ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
panel_mul_kernel(const T* P, int64_t m, const T* __restrict__ M, T* out, unsigned* __restrict__ amax_out) {
    constexpr int TR = 64;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    T (*Ms)[LP] = reinterpret_cast<T (*)[LP]>(dyn_smem);
    T (*Pt)[TR + 4] = reinterpret_cast<T (*)[TR + 4]>(dyn_smem + sizeof(T) * LP * LP);   // transposed tile: Pt[col][row]
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;   // rows 4ty..4ty+3, cols 4tx..4tx+3
    for (int i = tid; i < LP * LP; i += 256) Ms[i >> 6][i & 63] = M[i];
    const int64_t n_tiles = (m + TR - 1) / TR;
    float amax = 0.f;
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
                for (int y = 0; y < 4; y++) {
                    out[r * LP + 4 * tx + y] = acc[x][y];
                    amax = fmaxf(amax, fabsf((float)acc[x][y]));
                }
            }
        }
    }
    if (amax_out) {                      // bits of max |out| (zeroed by the caller)
#pragma unroll
        for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
        if ((tid & 31) == 0 && amax > 0.f) atomicMax(amax_out, __float_as_uint(amax));
    }
}

template <typename T>
void panel_mul(salg_ctx* ctx, const T* P, int64_t m, const T* d_M, T* out, unsigned* d_amax) {
    if (d_amax) SALG_CUDA(cudaMemsetAsync(d_amax, 0, 4, ctx->stream));
    if (m == 0) return;
    ProfScope ps(ctx, PROF_PANELMUL, 2.0 * (double)m * 60 * sizeof(T));
    int64_t want = ceil_div(m, 64);
    int64_t cap = (int64_t)ctx->sm_count * 4;
    constexpr int kSmem = (int)sizeof(T) * (LP * LP + LP * (64 + 4));
    set_max_dyn_smem(panel_mul_kernel<T>, (int)(kSmem));
    panel_mul_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, kSmem, ctx->stream>>>(P, m, d_M, out, d_amax);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}
template void panel_mul<float>(salg_ctx*, const float*, int64_t, const float*, float*, unsigned*);
template void panel_mul<double>(salg_ctx*, const double*, int64_t, const double*, double*, unsigned*);

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
// A = U diag(S) V^T, S descending.  Rows of W hold the columns of A V (so a column rotation touches two
// contiguous shared-memory rows); one warp per pair, round-robin tournament ordering.
// The ~59 rounds x ~8 sweeps are a serial chain of latencies (shuffle reductions, reciprocal / square root, barrier),
// and this kernel is replicated on every GPU of a row-sharded fit (0.75 ms in f64 throughout).  So the bulk of the sweeps
// runs in f32 (half the shuffles, hardware rsqrt / rcp), the accumulated V is then re-orthonormalised in f64 by two
// Newton-Schulz steps V <- V (3 I - V^T V) / 2 (f32 orthogonality error 1e-6 -> 1e-12 -> below f64 rounding),
// W = A V is recomputed in f64 and f64 sweeps finish the job (typically one rotating sweep and one confirming sweep).
template <typename F> struct JacMath;
template <> struct JacMath<double> {
    static __device__ __forceinline__ double rcp(double x) { return __drcp_rn(x); }
    static __device__ __forceinline__ double rsq(double x) { return rsqrt(x); }
};
template <> struct JacMath<float> {
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ float rsq(float x) { return rsqrtf(x); }
};

// returns the number of rotating sweeps; *s_rot is block-shared scratch.  EIGHT lanes per pair (lane s owns elements
// s, s + 8, ..., s + 56 of the two rows), four pairs per warp, 8 warps: the three dot products of a pair then cost three
// shuffle levels instead of five and a quarter of the shuffle instructions — with one warp per pair the 30 warps' 900
// shuffle instructions per round (f64: two each) ran into the SM's one-shuffle-per-clock limit (1.6 us per round).
// Only the first JAC_WARPS warps take part (named barrier); the caller __syncthreads() afterwards.
constexpr int JAC_WARPS = 8;
__device__ int g_jacobi_confirm = 0;          // SALG_JACOBI_CONFIRM=1: always finish with a sweep without rotations (round-2 behaviour)
template <typename F>
__device__ __forceinline__ int jacobi_sweeps(F (*W)[LP + 1], F (*Vt)[LP + 1], int ke, F tol, int max_sweeps, int* s_rot) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp >= JAC_WARPS) return 0;
    const int n_pairs = ke / 2;
    const int sub = lane & 7, pair = warp * 4 + (lane >> 3);
    const bool active = pair < n_pairs;
    auto bar = [] { asm volatile("bar.sync 1, %0;" ::"n"(JAC_WARPS * 32) : "memory"); };
    // Early exit on quadratic convergence: a sweep in which every pair already satisfied gamma^2 <= tol alpha beta (the
    // square root of the rotation threshold) leaves off-diagonal ratios of the order of their squares, i.e. below the
    // threshold — no confirming sweep without rotations is needed (one f64 and one or two f32 sweeps less per call).
    int sweep = 0;
    for (; sweep < max_sweeps; sweep++) {
        if (tid == 0) *s_rot = 0;
        bar();
        for (int round = 0; round < ke - 1; round++) {
            int p = 0, q = 1;
            if (active) {
                if (pair == 0) {
                    p = round % (ke - 1);
                    q = ke - 1;
                } else {
                    p = (round + pair) % (ke - 1);
                    q = (round - pair + (ke - 1)) % (ke - 1);
                }
                if (p > q) { int t = p; p = q; q = t; }
            }
            F wp[8], wq[8];
            F alpha = F(0), beta = F(0), gamma = F(0);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                wp[i] = W[p][8 * i + sub];
                wq[i] = W[q][8 * i + sub];
                alpha = fma(wp[i], wp[i], alpha);
                beta = fma(wq[i], wq[i], beta);
                gamma = fma(wp[i], wq[i], gamma);
            }
#pragma unroll
            for (int o = 4; o; o >>= 1) {
                alpha += __shfl_xor_sync(0xFFFFFFFFu, alpha, o);
                beta += __shfl_xor_sync(0xFFFFFFFFu, beta, o);
                gamma += __shfl_xor_sync(0xFFFFFFFFu, gamma, o);
            }
            // |gamma| > tol sqrt(alpha beta)  <=>  gamma^2 > tol^2 alpha beta  (no square root on the chain)
            // (columns whose squared norm is in or near the subnormal range are left to the f64 phase / are zero: the products
            // alpha * beta and the reciprocal of 2 gamma would under- / overflow: inf * 0 = NaN in the rotation parameters)
            const F tiny = sizeof(F) == 4 ? F(1e-30) : F(1e-280);
            if (active && alpha > tiny && beta > tiny && gamma * gamma > tol * tol * alpha * beta && gamma != F(0)) {
                F zeta = (beta - alpha) * JacMath<F>::rcp(F(2) * gamma);
                // sqrt(1 + zeta^2); beyond `big` it IS |zeta| to working precision, and zeta^2 would overflow (columns whose
                // norms differ by 1e17 and more, e.g. the 1e-14 rows a rank-deficient f64 sketch leaves in R beside sigma ~ 500:
                // inf * rsqrt(inf) = NaN used to poison the whole factor)
                const F az = fabs(zeta);
                const F big = sizeof(F) == 4 ? F(1e8) : F(1e150);
                F w = az;
                if (!(az > big)) {
                    const F h2 = fma(zeta, zeta, F(1));
                    w = h2 * JacMath<F>::rsq(h2);
                }
                F t = (zeta >= F(0) ? F(1) : F(-1)) * JacMath<F>::rcp(fabs(zeta) + w);
                F c = JacMath<F>::rsq(fma(t, t, F(1))), sn = c * t;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    W[p][8 * i + sub] = c * wp[i] - sn * wq[i];
                    W[q][8 * i + sub] = sn * wp[i] + c * wq[i];
                    const F vp = Vt[p][8 * i + sub], vq = Vt[q][8 * i + sub];
                    Vt[p][8 * i + sub] = c * vp - sn * vq;
                    Vt[q][8 * i + sub] = sn * vp + c * vq;
                }
                if (sub == 0) atomicOr(s_rot, (gamma * gamma > tol * alpha * beta) ? 3 : 1);   // bit 1: far from converged
            }
            bar();
        }
        int any = *s_rot;
        bar();
        if (!any) break;
        if (!(any & 2) && !g_jacobi_confirm) {
            sweep++;
            break;
        }
    }
    return sweep;
}

// C[i][j] (64 x 64, ld 65) = sum_m X[i][m] * Y[m][j]  (transX: X[m][i];  transY: Y[j][m]); 1024 threads, 4 outputs each
template <bool TX, bool TY>
__device__ __forceinline__ void mat64_smem(const double (*X)[LP + 1], const double (*Y)[LP + 1], double (*Cm)[LP + 1]) {
    const int tid = threadIdx.x;
    const int j = tid & 63, i0 = (tid >> 6) * 4;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int m = 0; m < LP; m++) {
        const double y = TY ? Y[j][m] : Y[m][j];
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = fma(TX ? X[m][i0 + u] : X[i0 + u][m], y, acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) Cm[i0 + u][j] = acc[u];
}

constexpr int JAC_SMEM = (4 * LP * (LP + 1)) * 8 + (2 * LP * (LP + 1)) * 4;

__global__ void __launch_bounds__(1024)
jacobi_svd64_kernel(const double* __restrict__ A, int k, double* __restrict__ U, double* __restrict__ S,
                    double* __restrict__ V, int* __restrict__ flag, int f32_sweeps, float tol32, int* __restrict__ dbg_sweeps) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double (*W)[LP + 1] = reinterpret_cast<double (*)[LP + 1]>(dyn_smem);
    double (*Vt)[LP + 1] = W + LP;
    double (*At)[LP + 1] = Vt + LP;        // At[c][r] = A[r][c]
    double (*T1)[LP + 1] = At + LP;        // scratch
    float (*W32)[LP + 1] = reinterpret_cast<float (*)[LP + 1]>(T1 + LP);
    float (*V32)[LP + 1] = W32 + LP;
    __shared__ double sig[LP];
    __shared__ int order[LP];
    __shared__ int s_rot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < LP * LP; i += 1024) {
        int r = i >> 6, c = i & 63;
        const double a = (r < k && c < k) ? A[r * LP + c] : 0.0;
        W[c][r] = a;
        At[c][r] = a;
        W32[c][r] = (float)a;
        Vt[r][c] = (r == c) ? 1.0 : 0.0;
        V32[r][c] = (r == c) ? 1.f : 0.f;
    }
    __syncthreads();
    const int ke = (k + 1) & ~1;        // even number of players (a zero column pads odd k)
    if (f32_sweeps > 0) {
        // scale-free in f32: the rotations only see ratios, but keep the entries in range
        const int s32 = jacobi_sweeps<float>(W32, V32, ke, tol32, f32_sweeps, &s_rot);
        if (dbg_sweeps && tid == 0) dbg_sweeps[0] = s32;
        __syncthreads();
        // Vt <- orth(V32): two Newton-Schulz steps in f64.  Vt rows are the columns of V: (V^T V)[i][j] = <Vt[i], Vt[j]>
        for (int i = tid; i < LP * LP; i += 1024) Vt[i >> 6][i & 63] = (double)V32[i >> 6][i & 63];
        __syncthreads();
        for (int it = 0; it < 2; it++) {
            mat64_smem<false, true>(Vt, Vt, T1);                 // T1 = Vt Vt^T = V^T V
            __syncthreads();
            for (int i = tid; i < LP * LP; i += 1024) {
                const int r = i >> 6, c = i & 63;
                T1[r][c] = (r == c ? 1.5 : 0.0) - 0.5 * T1[r][c];
            }
            __syncthreads();
            mat64_smem<false, false>(T1, Vt, W);                 // V <- V (1.5 I - 0.5 V^T V)  <=>  Vt <- (...)^T Vt, symmetric
            __syncthreads();
            for (int i = tid; i < LP * LP; i += 1024) Vt[i >> 6][i & 63] = W[i >> 6][i & 63];
            __syncthreads();
        }
        // W[p][r] = sum_j A[r][j] V[j][p] = sum_j At[j][r] Vt[p][j]
        mat64_smem<false, false>(Vt, At, W);
        __syncthreads();
    }
    const double tol = 1.5e-14;   // ~ k * eps: below the rounding noise of the k-term dot products
    const int sweep = jacobi_sweeps<double>(W, Vt, ke, tol, 30, &s_rot);
    __syncthreads();
    if (tid == 0 && sweep >= 30) atomicOr(flag, 2);
    if (dbg_sweeps && tid == 0) dbg_sweeps[1] = sweep;
    // singular values = row norms of W
    if (warp < 2) {
        int p = tid;   // 0..63
        double s = 0.0;
        for (int i = 0; i < LP; i++) s += W[p][i] * W[p][i];
        sig[p] = (p < k) ? sqrt(s) : -1.0;
    }
    __syncthreads();
    if (tid < LP) {
        // rank by descending sigma (stable on index); a non-finite value sorts last, so `order` is a permutation whatever
        // the input held (a NaN used to leave entries of `order` unset: out-of-range reads below)
        int rank = 0;
        const double raw = sig[tid];
        const double me = isfinite(raw) ? raw : -2.0;
        for (int j = 0; j < LP; j++) {
            const double oraw = sig[j];
            const double o = isfinite(oraw) ? oraw : -2.0;
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
    static const int f32_sweeps = getenv("SALG_JACOBI_F32") ? atoi(getenv("SALG_JACOBI_F32")) : 10;
    static const float tol32 = getenv("SALG_JACOBI_TOL32") ? (float)atof(getenv("SALG_JACOBI_TOL32")) : 4e-6f;
    static const bool dbg = getenv("SALG_JACOBI_DBG") != nullptr;
    static bool confirm_set = false;
    if (!confirm_set) {
        const int v = getenv("SALG_JACOBI_CONFIRM") ? 1 : 0;
        SALG_CUDA(cudaMemcpyToSymbolAsync(g_jacobi_confirm, &v, sizeof(int), 0, cudaMemcpyHostToDevice, ctx->stream));
        SALG_CUDA(cudaStreamSynchronize(ctx->stream));
        confirm_set = true;
    }
    DevBuf<int> d_dbg(dbg ? 2 : 0, ctx->stream);
    set_max_dyn_smem(jacobi_svd64_kernel, (int)(JAC_SMEM));
    jacobi_svd64_kernel<<<1, 1024, JAC_SMEM, ctx->stream>>>(d_A, k, d_U, d_S, d_V, d_flag, f32_sweeps, tol32,
                                                            dbg ? d_dbg.get() : nullptr);
    if (dbg) {
        int h[2] = {0, 0};
        cudaMemcpyAsync(h, d_dbg.get(), 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[jacobi] k=%d f32 sweeps %d (tol %.1e), f64 sweeps %d\n", k, h[0], (double)tol32, h[1]);
    }
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
        chol_inv<T>(ctx, G.get(), k, R, Ri, RiT.get(), d_flag, passes > 1 && pass == passes - 1);   // (last of two: drop what is still dependent)
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
