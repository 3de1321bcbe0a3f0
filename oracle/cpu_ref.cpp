// cpu_ref.cpp — multi-threaded CPU restatement of the reference's sparse-PCA path (C++17 + OpenMP).
//
// TEST INFRASTRUCTURE ONLY (like everything under oracle/): used by tests/ as a checker, by
// __graft_entry__.smoke() and by bench.py's `cpu_baseline` / `--impl reference` legs as the timed CPU
// baseline.  The product (single-algebra_b200/) never links, loads or calls it.
//
// What it restates (reference = /root/reference, crate single_algebra 0.9.2):
//   * SparsePCA::fit / MaskedSparsePCA::fit around the SVD: the three statistics passes sum_col, sum_col,
//     sum_col_squared over ALL columns (src/dimred/pca/sparse/mod.rs:106-131, sparse_masked/mod.rs:275-311),
//     mean_ = sum / n, total_var over the kept columns, explained_variance = s^2 / (n - 1) (:210-216 / :379-382),
//     svd_flip(u, vt, false) (:203 / :364), components_ = vt.
//   * single-svdlib 1.0.9 `randomized_svd` as called at pca/sparse/mod.rs:170-180 and
//     sparse_masked/mod.rs:341-351 (source un-vendored; published algorithm): column means of the
//     operator (one more pass), Y = A_c Om, q x { Y = qr(Y).q; Z = A_c^T Y; Z = qr(Z).q; Y = A_c Z }, Q = qr(Y).q,
//     B = Q^T A_c, SVD(B), U = Q U_B; A_c X = A X - 1 (mu^T X), A_c^T Y = A^T Y - mu (1^T Y) (never formed).
//     The products are row-parallel like the reference's Rayon loops (src/sparse/csr.rs:286-309,
//     sparse_masked/mod.rs:468-481); a MASKED operator streams the unmasked rows and tests the mask per stored
//     entry on every product, as `MaskedCSRMatrix` does.
//   * the normaliser is an unblocked Householder QR (nalgebra `qr()`), here with every reflector's dot products
//     and updates parallel over the rows — nalgebra's is serial, so this port is the GENEROUS side of a stand-in.
//   * transform: the intended projection (X - 1 mu^T) V^T on the kept columns.
// Same host-generated Omega as the CUDA path (explicit input).  Parity unpinned for the SVD engine (no golden
// vectors in the reference, no Rust toolchain): this file is cross-checked against oracle/oracle.py in
// tests/test_cpu_ref.py, which in turn is cross-checked against scikit-learn / LAPACK.
//
// Also here: the host-side synthetic-matrix generator (bit-identical to single-algebra_b200/synth.py and
// csrc/synth.cu) so that the reference arm of bench.py builds its input without touching the CUDA library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <omp.h>

namespace {

using clk = std::chrono::steady_clock;
static double now_s() { return std::chrono::duration<double>(clk::now().time_since_epoch()).count(); }

template <typename T>
struct Op {
    int64_t nrows, ncols;        // ncols = ALL columns of the stored matrix
    const int64_t* ptr;
    const uint32_t* idx;
    const T* val;
    const int32_t* remap;        // column -> compact id, -1 = masked out; nullptr = no mask
    int64_t n_eff;               // columns of the operator
};

// Y (nrows x L) = A X - 1 corr^T         (X: n_eff x L, row-major)
template <typename T>
void mul(const Op<T>& A, const T* __restrict__ X, int L, const double* corr, T* __restrict__ Y) {
#pragma omp parallel for schedule(dynamic, 512)
    for (int64_t r = 0; r < A.nrows; r++) {
        T acc[64];
        for (int j = 0; j < L; j++) acc[j] = T(0);
        for (int64_t p = A.ptr[r]; p < A.ptr[r + 1]; p++) {
            int64_t c = A.idx[p];
            if (A.remap) {
                int32_t m = A.remap[c];
                if (m < 0) continue;
                c = m;
            }
            const T v = A.val[p];
            const T* __restrict__ x = X + c * L;
            for (int j = 0; j < L; j++) acc[j] += v * x[j];
        }
        T* y = Y + r * L;
        if (corr)
            for (int j = 0; j < L; j++) y[j] = (T)((double)acc[j] - corr[j]);
        else
            for (int j = 0; j < L; j++) y[j] = acc[j];
    }
}

// Z (n_eff x L) = A^T Y - mu cs^T, cs = 1^T Y
template <typename T>
void mul_t(const Op<T>& A, const T* __restrict__ Y, int L, const T* mu, T* __restrict__ Z) {
    const int64_t n = A.n_eff;
    const int nt = omp_get_max_threads();
    std::vector<std::vector<T>> loc(nt);
    std::vector<std::vector<double>> lcs(nt);
#pragma omp parallel
    {
        const int t = omp_get_thread_num();
        loc[t].assign((size_t)n * L, T(0));
        lcs[t].assign(L, 0.0);
        T* z = loc[t].data();
        double* cs = lcs[t].data();
#pragma omp for schedule(dynamic, 512)
        for (int64_t r = 0; r < A.nrows; r++) {
            const T* __restrict__ y = Y + r * L;
            for (int j = 0; j < L; j++) cs[j] += (double)y[j];
            for (int64_t p = A.ptr[r]; p < A.ptr[r + 1]; p++) {
                int64_t c = A.idx[p];
                if (A.remap) {
                    int32_t m = A.remap[c];
                    if (m < 0) continue;
                    c = m;
                }
                const T v = A.val[p];
                T* __restrict__ zz = z + c * L;
                for (int j = 0; j < L; j++) zz[j] += v * y[j];
            }
        }
    }
    std::vector<double> cs(L, 0.0);
    for (int t = 0; t < nt; t++)
        for (int j = 0; j < L; j++) cs[j] += lcs[t][j];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        for (int j = 0; j < L; j++) {
            double s = 0.0;
            for (int t = 0; t < nt; t++) s += (double)loc[t][(size_t)i * L + j];
            if (mu) s -= (double)mu[i] * cs[j];
            Z[(size_t)i * L + j] = (T)s;
        }
    }
}

// column sums (or sums of squares) over ALL stored columns, f64 accumulation
template <typename T>
void col_sums(const Op<T>& A, bool squared, double* out) {
    const int nt = omp_get_max_threads();
    std::vector<std::vector<double>> loc(nt);
#pragma omp parallel
    {
        const int t = omp_get_thread_num();
        loc[t].assign((size_t)A.ncols, 0.0);
        double* s = loc[t].data();
#pragma omp for schedule(static)
        for (int64_t r = 0; r < A.nrows; r++)
            for (int64_t p = A.ptr[r]; p < A.ptr[r + 1]; p++) {
                const double v = (double)A.val[p];
                s[A.idx[p]] += squared ? v * v : v;
            }
    }
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < A.ncols; c++) {
        double s = 0.0;
        for (int t = 0; t < nt; t++) s += loc[t][c];
        out[c] = s;
    }
}

// Unblocked Householder QR of the m x L panel P (row-major, overwritten by Q, thin); R (L x L, row-major, f64)
// optional.  Reflector j: v = P[j:, j] with v_j = 1; dot products and updates parallel over rows.
template <typename T>
void householder_qr(T* __restrict__ P, int64_t m, int L, double* R) {
    const int K = (int)std::min<int64_t>(m, L);
    std::vector<double> tau(K, 0.0), Rm((size_t)L * L, 0.0);
    std::vector<double> w(L);
    for (int j = 0; j < K; j++) {
        // w_k = sum_{i >= j} P[i][j] * P[i][k], k = j .. L-1   (w_j = squared norm of the column tail)
        std::fill(w.begin(), w.end(), 0.0);
        const int nk = L - j;
#pragma omp parallel
        {
            double lw[64] = {0};
#pragma omp for schedule(static) nowait
            for (int64_t i = j; i < m; i++) {
                const T* row = P + i * L + j;
                const double a = (double)row[0];
                for (int k = 0; k < nk; k++) lw[k] += a * (double)row[k];
            }
#pragma omp critical
            for (int k = 0; k < nk; k++) w[j + k] += lw[k];
        }
        const double alpha = (double)P[(int64_t)j * L + j];
        const double norm = std::sqrt(w[j]);
        if (norm == 0.0) { tau[j] = 0.0; continue; }
        const double beta = alpha >= 0 ? -norm : norm;
        const double v0 = alpha - beta;               // unnormalised v_j
        tau[j] = (beta - alpha) / beta;
        // v = x - beta e_j, normalised by v0 so that v_j = 1;  v^T P[:,k] = (w_k - beta * P[j][k]) / v0
        Rm[(size_t)j * L + j] = beta;
        std::vector<double> coef(L, 0.0);
        for (int k = j + 1; k < L; k++) {
            const double vtp = (w[k] - beta * (double)P[(int64_t)j * L + k]) / v0;
            coef[k] = tau[j] * vtp;
        }
        // update trailing columns: P[i][k] -= v_i coef_k, v_i = P[i][j] / v0 (i > j), v_j = 1; store v in column j
        for (int k = j + 1; k < L; k++) {
            double pjk = (double)P[(int64_t)j * L + k] - coef[k];
            Rm[(size_t)j * L + k] = pjk;
            P[(int64_t)j * L + k] = (T)pjk;
        }
        P[(int64_t)j * L + j] = T(1);
        const double inv = 1.0 / v0;
#pragma omp parallel for schedule(static)
        for (int64_t i = j + 1; i < m; i++) {
            T* row = P + i * L;
            const double vi = (double)row[j] * inv;
            row[j] = (T)vi;
            for (int k = j + 1; k < L; k++) row[k] = (T)((double)row[k] - vi * coef[k]);
        }
    }
    if (R) std::memcpy(R, Rm.data(), sizeof(double) * (size_t)L * L);
    // form thin Q = H_0 ... H_{K-1} [I; 0] in a second buffer, then copy back
    std::vector<T> Q((size_t)m * L, T(0));
    for (int k = 0; k < K; k++) Q[(size_t)k * L + k] = T(1);
    for (int j = K - 1; j >= 0; j--) {
        if (tau[j] == 0.0) continue;
        // columns k >= j of Q can be non-zero below row j
        const int nk = L - j;
        std::fill(w.begin(), w.end(), 0.0);
#pragma omp parallel
        {
            double lw[64] = {0};
#pragma omp for schedule(static) nowait
            for (int64_t i = j; i < m; i++) {
                const double vi = (i == j) ? 1.0 : (double)P[i * L + j];
                const T* q = Q.data() + i * L + j;
                for (int k = 0; k < nk; k++) lw[k] += vi * (double)q[k];
            }
#pragma omp critical
            for (int k = 0; k < nk; k++) w[j + k] += lw[k];
        }
        const double tj = tau[j];
#pragma omp parallel for schedule(static)
        for (int64_t i = j; i < m; i++) {
            const double vi = (i == j) ? 1.0 : (double)P[i * L + j];
            T* q = Q.data() + i * L + j;
            for (int k = 0; k < nk; k++) q[k] = (T)((double)q[k] - tj * vi * w[j + k]);
        }
    }
    std::memcpy(P, Q.data(), sizeof(T) * (size_t)m * L);
}

// one-sided Jacobi SVD of the n x n matrix M (row-major f64): M = U diag(s) V^T, s descending
void jacobi_svd(const double* M, int n, double* U, double* s, double* V) {
    std::vector<double> A(M, M + (size_t)n * n), W((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) W[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double a = 0, b = 0, c = 0;
                for (int i = 0; i < n; i++) {
                    const double x = A[(size_t)i * n + p], y = A[(size_t)i * n + q];
                    a += x * x; b += y * y; c += x * y;
                }
                if (std::fabs(c) <= 1e-300 || std::fabs(c) <= 1e-16 * std::sqrt(a * b)) continue;
                off = std::max(off, std::fabs(c) / std::sqrt(a * b));
                const double zeta = (b - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < n; i++) {
                    double x = A[(size_t)i * n + p], y = A[(size_t)i * n + q];
                    A[(size_t)i * n + p] = cs * x - sn * y;
                    A[(size_t)i * n + q] = sn * x + cs * y;
                    x = W[(size_t)i * n + p]; y = W[(size_t)i * n + q];
                    W[(size_t)i * n + p] = cs * x - sn * y;
                    W[(size_t)i * n + q] = sn * x + cs * y;
                }
            }
        if (off < 1e-15) break;
    }
    std::vector<double> nr(n);
    std::vector<int> ord(n);
    for (int j = 0; j < n; j++) {
        double a = 0;
        for (int i = 0; i < n; i++) a += A[(size_t)i * n + j] * A[(size_t)i * n + j];
        nr[j] = std::sqrt(a);
        ord[j] = j;
    }
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return nr[a] > nr[b]; });
    for (int jj = 0; jj < n; jj++) {
        const int j = ord[jj];
        s[jj] = nr[j];
        for (int i = 0; i < n; i++) {
            U[(size_t)i * n + jj] = nr[j] > 0 ? A[(size_t)i * n + j] / nr[j] : 0.0;
            V[(size_t)i * n + jj] = W[(size_t)i * n + j];
        }
    }
}

template <typename T>
int pca_fit(int64_t nrows, int64_t ncols, const int64_t* ptr, const uint32_t* idx, const T* val, const uint8_t* mask,
            int k, int p, int q, int center, const T* omega, T* components, double* sv, double* ev, double* mean,
            double* total_var, T* scores, double* timings) {
    if (nrows < 2 || ncols < 1) return 1;
    double t0 = now_s();
    std::vector<int32_t> remap;
    int64_t n_eff = ncols;
    if (mask) {
        remap.resize((size_t)ncols);
        n_eff = 0;
        for (int64_t c = 0; c < ncols; c++) remap[c] = mask[c] ? (int32_t)n_eff++ : -1;
    }
    if (n_eff < 1) return 2;
    Op<T> A{nrows, ncols, ptr, idx, val, mask ? remap.data() : nullptr, n_eff};
    const int rank = (int)std::min<int64_t>(k, std::min<int64_t>(nrows, n_eff));
    const int L = rank + p;
    if (L > 64) return 3;
    // ---- the reference's three statistics passes (pca/sparse/mod.rs:107,121,122; sparse_masked/mod.rs:279,299,300)
    std::vector<double> s1((size_t)ncols), s2((size_t)ncols), sq((size_t)ncols);
    col_sums(A, false, s1.data());
    col_sums(A, false, s2.data());
    col_sums(A, true, sq.data());
    const double n_d = (double)nrows;
    double tv = 0.0;
    for (int64_t c = 0; c < ncols; c++) {
        mean[c] = center ? s1[c] / n_d : 0.0;
        if (!mask || mask[c]) {
            const double m = s2[c] / n_d;
            tv += (sq[c] - m * s2[c]) / (n_d - 1.0);
        }
    }
    double t1 = now_s();
    // ---- svdlib: column means of the operator (its own pass), then the power iteration
    std::vector<T> mu((size_t)n_eff, T(0));
    if (center) {
        std::vector<double> s3((size_t)ncols);
        col_sums(A, false, s3.data());
        for (int64_t c = 0; c < ncols; c++) {
            const int64_t m = mask ? remap[c] : c;
            if (m >= 0) mu[m] = (T)(s3[c] / n_d);
        }
    }
    double t_stats = now_s();
    std::vector<T> Y((size_t)nrows * L), Z((size_t)n_eff * L);
    std::vector<double> corr(L);
    auto set_corr = [&](const T* X) {
        for (int j = 0; j < L; j++) corr[j] = 0.0;
        if (!center) return;
        for (int64_t i = 0; i < n_eff; i++)
            for (int j = 0; j < L; j++) corr[j] += (double)mu[i] * (double)X[(size_t)i * L + j];
    };
    double t_mul = 0, t_qr = 0, tt;
    tt = now_s(); set_corr(omega); mul(A, omega, L, center ? corr.data() : nullptr, Y.data()); t_mul += now_s() - tt;
    for (int it = 0; it < q; it++) {
        tt = now_s(); householder_qr(Y.data(), nrows, L, nullptr); t_qr += now_s() - tt;
        tt = now_s(); mul_t(A, Y.data(), L, center ? mu.data() : nullptr, Z.data()); t_mul += now_s() - tt;
        tt = now_s(); householder_qr(Z.data(), n_eff, L, nullptr); t_qr += now_s() - tt;
        tt = now_s(); set_corr(Z.data()); mul(A, Z.data(), L, center ? corr.data() : nullptr, Y.data()); t_mul += now_s() - tt;
    }
    tt = now_s(); householder_qr(Y.data(), nrows, L, nullptr); t_qr += now_s() - tt;
    tt = now_s(); mul_t(A, Y.data(), L, center ? mu.data() : nullptr, Z.data()); t_mul += now_s() - tt;   // B^T (n_eff x L)
    // ---- SVD of B: B^T = Q_B R_B, R_B = U_R S V_R^T  =>  B = V_R S (Q_B U_R)^T
    double t_svd0 = now_s();
    std::vector<double> Rb((size_t)L * L), Ur((size_t)L * L), Vr((size_t)L * L), S(L);
    householder_qr(Z.data(), n_eff, L, Rb.data());
    jacobi_svd(Rb.data(), L, Ur.data(), S.data(), Vr.data());
    // V (n_eff x rank) = Q_B U_R[:, :rank]
    std::vector<T> V((size_t)n_eff * rank);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_eff; i++)
        for (int c = 0; c < rank; c++) {
            double s = 0.0;
            for (int j = 0; j < L; j++) s += (double)Z[(size_t)i * L + j] * Ur[(size_t)j * L + c];
            V[(size_t)i * rank + c] = (T)s;
        }
    // svd_flip(u, vt, false): the largest-magnitude entry of every component becomes positive
    for (int c = 0; c < rank; c++) {
        int64_t best = 0;
        double bv = -1.0;
        for (int64_t i = 0; i < n_eff; i++) {
            const double a = std::fabs((double)V[(size_t)i * rank + c]);
            if (a > bv) { bv = a; best = i; }
        }
        if (V[(size_t)best * rank + c] < 0)
            for (int64_t i = 0; i < n_eff; i++) V[(size_t)i * rank + c] = -V[(size_t)i * rank + c];
    }
    for (int c = 0; c < rank; c++) {
        sv[c] = S[c];
        ev[c] = S[c] * S[c] / (n_d - 1.0);
        for (int64_t i = 0; i < n_eff; i++) components[(size_t)c * n_eff + i] = V[(size_t)i * rank + c];
    }
    if (!center) {
        tv = 0.0;
        for (int c = 0; c < rank; c++) tv += ev[c];
    }
    *total_var = tv;
    double t_svd = now_s() - t_svd0;
    // ---- transform of the fitted rows: (X - 1 mu^T) V
    double t_tr0 = now_s();
    if (scores) {
        std::vector<double> cr(rank, 0.0);
        if (center)
            for (int64_t i = 0; i < n_eff; i++)
                for (int c = 0; c < rank; c++) cr[c] += (double)mu[i] * (double)V[(size_t)i * rank + c];
        mul(A, V.data(), rank, center ? cr.data() : nullptr, scores);
    }
    double t_end = now_s();
    if (timings) {
        timings[0] = t1 - t0;             // three statistics passes
        timings[1] = t_stats - t1;        // svdlib's column means
        timings[2] = t_mul;               // 2q + 2 products
        timings[3] = t_qr;                // 2q + 1 Householder QRs
        timings[4] = t_svd;               // QR + Jacobi of B, V, flip
        timings[5] = t_end - t_tr0;       // transform
        timings[6] = t_end - t0;
    }
    return 0;
}

// ---- synthetic generator (bit-identical to synth.py::generate_rows / synth.cu) ----------------------------------------
inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
struct Synth {
    uint64_t seed; int64_t row0, nrows, ncols; int n_clusters;
    const uint8_t* base_level; const int32_t* sf_offset; const uint32_t* cdf;
};
inline void row_meta(const Synth& a, int64_t row, int& cluster, int& sfo, uint64_t& hrow) {
    const uint64_t r = (uint64_t)row;
    const uint64_t hr = mix64((a.seed ^ 0xA5A5A5A5DEADBEEFULL) + r * 0x9E3779B97F4A7C15ULL);
    cluster = (int)((hr & 0xFFFFULL) % (uint64_t)a.n_clusters);
    sfo = a.sf_offset[(hr >> 16) & 15ULL];
    hrow = mix64(a.seed + r * 0x9E3779B97F4A7C15ULL);
}
inline int cell_value(const Synth& a, const uint8_t* lvl_row, int sfo, uint64_t hrow, int64_t col) {
    const uint64_t h = mix64(hrow ^ ((uint64_t)col * 0xD1B54A32D192ED03ULL + 0x8CB92BA72F3D8DD7ULL));
    const uint32_t u = (uint32_t)(h >> 32);
    int lvl = (int)lvl_row[col] + sfo;
    lvl = lvl < 0 ? 0 : (lvl > 255 ? 255 : lvl);
    const uint32_t* c = a.cdf + lvl * 40;
    if (u < c[0]) return 0;
    int x = 1;
    while (x < 40 && u >= c[x]) x++;
    return x;
}

}  // namespace

extern "C" {

int cpuref_max_threads() { return omp_get_max_threads(); }
void cpuref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

int cpuref_pca_fit_f32(int64_t nrows, int64_t ncols, const int64_t* ptr, const uint32_t* idx, const float* val,
                       const uint8_t* mask, int k, int p, int q, int center, const float* omega, float* components,
                       double* sv, double* ev, double* mean, double* total_var, float* scores, double* timings) {
    return pca_fit<float>(nrows, ncols, ptr, idx, val, mask, k, p, q, center, omega, components, sv, ev, mean, total_var,
                          scores, timings);
}
int cpuref_pca_fit_f64(int64_t nrows, int64_t ncols, const int64_t* ptr, const uint32_t* idx, const double* val,
                       const uint8_t* mask, int k, int p, int q, int center, const double* omega, double* components,
                       double* sv, double* ev, double* mean, double* total_var, double* scores, double* timings) {
    return pca_fit<double>(nrows, ncols, ptr, idx, val, mask, k, p, q, center, omega, components, sv, ev, mean, total_var,
                           scores, timings);
}

// sum_col / sum_col_squared (src/sparse/csr.rs:259-312, 558-608), f64 accumulation
void cpuref_col_sums_f32(int64_t nrows, int64_t ncols, const int64_t* ptr, const uint32_t* idx, const float* val,
                         int squared, double* out) {
    Op<float> A{nrows, ncols, ptr, idx, val, nullptr, ncols};
    col_sums(A, squared != 0, out);
}

// The preprocessing chain single-rust runs before PCA, as the reference's FIVE separate passes:
// sum_row (src/sparse/csr.rs:314-392, row-parallel), normalize ROW (:1013-1068: scale = target / sum where sum > 0, rows
// with scale <= 0 untouched; serial in the reference, row-parallel here), log1p as ln(fl(1 + v)) (:1070-1079),
// sum_col, sum_col_squared (:259-312, 558-608).  Values are updated in place.
void cpuref_preprocess_f32(int64_t nrows, int64_t ncols, const int64_t* ptr, const uint32_t* idx, float* val,
                           float target, double* sum, double* sumsq) {
    std::vector<float> rs((size_t)nrows);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrows; r++) {
        float a = 0.f;
        for (int64_t p = ptr[r]; p < ptr[r + 1]; p++) a += val[p];
        rs[r] = a;
    }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrows; r++) {
        const float scale = rs[r] > 0.f ? target / rs[r] : 0.f;
        if (scale > 0.f)
            for (int64_t p = ptr[r]; p < ptr[r + 1]; p++) val[p] = val[p] * scale;
    }
    const int64_t nnz = ptr[nrows];
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < nnz; p++) {
        const float t = 1.0f + val[p];
        val[p] = std::log(t);
    }
    Op<float> A{nrows, ncols, ptr, idx, val, nullptr, ncols};
    col_sums(A, false, sum);
    col_sums(A, true, sumsq);
}

// rows [row0, row0 + nrows) of the synthetic matrix: pass 1 (idx == NULL) fills ptr[0..nrows] and returns the entry
// count; pass 2 writes idx / val (f32) at the offsets of ptr
int64_t cpuref_synth(uint64_t seed, int64_t row0, int64_t nrows, int64_t ncols, int n_clusters, const uint8_t* base_level,
                     const int32_t* sf_offset, const uint32_t* cdf, int64_t* ptr, uint32_t* idx, float* val) {
    Synth a{seed, row0, nrows, ncols, n_clusters, base_level, sf_offset, cdf};
    if (!idx) {
        ptr[0] = 0;
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t r = 0; r < nrows; r++) {
            int cluster, sfo; uint64_t hrow;
            row_meta(a, row0 + r, cluster, sfo, hrow);
            const uint8_t* lvl_row = base_level + (size_t)cluster * ncols;
            // first threshold of every level for this row's size factor (256 entries: stays in L1)
            uint32_t c0[256];
            for (int l = 0; l < 256; l++) {
                int lv = l + sfo;
                lv = lv < 0 ? 0 : (lv > 255 ? 255 : lv);
                c0[l] = cdf[lv * 40];
            }
            int64_t n = 0;
            for (int64_t c = 0; c < ncols; c++) {
                const uint64_t h = mix64(hrow ^ ((uint64_t)c * 0xD1B54A32D192ED03ULL + 0x8CB92BA72F3D8DD7ULL));
                n += (uint32_t)(h >> 32) >= c0[lvl_row[c]];
            }
            ptr[r + 1] = n;
        }
        for (int64_t r = 0; r < nrows; r++) ptr[r + 1] += ptr[r];
        return ptr[nrows];
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < nrows; r++) {
        int cluster, sfo; uint64_t hrow;
        row_meta(a, row0 + r, cluster, sfo, hrow);
        const uint8_t* lvl_row = base_level + (size_t)cluster * ncols;
        int64_t o = ptr[r];
        for (int64_t c = 0; c < ncols; c++) {
            const int x = cell_value(a, lvl_row, sfo, hrow, c);
            if (x) { idx[o] = (uint32_t)c; val[o] = (float)x; o++; }
        }
    }
    return ptr[nrows];
}

}  // extern "C"
