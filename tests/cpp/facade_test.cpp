// facade_test.cpp — the reference's own tests for the hot path, written against the C++ facade
// (include/single_algebra.hpp) exactly as the Rust tests are written against the crate:
//   test_csr_normalize                      src/sparse/csr.rs:1514-1550   (KAT-N1)
//   test_matrix_sum / test_csc_normalization / test_zero_elements   src/sparse/csc.rs:1123-1152, 1256-1301, 1303-1314 (KAT-S1, N2, L1)
//   test_random_matrix_sparse_svd_comp_random  src/dimred/pca/sparse/mod.rs:540-562 (asserts is_ok(); here at a reduced
//                                           size by default, `--full` runs the reference's 10M x 2500 at 1 %)
// plus the error behaviour of the masked type (pca/sparse_masked/mod.rs:258-262, 440-444) and plain-loop checks of
// MatrixSum / Log1P on a small ragged matrix (empty rows, an empty column).
// Needs a CUDA device (the library has no CPU fallback); prints one line per check and exits non-zero on failure.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <set>
#include <string>

#include "single_algebra.hpp"

using namespace single_algebra;

static int g_fail = 0;
#define CHECK(cond, ...)                                   \
    do {                                                   \
        if (!(cond)) {                                     \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);                      \
            std::printf("\n");                             \
            g_fail++;                                      \
        }                                                  \
    } while (0)

// CooMatrix::try_from_triplets + CsrMatrix::from(&coo) for triplets already sorted by (row, col)
template <typename T>
static CsrMatrix<T> from_triplets(std::size_t nrows, std::size_t ncols, const std::vector<std::size_t>& ri,
                                  const std::vector<std::size_t>& ci, const std::vector<T>& v) {
    std::vector<std::uint64_t> off(nrows + 1, 0), idx(ci.begin(), ci.end());
    for (auto r : ri) off[r + 1]++;
    for (std::size_t r = 0; r < nrows; r++) off[r + 1] += off[r];
    return CsrMatrix<T>(nrows, ncols, off, idx, v);
}

// src/sparse/csr.rs:1514-1550
static void test_csr_normalize() {
    const std::vector<std::size_t> ri{0, 0, 1, 1, 2}, ci{0, 1, 1, 2, 2};
    const std::vector<double> v{2.0, 3.0, 4.0, 1.0, 2.0};
    auto csr = from_triplets<double>(3, 3, ri, ci, v);
    csr.normalize(std::vector<double>{2.0, 7.0, 3.0}, 1.0, Direction::COLUMN);
    const double ec[5] = {1.0, 3.0 / 7.0, 4.0 / 7.0, 1.0 / 3.0, 2.0 / 3.0};
    for (int i = 0; i < 5; i++) CHECK(std::fabs(csr.values()[i] - ec[i]) < 1e-10, "COLUMN value %d = %.17g", i, csr.values()[i]);
    auto csr2 = from_triplets<double>(3, 3, ri, ci, v);
    csr2.normalize(std::vector<double>{5.0, 5.0, 2.0}, 1.0, Direction::ROW);
    const double er[5] = {0.4, 0.6, 0.8, 0.2, 1.0};
    for (int i = 0; i < 5; i++) CHECK(std::fabs(csr2.values()[i] - er[i]) < 1e-10, "ROW value %d = %.17g", i, csr2.values()[i]);
    // a `sums` shorter than the normalised dimension: the reference panics, the facade returns an error
    bool threw = false;
    try {
        csr2.normalize(std::vector<double>{1.0}, 1.0, Direction::ROW);
    } catch (const Error& e) {
        threw = e.code == SALG_ERR_BAD_ARG;
    }
    CHECK(threw, "short sums must be an Error");
    std::printf("ok test_csr_normalize\n");
}

// MatrixSum + Log1P against plain loops on a ragged matrix (row 1 and row 4 empty, column 3 empty)
template <typename T>
static void test_sums_and_log1p(const char* name) {
    const std::size_t nr = 6, nc = 5;
    const std::vector<std::size_t> ri{0, 0, 0, 2, 2, 3, 5, 5, 5, 5}, ci{0, 2, 4, 1, 2, 0, 0, 1, 2, 4};
    const std::vector<T> v{1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
    auto a = from_triplets<T>(nr, nc, ri, ci, v);
    std::vector<double> sc(nc, 0), sq(nc, 0), sr(nr, 0);
    for (std::size_t i = 0; i < v.size(); i++) {
        sc[ci[i]] += v[i];
        sq[ci[i]] += (double)v[i] * v[i];
        sr[ri[i]] += v[i];
    }
    auto g_sc = a.sum_col();
    auto g_sq = a.sum_col_squared();
    auto g_sr = a.sum_row();
    for (std::size_t c = 0; c < nc; c++) {
        CHECK(g_sc[c] == (T)sc[c], "%s sum_col[%zu] = %g, want %g", name, c, (double)g_sc[c], sc[c]);
        CHECK(g_sq[c] == (T)sq[c], "%s sum_col_squared[%zu] = %g, want %g", name, c, (double)g_sq[c], sq[c]);
    }
    for (std::size_t r = 0; r < nr; r++) CHECK(g_sr[r] == (T)sr[r], "%s sum_row[%zu] = %g, want %g", name, r, (double)g_sr[r], sr[r]);
    a.log1p_normalize();      // v <- ln(fl(1 + v)) (src/sparse/csr.rs:1074-1075)
    const double tol = sizeof(T) == 4 ? 2e-7 : 1e-15;
    for (std::size_t i = 0; i < v.size(); i++) {
        const double want = std::log((double)(T)(T(1) + v[i]));
        CHECK(std::fabs(a.values()[i] - want) <= tol * std::fabs(want), "%s log1p value %zu = %.17g, want %.17g", name, i,
              (double)a.values()[i], want);
    }
    // values_mut() hands the host values out: the next operation must see the edit
    a.values_mut()[0] = T(100);
    auto again = a.sum_row();
    double want0 = 100.0 + a.values()[1] + a.values()[2];
    CHECK(std::fabs(again[0] - want0) <= 1e-5 * want0, "%s sum_row after values_mut = %g, want %g", name, (double)again[0], want0);
    std::printf("ok test_sums_and_log1p<%s>\n", name);
}

// CooMatrix -> CscMatrix for small triplet lists (any order)
static CscMatrix<double> csc_from_triplets(std::size_t nrows, std::size_t ncols, const std::vector<std::size_t>& ri,
                                           const std::vector<std::size_t>& ci, const std::vector<double>& v) {
    std::vector<std::uint64_t> off(ncols + 1, 0), idx;
    std::vector<double> val;
    for (std::size_t c = 0; c < ncols; c++) {
        for (std::size_t r = 0; r < nrows; r++)
            for (std::size_t i = 0; i < v.size(); i++)
                if (ci[i] == c && ri[i] == r) {
                    idx.push_back(r);
                    val.push_back(v[i]);
                }
        off[c + 1] = idx.size();
    }
    return CscMatrix<double>(nrows, ncols, off, idx, val);
}

// src/sparse/csc.rs:1123-1152 (matrix of :1071-1093), :1256-1301, :1303-1314
static void test_csc_twins() {
    auto m = csc_from_triplets(3, 3, {0, 2, 1, 0, 2}, {0, 0, 1, 2, 2}, {1.0, 4.0, 3.0, 2.0, 5.0});
    CHECK((m.sum_col() == std::vector<double>{5.0, 3.0, 7.0}), "Column sums should match");
    CHECK((m.sum_row() == std::vector<double>{3.0, 3.0, 9.0}), "Row sums should match");
    CHECK((m.sum_col_squared() == std::vector<double>{17.0, 9.0, 29.0}), "Column sums of squares");
    const std::vector<std::size_t> ri{0, 1, 2, 0, 2}, ci{0, 1, 0, 2, 2};
    const std::vector<double> v{2.0, 4.0, 1.0, 3.0, 5.0};
    auto csc = csc_from_triplets(3, 3, ri, ci, v);
    csc.normalize(std::vector<double>{3.0, 4.0, 8.0}, 1.0, Direction::COLUMN);
    for (std::size_t c = 0; c < 3; c++) {
        double s = 0;
        for (auto j = csc.col_offsets()[c]; j < csc.col_offsets()[c + 1]; j++) s += csc.values()[j];
        CHECK(std::fabs(s - 1.0) < 1e-10, "column %zu sums to %g after COLUMN normalisation", c, s);
    }
    auto csc2 = csc_from_triplets(3, 3, ri, ci, v);
    csc2.normalize(std::vector<double>{5.0, 4.0, 6.0}, 1.0, Direction::ROW);
    double rs[3] = {0, 0, 0};
    for (std::size_t j = 0; j < csc2.nnz(); j++) rs[csc2.row_indices()[j]] += csc2.values()[j];
    for (int r = 0; r < 3; r++) CHECK(std::fabs(rs[r] - 1.0) < 1e-10, "row %d sums to %g after ROW normalisation", r, rs[r]);
    auto z = csc_from_triplets(2, 2, {0, 1}, {0, 1}, {0.0, 0.0});
    z.log1p_normalize();
    for (double x : z.values()) CHECK(std::fabs(x) < 1e-10, "ln(1 + 0) = %g", x);
    std::printf("ok test_csc_twins\n");
}

// create_sparse_matrix of the reference's test module (pca/sparse/mod.rs:493-537): uniform random positions without
// repeats, values uniform in (-10, 10) away from zero, seed 42 (a different generator: the stream is not part of the test)
static CsrMatrix<double> create_sparse_matrix(std::size_t rows, std::size_t cols, double density) {
    std::mt19937_64 rng(42);
    const std::size_t per_row = (std::size_t)std::llround(cols * density) > 0 ? (std::size_t)std::llround(cols * density) : 1;
    std::vector<std::uint64_t> off(rows + 1, 0), idx;
    std::vector<double> val;
    idx.reserve(rows * per_row);
    val.reserve(rows * per_row);
    std::uniform_int_distribution<std::size_t> col(0, cols - 1);
    std::uniform_real_distribution<double> uv(-10.0, 10.0);
    for (std::size_t r = 0; r < rows; r++) {
        std::set<std::size_t> cs;
        while (cs.size() < per_row) cs.insert(col(rng));
        for (auto c : cs) {
            double v;
            do v = uv(rng); while (std::fabs(v) <= 1e-10);
            idx.push_back(c);
            val.push_back(v);
        }
        off[r + 1] = idx.size();
    }
    return CsrMatrix<double>(rows, cols, std::move(off), std::move(idx), std::move(val));
}

template <typename T>
static double max_offdiag_of_gram(const Array2<T>& c) {   // |C C^T - I|_max : rows of components_ are orthonormal
    double worst = 0;
    for (std::size_t i = 0; i < c.rows; i++)
        for (std::size_t j = i; j < c.rows; j++) {
            double d = 0;
            for (std::size_t k = 0; k < c.cols; k++) d += (double)c(i, k) * c(j, k);
            worst = std::fmax(worst, std::fabs(d - (i == j ? 1.0 : 0.0)));
        }
    return worst;
}

// pca/sparse/mod.rs:540-562
static void test_random_matrix_sparse_svd_comp_random(bool full) {
    const std::size_t rows = full ? 10000000 : 20000, cols = 2500;
    auto random_matrix = create_sparse_matrix(rows, cols, 0.01);
    auto sparse_pca = SparsePCABuilder<double>::new_()
                          .random_seed(42)
                          .svd_method(SVDMethod::Random(10, 7, PowerIterationNormalizer::QR))
                          .n_components(50)
                          .verbose(false)
                          .center(true)
                          .tolerance(1e-4)
                          .alpha(1.5)
                          .build();
    bool ok = true;
    random_matrix.device();     // upload outside the stopwatch
    const auto t0 = std::chrono::steady_clock::now();
    try {
        sparse_pca.fit(random_matrix);
    } catch (const Error& e) {
        ok = false;
        std::printf("fit failed: %s\n", e.what());
    }
    CHECK(ok, "res_fit.is_ok()");
    std::printf("fit of %zu x %zu (%zu stored entries, f64, Random{10, 7, QR}, 50 components): %.3f s, device-resident input\n", rows,
                cols, random_matrix.nnz(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    if (ok) {
        CHECK(sparse_pca.components_->rows == 50 && sparse_pca.components_->cols == cols, "components_ shape %zu x %zu",
              sparse_pca.components_->rows, sparse_pca.components_->cols);
        CHECK(sparse_pca.mean_->size() == cols, "mean_ length");
        const auto& ev = *sparse_pca.explained_variance_;
        for (std::size_t i = 1; i < ev.size(); i++) CHECK(ev[i] <= ev[i - 1] && ev[i] > 0, "explained_variance_ descending");
        CHECK(max_offdiag_of_gram(*sparse_pca.components_) < 1e-9, "components_ rows orthonormal: %g",
              max_offdiag_of_gram(*sparse_pca.components_));
        auto cum = sparse_pca.cumulative_explained_variance_ratio();
        CHECK(std::fabs(cum.back() - 1.0) < 1e-12, "cumulative ratio ends at 1 (sum over the computed components)");
        if (!full) {
            // transform of the fit rows equals fit_transform's scores
            auto t = sparse_pca.transform(random_matrix);
            auto s = sparse_pca.fit_transform(random_matrix);
            double worst = 0, scale = 0;
            for (std::size_t i = 0; i < t.data.size(); i++) {
                worst = std::fmax(worst, std::fabs(t.data[i] - s.data[i]));
                scale = std::fmax(scale, std::fabs(s.data[i]));
            }
            CHECK(worst <= 1e-9 * scale, "transform vs fit_transform: %g of %g", worst, scale);
        }
    }
    std::printf("ok test_random_matrix_sparse_svd_comp_random (%zu x %zu)\n", rows, cols);
}

// MaskedSparsePCA: mask semantics and the reference's error strings
static void test_masked() {
    const std::size_t rows = 4000, cols = 600;
    auto x = create_sparse_matrix(rows, cols, 0.05);
    std::vector<bool> mask(cols, false);
    for (std::size_t c = 0; c < cols; c += 3) mask[c] = true;       // 200 kept
    auto pca = MaskedSparsePCABuilder<double>::new_()
                   .n_components(10)
                   .mask(mask)
                   .svd_method(SVDMethod::Random(10, 4, PowerIterationNormalizer::LU))
                   .build();
    bool threw = false;
    try {
        pca.transform(x);
    } catch (const Error& e) {
        threw = std::string(e.what()) == "Must be fitted before transform!";
    }
    CHECK(threw, "transform before fit");
    threw = false;
    try {
        pca.explained_variance_ratio();
    } catch (const Error& e) {
        threw = std::string(e.what()) == "Model must be fitted first!";     // pca/sparse_masked/mod.rs:578
    }
    CHECK(threw, "explained_variance_ratio before fit");
    auto scores = pca.fit_transform(x);
    CHECK(scores.rows == rows && scores.cols == 10, "scores shape");
    CHECK(pca.components_->rows == 10 && pca.components_->cols == 200, "components_ cover the kept columns only");
    CHECK(pca.mean_->size() == cols, "mean_ has the FULL column count (pca/sparse_masked/mod.rs:280-291)");
    auto sums = x.sum_col();
    for (std::size_t c = 0; c < cols; c += 37)
        CHECK(std::fabs((*pca.mean_)[c] - sums[c] / rows) <= 1e-12 * (1 + std::fabs(sums[c] / rows)), "mean_[%zu]", c);
    // wrong mask length (pca/sparse_masked/mod.rs:258-262)
    auto bad = MaskedSparsePCABuilder<double>::new_().n_components(5).mask(std::vector<bool>(cols - 1, true))
                   .svd_method(SVDMethod::Random(5, 2)).build();
    threw = false;
    try {
        bad.fit(x);
    } catch (const Error& e) {
        threw = e.code == SALG_ERR_MASK_LEN &&
                std::string(e.what()) == "The mask vector length and the number of features (columns) have to be the same!";
        if (!threw) std::printf("message was: %s\n", e.what());
    }
    CHECK(threw, "mask length mismatch error string");
    // default method is Lanczos (pca/mod.rs:64-68)
    CHECK(SVDMethod() == SVDMethod::Lanczos(), "SVDMethod::default() is Lanczos");
    auto lz = SparsePCABuilder<double>::new_().n_components(5).build();
    lz.fit(x);
    CHECK(lz.components_->rows == 5 && lz.components_->cols == cols, "Lanczos components_ shape");
    std::printf("ok test_masked\n");
}

int main(int argc, char** argv) {
    const bool full = argc > 1 && std::strcmp(argv[1], "--full") == 0;
    if (Context::device_count() == 0) {
        std::printf("no CUDA device: libsalg_b200 has no CPU fallback\n");
        return 2;
    }
    try {
        test_csr_normalize();
        test_sums_and_log1p<float>("f32");
        test_sums_and_log1p<double>("f64");
        test_csc_twins();
        test_random_matrix_sparse_svd_comp_random(full);
        test_masked();
    } catch (const Error& e) {
        std::printf("FAIL uncaught Error %d: %s\n", e.code, e.what());
        return 1;
    }
    if (g_fail) {
        std::printf("%d check(s) failed\n", g_fail);
        return 1;
    }
    std::printf("all facade tests passed\n");
    return 0;
}
