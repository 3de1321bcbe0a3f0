// single_algebra.hpp — C++17 host facade over the C ABI of libsalg_b200.so (include/salg.h).
//
// The reference (crate `single_algebra` 0.9.2) is compiled Rust with no FFI of its own; its operator API on the sparse-PCA
// hot path is the type surface
//   single_algebra::dimred::pca::{SparsePCA, SparsePCABuilder, MaskedSparsePCA, MaskedSparsePCABuilder, SVDMethod,
//                                 PowerIterationNormalizer}                      (src/dimred/pca/mod.rs:37-68)
//   single_algebra::sparse::MatrixSum  {sum_col, sum_col_squared, sum_row}        (src/sparse/mod.rs:67-102)
//   single_algebra::{Normalize, Log1P} {normalize, log1p_normalize}              (src/utils/mod.rs:6-17)
// The Rust toolchain is not in this image, so the facade a maintainer would write in Rust is mirrored here
// in C++: same names, same argument meaning, same defaults, same error strings.  Every method body is one or two calls into
// the C ABI; nothing is computed on the host except what the reference itself computes on tiny arrays (ratios, prefix sums).
// There is no CPU fallback: without a CUDA device every compute call throws Error{SALG_ERR_CUDA}.
//
// Header only.  Build:  g++ -std=c++17 -I include app.cpp -L single-algebra_b200 -lsalg_b200 -Wl,-rpath,<dir of the .so>
#ifndef SINGLE_ALGEBRA_HPP
#define SINGLE_ALGEBRA_HPP

#include <cstddef>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "salg.h"

namespace single_algebra {

// anyhow::Error of the reference: message + the C ABI's status code
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

namespace detail {
inline void check(int status) {
    if (status != SALG_OK) {
        const char* m = salg_last_error();
        throw Error(status, m && *m ? std::string(m) : "salg status " + std::to_string(status));
    }
}
// the two instantiations of the reference's generic T
template <typename T> struct Abi;
template <> struct Abi<float> {
    static constexpr auto csr_upload = salg_csr_upload_f32;
    static constexpr auto csr_download = salg_csr_download_f32;
    static constexpr auto sum_col = salg_sum_col_f32;
    static constexpr auto sum_row = salg_sum_row_f32;
    static constexpr auto normalize = salg_normalize_f32;
    static constexpr auto csc_upload = salg_csc_upload_f32;
    static constexpr auto csc_sum_col = salg_csc_sum_col_f32;
    static constexpr auto csc_sum_row = salg_csc_sum_row_f32;
    static constexpr auto csc_normalize = salg_csc_normalize_f32;
    static constexpr auto pca_fit = salg_pca_fit_f32;
    static constexpr auto pca_components = salg_pca_components_f32;
    static constexpr auto pca_transform = salg_pca_transform_f32;
    static constexpr auto pca_fit_scores = salg_pca_fit_scores_f32;
};
template <> struct Abi<double> {
    static constexpr auto csr_upload = salg_csr_upload_f64;
    static constexpr auto csr_download = salg_csr_download_f64;
    static constexpr auto sum_col = salg_sum_col_f64;
    static constexpr auto sum_row = salg_sum_row_f64;
    static constexpr auto normalize = salg_normalize_f64;
    static constexpr auto csc_upload = salg_csc_upload_f64;
    static constexpr auto csc_sum_col = salg_csc_sum_col_f64;
    static constexpr auto csc_sum_row = salg_csc_sum_row_f64;
    static constexpr auto csc_normalize = salg_csc_normalize_f64;
    static constexpr auto pca_fit = salg_pca_fit_f64;
    static constexpr auto pca_components = salg_pca_components_f64;
    static constexpr auto pca_transform = salg_pca_transform_f64;
    static constexpr auto pca_fit_scores = salg_pca_fit_scores_f64;
};
}  // namespace detail

// One GPU + stream + workspaces.  The reference's analogue is the ambient Rayon pool; here the default context is created
// on device 0 at first use and lives for the process.
class Context {
public:
    explicit Context(int device = 0) { detail::check(salg_ctx_create(device, &h_)); }
    ~Context() { if (h_) salg_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    salg_ctx* handle() const { return h_; }
    void sync() const { detail::check(salg_ctx_sync(h_)); }
    static Context& default_context() {
        static Context ctx(0);
        return ctx;
    }
    static int device_count() {
        int n = 0;
        return salg_device_count(&n) == SALG_OK ? n : 0;
    }

private:
    salg_ctx* h_ = nullptr;
};

// single_utilities::types::Direction as used by Normalize::normalize (src/sparse/csr.rs:1032, 1046)
enum class Direction { ROW = SALG_ROW, COLUMN = SALG_COLUMN };

// single_svdlib::randomized::PowerIterationNormalizer (re-export src/dimred/pca/mod.rs:41)
enum class PowerIterationNormalizer { QR = SALG_NORM_QR, LU = SALG_NORM_LU, None = SALG_NORM_NONE };

// SVDMethod (src/dimred/pca/mod.rs:49-62); Default = Lanczos (:64-68)
struct SVDMethod {
    enum Kind { KLanczos = SALG_SVD_LANCZOS, KRandom = SALG_SVD_RANDOM } kind = KLanczos;
    std::size_t n_oversamples = 0;
    std::size_t n_power_iterations = 0;
    PowerIterationNormalizer normalizer = PowerIterationNormalizer::QR;
    static SVDMethod Lanczos() { return SVDMethod{}; }
    static SVDMethod Random(std::size_t n_oversamples, std::size_t n_power_iterations,
                            PowerIterationNormalizer normalizer = PowerIterationNormalizer::QR) {
        SVDMethod m;
        m.kind = KRandom;
        m.n_oversamples = n_oversamples;
        m.n_power_iterations = n_power_iterations;
        m.normalizer = normalizer;
        return m;
    }
    bool operator==(const SVDMethod& o) const {
        return kind == o.kind && (kind == KLanczos || (n_oversamples == o.n_oversamples &&
                                                       n_power_iterations == o.n_power_iterations && normalizer == o.normalizer));
    }
};

// what SparsePCA::transform / MaskedSparsePCA::transform compute
enum class TransformMode { Exact = SALG_TRANSFORM_EXACT, ReferenceCompat = SALG_TRANSFORM_REFERENCE_COMPAT };

// nalgebra_sparse::CsrMatrix<T>: usize row offsets [nrows + 1], usize column indices [nnz] strictly increasing inside a
// row, values T [nnz].  The host arrays are the matrix; a device copy is made at first use and dropped whenever the host
// values are handed out mutably, so `&mut self` semantics of normalize / log1p_normalize hold on the host copy.
template <typename T>
class CsrMatrix {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "T is f32 or f64");

public:
    // CsrMatrix::try_from_csr_data; the layout is validated on the device at first use (offsets monotone and ending at nnz,
    // indices strictly increasing per row and < ncols), an invalid matrix throws Error{SALG_ERR_BAD_ARG} there
    CsrMatrix(std::size_t nrows, std::size_t ncols, std::vector<std::uint64_t> row_offsets,
              std::vector<std::uint64_t> col_indices, std::vector<T> values, Context* ctx = nullptr)
        : nrows_(nrows), ncols_(ncols), off_(std::move(row_offsets)), idx_(std::move(col_indices)), val_(std::move(values)),
          ctx_(ctx) {
        if (off_.size() != nrows_ + 1) throw Error(SALG_ERR_BAD_ARG, "row_offsets must have nrows + 1 entries");
        if (idx_.size() != val_.size()) throw Error(SALG_ERR_BAD_ARG, "col_indices and values differ in length");
    }
    ~CsrMatrix() { drop_device(); }
    CsrMatrix(const CsrMatrix&) = delete;
    CsrMatrix& operator=(const CsrMatrix&) = delete;
    CsrMatrix(CsrMatrix&& o) noexcept { *this = std::move(o); }
    CsrMatrix& operator=(CsrMatrix&& o) noexcept {
        if (this != &o) {
            drop_device();
            nrows_ = o.nrows_; ncols_ = o.ncols_;
            off_ = std::move(o.off_); idx_ = std::move(o.idx_); val_ = std::move(o.val_);
            ctx_ = o.ctx_; dev_ = o.dev_; o.dev_ = nullptr;
        }
        return *this;
    }

    std::size_t nrows() const { return nrows_; }
    std::size_t ncols() const { return ncols_; }
    std::size_t nnz() const { return val_.size(); }
    const std::vector<std::uint64_t>& row_offsets() const { return off_; }
    const std::vector<std::uint64_t>& col_indices() const { return idx_; }
    const std::vector<T>& values() const { return val_; }
    std::vector<T>& values_mut() {
        drop_device();
        return val_;
    }

    // ---- MatrixSum (src/sparse/mod.rs:67-102) ----
    // sum_col (src/sparse/csr.rs:259-312)
    std::vector<T> sum_col() const {
        std::vector<T> s(ncols_);
        detail::check(detail::Abi<T>::sum_col(ctx().handle(), device(), s.data(), nullptr));
        return s;
    }
    // sum_col_squared (src/sparse/csr.rs:558-608)
    std::vector<T> sum_col_squared() const {
        std::vector<T> s(ncols_), q(ncols_);
        detail::check(detail::Abi<T>::sum_col(ctx().handle(), device(), s.data(), q.data()));
        return q;
    }
    // sum_row (src/sparse/csr.rs:314-392)
    std::vector<T> sum_row() const {
        std::vector<T> s(nrows_);
        detail::check(detail::Abi<T>::sum_row(ctx().handle(), device(), s.data()));
        return s;
    }

    // ---- Normalize::normalize (src/sparse/csr.rs:1013-1068; trait src/utils/mod.rs:6-13), U = T ----
    // scale[i] = sums[i] > 0 ? target / sums[i] : 0; entries are rescaled only where scale > 0.  A `sums` shorter than the
    // normalised dimension is an Error (the reference indexes it unchecked and panics).
    void normalize(const std::vector<T>& sums, T target, Direction direction) {
        detail::check(detail::Abi<T>::normalize(ctx().handle(), device(), sums.data(), (std::int64_t)sums.size(), target,
                                                (int)direction));
        pull_values();
    }
    // U = f64 sums on an f32 matrix (the reference is generic over U)
    template <typename TT = T, typename = std::enable_if_t<std::is_same<TT, float>::value>>
    void normalize(const std::vector<double>& sums, double target, Direction direction) {
        detail::check(salg_normalize_f32_u64(ctx().handle(), device(), sums.data(), (std::int64_t)sums.size(), target,
                                             (int)direction));
        pull_values();
    }
    // ---- Log1P::log1p_normalize (src/sparse/csr.rs:1070-1079): v <- ln(fl(1 + v)) ----
    void log1p_normalize() {
        detail::check(salg_log1p(ctx().handle(), device()));
        pull_values();
    }

    Context& ctx() const { return ctx_ ? *ctx_ : Context::default_context(); }
    // device-resident copy (uploaded at first use)
    salg_csr* device() const {
        if (!dev_) {
            detail::check(detail::Abi<T>::csr_upload(ctx().handle(), (std::int64_t)nrows_, (std::int64_t)ncols_,
                                                     (std::int64_t)val_.size(), off_.data(), idx_.data(), val_.data(),
                                                     &dev_));
        }
        return dev_;
    }
    void drop_device() const {
        if (dev_) {
            salg_csr_free(dev_);
            dev_ = nullptr;
        }
    }

private:
    void pull_values() {      // the in-place operations ran on the device copy: bring the values back (&mut self)
        detail::check(detail::Abi<T>::csr_download(ctx().handle(), dev_, nullptr, nullptr, val_.data()));
    }
    std::size_t nrows_ = 0, ncols_ = 0;
    std::vector<std::uint64_t> off_, idx_;
    std::vector<T> val_;
    Context* ctx_ = nullptr;
    mutable salg_csr* dev_ = nullptr;
};

// nalgebra_sparse::CscMatrix<T>: usize column offsets [ncols + 1], usize row indices [nnz] strictly increasing inside a
// column, values T [nnz] — the CSC twins of the traits.  The device keeps it as the CSR of A^T.
template <typename T>
class CscMatrix {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "T is f32 or f64");

public:
    CscMatrix(std::size_t nrows, std::size_t ncols, std::vector<std::uint64_t> col_offsets,
              std::vector<std::uint64_t> row_indices, std::vector<T> values, Context* ctx = nullptr)
        : nrows_(nrows), ncols_(ncols), off_(std::move(col_offsets)), idx_(std::move(row_indices)), val_(std::move(values)),
          ctx_(ctx) {
        if (off_.size() != ncols_ + 1) throw Error(SALG_ERR_BAD_ARG, "col_offsets must have ncols + 1 entries");
        if (idx_.size() != val_.size()) throw Error(SALG_ERR_BAD_ARG, "row_indices and values differ in length");
    }
    ~CscMatrix() { drop_device(); }
    CscMatrix(const CscMatrix&) = delete;
    CscMatrix& operator=(const CscMatrix&) = delete;
    CscMatrix(CscMatrix&& o) noexcept
        : nrows_(o.nrows_), ncols_(o.ncols_), off_(std::move(o.off_)), idx_(std::move(o.idx_)), val_(std::move(o.val_)),
          ctx_(o.ctx_), dev_(o.dev_) {
        o.dev_ = nullptr;
    }

    std::size_t nrows() const { return nrows_; }
    std::size_t ncols() const { return ncols_; }
    std::size_t nnz() const { return val_.size(); }
    const std::vector<std::uint64_t>& col_offsets() const { return off_; }
    const std::vector<std::uint64_t>& row_indices() const { return idx_; }
    const std::vector<T>& values() const { return val_; }
    std::vector<T>& values_mut() {
        drop_device();
        return val_;
    }

    // MatrixSum for CscMatrix: sum_col (src/sparse/csc.rs:157-197), sum_col_squared (:323-335), sum_row (:199-220)
    std::vector<T> sum_col() const {
        std::vector<T> s(ncols_);
        detail::check(detail::Abi<T>::csc_sum_col(ctx().handle(), device(), s.data(), nullptr));
        return s;
    }
    std::vector<T> sum_col_squared() const {
        std::vector<T> q(ncols_);
        detail::check(detail::Abi<T>::csc_sum_col(ctx().handle(), device(), nullptr, q.data()));
        return q;
    }
    std::vector<T> sum_row() const {
        std::vector<T> s(nrows_);
        detail::check(detail::Abi<T>::csc_sum_row(ctx().handle(), device(), s.data()));
        return s;
    }
    // Normalize::normalize for CscMatrix (src/sparse/csc.rs:680-735): ROW indexes `sums` by row_indices, COLUMN by column
    void normalize(const std::vector<T>& sums, T target, Direction direction) {
        detail::check(detail::Abi<T>::csc_normalize(ctx().handle(), device(), sums.data(), (std::int64_t)sums.size(), target,
                                                    (int)direction));
        pull_values();
    }
    // Log1P::log1p_normalize for CscMatrix (src/sparse/csc.rs:737-746)
    void log1p_normalize() {
        detail::check(salg_log1p(ctx().handle(), device()));
        pull_values();
    }

    Context& ctx() const { return ctx_ ? *ctx_ : Context::default_context(); }
    salg_csr* device() const {
        if (!dev_) {
            detail::check(detail::Abi<T>::csc_upload(ctx().handle(), (std::int64_t)nrows_, (std::int64_t)ncols_,
                                                     (std::int64_t)val_.size(), off_.data(), idx_.data(), val_.data(),
                                                     &dev_));
        }
        return dev_;
    }
    void drop_device() const {
        if (dev_) {
            salg_csr_free(dev_);
            dev_ = nullptr;
        }
    }

private:
    void pull_values() {
        detail::check(detail::Abi<T>::csr_download(ctx().handle(), dev_, nullptr, nullptr, val_.data()));
    }
    std::size_t nrows_ = 0, ncols_ = 0;
    std::vector<std::uint64_t> off_, idx_;
    std::vector<T> val_;
    Context* ctx_ = nullptr;
    mutable salg_csr* dev_ = nullptr;
};

// row-major dense matrix returned by the PCA types (ndarray::Array2<T> in the reference)
template <typename T>
struct Array2 {
    std::size_t rows = 0, cols = 0;
    std::vector<T> data;
    T& operator()(std::size_t r, std::size_t c) { return data[r * cols + c]; }
    const T& operator()(std::size_t r, std::size_t c) const { return data[r * cols + c]; }
};

namespace detail {
// state and methods SparsePCA and MaskedSparsePCA share (the reference duplicates them in two files)
template <typename T>
class PcaBase {
public:
    ~PcaBase() { free_model(); }
    PcaBase(const PcaBase&) = delete;
    PcaBase& operator=(const PcaBase&) = delete;
    PcaBase(PcaBase&& o) noexcept { move_from(std::move(o)); }
    PcaBase& operator=(PcaBase&& o) noexcept {
        if (this != &o) {
            free_model();
            move_from(std::move(o));
        }
        return *this;
    }

    // fitted state: components_ d x n_features (pca/sparse/mod.rs:208), explained_variance_ (:210-216), mean_ with the
    // FULL column count (masked: pca/sparse_masked/mod.rs:280-291)
    std::optional<Array2<T>> components_;
    std::optional<std::vector<T>> explained_variance_;
    std::optional<std::vector<T>> mean_;
    std::vector<double> singular_values_;
    double total_var_ = 0.0;

    // `transform(&self, x)` (pca/sparse/mod.rs:255-285, pca/sparse_masked/mod.rs:438-546): n x d scores.
    // Default: the projection (X - 1 mu^T) V^T; TransformMode::ReferenceCompat reproduces the reference's loops bug for
    // bug.
    Array2<T> transform(const CsrMatrix<T>& x, TransformMode mode = TransformMode::Exact) const {
        if (masked_ && x.ncols() != mask_len_)      // checked before the fitted state (pca/sparse_masked/mod.rs:440-444)
            throw Error(SALG_ERR_MASK_LEN, "The mask vector length and the number of features (columns) have to be the same!");
        if (!model_) throw Error(SALG_ERR_NOT_FITTED, "Must be fitted before transform!");
        Array2<T> out{x.nrows(), components_->rows, std::vector<T>(x.nrows() * components_->rows)};
        check(Abi<T>::pca_transform(x.ctx().handle(), model_, x.device(), (int)mode, out.data.data()));
        return out;
    }
    // `feature_importances` (pca/sparse/mod.rs:295-302): squared loadings
    Array2<T> feature_importances() const {
        if (!components_) throw Error(SALG_ERR_NOT_FITTED, "Model must be fitted first!");          // :299
        Array2<T> out = *components_;
        for (auto& v : out.data) v = v * v;
        return out;
    }
    // `explained_variance_ratio` (pca/sparse/mod.rs:312-322): normalised by the sum over the COMPUTED components
    std::vector<T> explained_variance_ratio() const {
        if (!explained_variance_) throw Error(SALG_ERR_NOT_FITTED, "Model must be fitted first!");  // :316
        T total = T(0);
        for (T v : *explained_variance_) total += v;
        std::vector<T> r = *explained_variance_;
        for (auto& v : r) v /= total;
        return r;
    }
    // `cumulative_explained_variance_ratio` (pca/sparse/mod.rs:333-343)
    std::vector<T> cumulative_explained_variance_ratio() const {
        std::vector<T> r = explained_variance_ratio();
        for (std::size_t i = 1; i < r.size(); i++) r[i] += r[i - 1];
        return r;
    }
    // bit 1: a Cholesky pivot was floored; bit 2: Jacobi sweep limit; bit 4: Lanczos returned early (salg.h)
    int numeric_flags() const {
        int f = 0;
        if (model_) check(salg_pca_numeric_flags(model_, &f));
        return f;
    }

protected:
    PcaBase() = default;
    // one salg_pca_fit call; `omega` (extension, nullable): the host Gaussian test matrix shared with the oracle, row-major
    // omega_rows x omega_cols — the "same host-generated Omega" of the parity contract
    void fit_impl(const CsrMatrix<T>& x, const std::vector<bool>* mask, bool keep_scores, const T* omega,
                  std::size_t omega_rows, std::size_t omega_cols) {
        salg_pca_params p;
        check(salg_pca_params_default(&p));
        p.n_components = (std::int32_t)n_components;
        p.svd_method = (std::int32_t)svdmethod.kind;
        p.n_oversamples = (std::int32_t)svdmethod.n_oversamples;
        p.n_power_iterations = (std::int32_t)svdmethod.n_power_iterations;
        p.normalizer = (std::int32_t)svdmethod.normalizer;
        p.center = center ? 1 : 0;
        p.verbose = verbose ? 1 : 0;
        p.random_seed = random_seed;
        p.alpha = (double)alpha;
        p.tolerance = (double)tolerance;
        p.keep_scores = keep_scores ? 1 : 0;
        std::vector<std::uint8_t> m8;
        if (mask) m8.assign(mask->begin(), mask->end());
        free_model();
        components_.reset();
        explained_variance_.reset();
        mean_.reset();
        check(Abi<T>::pca_fit(x.ctx().handle(), x.device(), &p, mask ? m8.data() : nullptr, mask ? (std::int64_t)m8.size() : 0,
                              omega, (std::int64_t)omega_rows, (std::int64_t)omega_cols, &model_));
        ctx_ = x.ctx().handle();
        std::int64_t d = 0, n_eff = 0, ncols = 0;
        int dt = 0;
        check(salg_pca_dims(model_, &d, &n_eff, &ncols, &dt));
        Array2<T> comp{(std::size_t)d, (std::size_t)n_eff, std::vector<T>((std::size_t)(d * n_eff))};
        check(Abi<T>::pca_components(model_, comp.data.data()));
        std::vector<double> ev((std::size_t)d), mean((std::size_t)ncols);
        singular_values_.assign((std::size_t)d, 0.0);
        check(salg_pca_singular_values_f64(model_, singular_values_.data()));
        check(salg_pca_explained_variance_f64(model_, ev.data()));
        check(salg_pca_mean_f64(model_, mean.data()));
        check(salg_pca_total_var(model_, &total_var_));
        components_ = std::move(comp);
        explained_variance_ = std::vector<T>(ev.begin(), ev.end());
        mean_ = std::vector<T>(mean.begin(), mean.end());
    }
    Array2<T> fit_scores(std::size_t nrows) {
        Array2<T> out{nrows, components_->rows, std::vector<T>(nrows * components_->rows)};
        check(Abi<T>::pca_fit_scores(ctx_, model_, out.data.data()));
        return out;
    }
    void free_model() {
        if (model_) salg_pca_free(model_);
        model_ = nullptr;
    }
    void move_from(PcaBase&& o) {
        components_ = std::move(o.components_);
        explained_variance_ = std::move(o.explained_variance_);
        mean_ = std::move(o.mean_);
        singular_values_ = std::move(o.singular_values_);
        total_var_ = o.total_var_;
        n_components = o.n_components; alpha = o.alpha; tolerance = o.tolerance; random_seed = o.random_seed;
        center = o.center; verbose = o.verbose; svdmethod = o.svdmethod;
        model_ = o.model_; ctx_ = o.ctx_;
        masked_ = o.masked_; mask_len_ = o.mask_len_;
        o.model_ = nullptr;
    }

    std::size_t n_components = 50;
    T alpha = T(1);
    T tolerance = T(1e-6);
    std::uint32_t random_seed = 42;
    bool center = true, verbose = false;
    SVDMethod svdmethod;
    salg_pca* model_ = nullptr;
    salg_ctx* ctx_ = nullptr;
    bool masked_ = false;            // MaskedSparsePCA: transform checks the mask length first
    std::size_t mask_len_ = 0;
};
}  // namespace detail

// ---- SparsePCA<T> (src/dimred/pca/sparse/mod.rs:33-358) ----
template <typename T> class SparsePCABuilder;
template <typename T>
class SparsePCA : public detail::PcaBase<T> {
public:
    // SparsePCA::new (pca/sparse/mod.rs:63-84); `None` for tolerance / seed takes 1e-6 / 42
    SparsePCA(std::size_t n_components, T alpha, std::optional<T> tollerance, std::optional<std::uint32_t> random_seed,
              bool center, bool verbose, SVDMethod svdmethod) {
        this->n_components = n_components;
        this->alpha = alpha;
        this->tolerance = tollerance.value_or(T(1e-6));
        this->random_seed = random_seed.value_or(42u);
        this->center = center;
        this->verbose = verbose;
        this->svdmethod = svdmethod;
    }
    // `fit(&mut self, x)` (pca/sparse/mod.rs:102-242)
    SparsePCA& fit(const CsrMatrix<T>& x, const T* omega = nullptr, std::size_t omega_rows = 0, std::size_t omega_cols = 0) {
        this->fit_impl(x, nullptr, false, omega, omega_rows, omega_cols);
        return *this;
    }
    // `fit_transform(&mut self, x)` (pca/sparse/mod.rs:355-358): fit, then the projection of the same rows (computed inside
    // the fit call while the operator is resident)
    Array2<T> fit_transform(const CsrMatrix<T>& x, const T* omega = nullptr, std::size_t omega_rows = 0,
                            std::size_t omega_cols = 0, TransformMode mode = TransformMode::Exact) {
        if (mode != TransformMode::Exact) {          // fit + transform literally: the projection kept by the fit is the exact one
            this->fit_impl(x, nullptr, false, omega, omega_rows, omega_cols);
            return this->transform(x, mode);
        }
        this->fit_impl(x, nullptr, true, omega, omega_rows, omega_cols);
        return this->fit_scores(x.nrows());
    }
};

// SparsePCABuilder (pca/sparse/mod.rs:375-484); defaults :388-403
template <typename T>
class SparsePCABuilder {
public:
    SparsePCABuilder() = default;
    static SparsePCABuilder new_() { return SparsePCABuilder(); }
    SparsePCABuilder& n_components(std::size_t n) { n_components_ = n; return *this; }
    SparsePCABuilder& alpha(T a) { alpha_ = a; return *this; }
    SparsePCABuilder& tolerance(T t) { tolerance_ = t; return *this; }
    SparsePCABuilder& random_seed(std::uint32_t s) { random_seed_ = s; return *this; }
    SparsePCABuilder& center(bool c) { center_ = c; return *this; }
    SparsePCABuilder& verbose(bool v) { verbose_ = v; return *this; }
    SparsePCABuilder& svd_method(SVDMethod m) { svdmethod_ = m; return *this; }
    SparsePCA<T> build() const {
        return SparsePCA<T>(n_components_, alpha_, tolerance_, random_seed_, center_, verbose_, svdmethod_);
    }

private:
    std::size_t n_components_ = 50;
    T alpha_ = T(1);
    T tolerance_ = T(1e-6);
    std::uint32_t random_seed_ = 42;
    bool center_ = true, verbose_ = false;
    SVDMethod svdmethod_;
};

// ---- MaskedSparsePCA<T> (src/dimred/pca/sparse_masked/mod.rs:179-619) ----
template <typename T>
class MaskedSparsePCA : public detail::PcaBase<T> {
public:
    // MaskedSparsePCA::new (pca/sparse_masked/mod.rs:214-237)
    MaskedSparsePCA(std::size_t n_components, T alpha, std::optional<T> tolerance, std::optional<std::uint32_t> random_seed,
                    bool center, bool verbose, std::vector<bool> mask, SVDMethod svdmethod)
        : mask_(std::move(mask)) {
        this->masked_ = true;
        this->mask_len_ = mask_.size();
        this->n_components = n_components;
        this->alpha = alpha;
        this->tolerance = tolerance.value_or(T(1e-6));
        this->random_seed = random_seed.value_or(42u);
        this->center = center;
        this->verbose = verbose;
        this->svdmethod = svdmethod;
    }
    // `fit(&mut self, x)` (pca/sparse_masked/mod.rs:255-419): a mask whose length differs from x.ncols() is the reference's
    // "The mask vector length and the number of features (columns) have to be the same!" (:258-262)
    MaskedSparsePCA& fit(const CsrMatrix<T>& x, const T* omega = nullptr, std::size_t omega_rows = 0,
                         std::size_t omega_cols = 0) {
        this->fit_impl(x, &mask_, false, omega, omega_rows, omega_cols);
        return *this;
    }
    // `fit_transform(&mut self, x)` (pca/sparse_masked/mod.rs:605-619)
    Array2<T> fit_transform(const CsrMatrix<T>& x, const T* omega = nullptr, std::size_t omega_rows = 0,
                            std::size_t omega_cols = 0, TransformMode mode = TransformMode::Exact) {
        if (mode != TransformMode::Exact) {
            this->fit_impl(x, &mask_, false, omega, omega_rows, omega_cols);
            return this->transform(x, mode);
        }
        this->fit_impl(x, &mask_, true, omega, omega_rows, omega_cols);
        return this->fit_scores(x.nrows());
    }
    const std::vector<bool>& mask() const { return mask_; }

private:
    std::vector<bool> mask_;
};

// MaskedSparsePCABuilder (pca/sparse_masked/mod.rs:37-160); defaults :51-67
template <typename T>
class MaskedSparsePCABuilder {
public:
    MaskedSparsePCABuilder() = default;
    static MaskedSparsePCABuilder new_() { return MaskedSparsePCABuilder(); }
    MaskedSparsePCABuilder& n_components(std::size_t n) { n_components_ = n; return *this; }
    MaskedSparsePCABuilder& alpha(T a) { alpha_ = a; return *this; }
    MaskedSparsePCABuilder& tolerance(T t) { tolerance_ = t; return *this; }
    MaskedSparsePCABuilder& random_seed(std::uint32_t s) { random_seed_ = s; return *this; }
    MaskedSparsePCABuilder& center(bool c) { center_ = c; return *this; }
    MaskedSparsePCABuilder& verbose(bool v) { verbose_ = v; return *this; }
    MaskedSparsePCABuilder& mask(std::vector<bool> m) { mask_ = std::move(m); return *this; }
    MaskedSparsePCABuilder& svd_method(SVDMethod m) { svdmethod_ = m; return *this; }
    MaskedSparsePCA<T> build() const {
        return MaskedSparsePCA<T>(n_components_, alpha_, tolerance_, random_seed_, center_, verbose_, mask_, svdmethod_);
    }

private:
    std::size_t n_components_ = 50;
    T alpha_ = T(1);
    T tolerance_ = T(1e-6);
    std::uint32_t random_seed_ = 42;
    bool center_ = true, verbose_ = false;
    std::vector<bool> mask_;
    SVDMethod svdmethod_;
};

}  // namespace single_algebra
#endif  // SINGLE_ALGEBRA_HPP
