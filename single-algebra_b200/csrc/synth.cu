// synth.cu — device-side synthetic count-matrix generator (benchmark input; not a reference function).
// Bit-identical to single-algebra_b200/synth.py::generate_rows for the same tables: all per-cell work is
// 64-bit integer hashing and integer comparisons against host-built tables, so any row range of the
// 1M x 30k / 4M x 33k benchmark matrices can be regenerated on the host for a parity spot check
// (SURVEY §7 hard part 6).
#include "common.cuh"

namespace salg {

constexpr int SY_LEVELS = 256;
constexpr int SY_KCDF = 40;
constexpr int SY_NSF = 16;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}

struct SynthArgs {
    uint64_t seed;
    int64_t row0, nrows, ncols;
    int n_clusters;
    const uint8_t* base_level;   // [n_clusters * ncols]
    const int32_t* sf_offset;    // [16]
    const uint32_t* cdf;         // [256 * 40]
};

__device__ __forceinline__ void row_meta(const SynthArgs& a, int64_t row, int& cluster, int& sfo, uint64_t& hrow) {
    uint64_t r = (uint64_t)row;
    uint64_t hr = mix64((a.seed ^ 0xA5A5A5A5DEADBEEFULL) + r * 0x9E3779B97F4A7C15ULL);
    cluster = (int)((hr & 0xFFFFULL) % (uint64_t)a.n_clusters);
    int sf = (int)((hr >> 16) & (uint64_t)(SY_NSF - 1));
    sfo = a.sf_offset[sf];
    hrow = mix64(a.seed + r * 0x9E3779B97F4A7C15ULL);
}

// value of cell (row, col): 0 when not stored
__device__ __forceinline__ int cell_value(const SynthArgs& a, const uint8_t* __restrict__ lvl_row, int sfo,
                                          uint64_t hrow, int64_t col) {
    uint64_t h = mix64(hrow ^ ((uint64_t)col * 0xD1B54A32D192ED03ULL + 0x8CB92BA72F3D8DD7ULL));
    uint32_t u = (uint32_t)(h >> 32);
    int lvl = (int)lvl_row[col] + sfo;
    lvl = lvl < 0 ? 0 : (lvl > SY_LEVELS - 1 ? SY_LEVELS - 1 : lvl);
    const uint32_t* c = a.cdf + lvl * SY_KCDF;
    if (u < c[0]) return 0;
    int x = 1;
    while (x < SY_KCDF && u >= c[x]) x++;
    return x;
}

__global__ void synth_count_kernel(SynthArgs a, int64_t* __restrict__ cnt) {
    int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < a.nrows; r += nw) {
        int cluster, sfo;
        uint64_t hrow;
        row_meta(a, a.row0 + r, cluster, sfo, hrow);
        const uint8_t* lvl_row = a.base_level + (size_t)cluster * a.ncols;
        int n = 0;
        for (int64_t c = lane; c < a.ncols; c += 32) n += cell_value(a, lvl_row, sfo, hrow, c) != 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
        if (lane == 0) cnt[r] = n;
    }
    if (w == 0 && lane == 0) cnt[a.nrows] = 0;
}

template <typename T>
__global__ void synth_fill_kernel(SynthArgs a, const int64_t* __restrict__ ptr, uint32_t* __restrict__ col,
                                  T* __restrict__ val) {
    int lane = threadIdx.x & 31;
    unsigned lt = (1u << lane) - 1u;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w; r < a.nrows; r += nw) {
        int cluster, sfo;
        uint64_t hrow;
        row_meta(a, a.row0 + r, cluster, sfo, hrow);
        const uint8_t* lvl_row = a.base_level + (size_t)cluster * a.ncols;
        int64_t o = ptr[r];
        for (int64_t base = 0; base < a.ncols; base += 32) {
            int64_t c = base + lane;
            int x = 0;
            if (c < a.ncols) x = cell_value(a, lvl_row, sfo, hrow, c);
            unsigned b = __ballot_sync(0xFFFFFFFFu, x != 0);
            if (x != 0) {
                int64_t q = o + __popc(b & lt);
                col[q] = (uint32_t)c;
                val[q] = (T)x;
            }
            o += __popc(b);
        }
    }
}

}  // namespace salg

using namespace salg;

extern "C" int salg_csr_synth(salg_ctx* ctx, int dtype, uint64_t seed, int64_t row0, int64_t nrows, int64_t ncols,
                              int32_t n_clusters, const uint8_t* base_level, const int32_t* sf_offset,
                              const uint32_t* cdf, salg_csr** out) {
    return guarded([&] {
        SALG_REQUIRE(ctx && out && base_level && sf_offset && cdf, SALG_ERR_BAD_ARG, "NULL argument");
        SALG_REQUIRE(dtype == SALG_F32 || dtype == SALG_F64, SALG_ERR_BAD_ARG, "bad dtype");
        SALG_REQUIRE(nrows >= 0 && ncols >= 1 && n_clusters >= 1 && row0 >= 0, SALG_ERR_BAD_ARG, "bad shape");
        SALG_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        DevBuf<uint8_t> d_lvl((size_t)n_clusters * ncols, st);
        DevBuf<int32_t> d_sf(SY_NSF, st);
        DevBuf<uint32_t> d_cdf(SY_LEVELS * SY_KCDF, st);
        SALG_CUDA(cudaMemcpyAsync(d_lvl.get(), base_level, (size_t)n_clusters * ncols, cudaMemcpyHostToDevice, st));
        SALG_CUDA(cudaMemcpyAsync(d_sf.get(), sf_offset, SY_NSF * 4, cudaMemcpyHostToDevice, st));
        SALG_CUDA(cudaMemcpyAsync(d_cdf.get(), cdf, SY_LEVELS * SY_KCDF * 4, cudaMemcpyHostToDevice, st));
        SynthArgs a{seed, row0, nrows, ncols, n_clusters, d_lvl.get(), d_sf.get(), d_cdf.get()};
        DevBuf<int64_t> cnt((size_t)nrows + 1, st);
        int64_t want = ceil_div((nrows + 1) * 32, 256);
        int64_t cap = (int64_t)ctx->sm_count * 8;
        unsigned grid = (unsigned)(want < cap ? want : cap);
        synth_count_kernel<<<grid, 256, 0, st>>>(a, cnt.get());
        ctx->n_launch++;
        SALG_CUDA(cudaGetLastError());
        int64_t* ptr = (int64_t*)dev_alloc(ctx, (size_t)(nrows + 1) * 8);
        salg_csr* c = nullptr;
        try {
            exclusive_scan_i64(ctx, cnt.get(), ptr, nrows + 1);
            int64_t nnz = 0;
            SALG_CUDA(cudaMemcpyAsync(&nnz, ptr + nrows, 8, cudaMemcpyDeviceToHost, st));
            SALG_CUDA(cudaStreamSynchronize(st));
            c = new salg_csr();
            c->ctx = ctx;
            c->dtype = dtype;
            c->nrows = nrows;
            c->ncols = ncols;
            c->nnz = nnz;
            c->row_ptr = ptr;
            ptr = nullptr;
            size_t es = dtype == SALG_F64 ? 8 : 4;
            c->col = (uint32_t*)dev_alloc(ctx, ((size_t)nnz + 16) * 4);
            c->val = dev_alloc(ctx, ((size_t)nnz + 16) * es);
            SALG_CUDA(cudaMemsetAsync(c->col + nnz, 0, 16 * 4, st));
            SALG_CUDA(cudaMemsetAsync((char*)c->val + (size_t)nnz * es, 0, 16 * es, st));
            if (nrows) {
                if (dtype == SALG_F64) { synth_fill_kernel<double><<<grid, 256, 0, st>>>(a, c->row_ptr, c->col, (double*)c->val); ctx->n_launch++; }
                else { synth_fill_kernel<float><<<grid, 256, 0, st>>>(a, c->row_ptr, c->col, (float*)c->val); ctx->n_launch++; }
                SALG_CUDA(cudaGetLastError());
            }
            SALG_CUDA(cudaStreamSynchronize(st));
        } catch (...) {
            dev_free(ctx, ptr);
            if (c) csr_destroy(c);
            throw;
        }
        *out = c;
    });
}
