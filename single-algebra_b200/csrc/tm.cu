// tm.cu — sparse x panel products with the SPARSE operand expanded straight into tensor memory (tcgen05.mma, A from TMEM).
//
// Why a second generation: the tile-densified kernels of tc.cu materialise every 128 x 64 tile as a zeroed dense fp16 tile in
// SHARED memory (clear + scatter + tensor-core read = 48 KB of shared-memory traffic for ~570 useful entries) and measured
// shared-memory / hand-off bound at 0.34 of the HBM roofline.  Here the dense tile never exists in shared
// memory: lane (= TMEM data path) i of a 128 x 128 tile is owned by ONE thread, which expands its compressed row —
// a 32-bit mask of non-empty QUADS (4 consecutive k) + the packed quads — into registers with predicated 8-byte
// shared-memory loads and writes them to TMEM with tcgen05.st.  The tensor core reads the sparse operand from TMEM
// (M = 128 lanes, K = 16 per instruction) and only the DENSE panel slice from shared memory.
//
//   A X   : lanes = operator rows,    k = operator columns; B = pre-split panel slice  [N = 64 | 64 (hi | lo)][K = 128]
//           accumulator D[row][panel column] (hi and lo products summed by the tensor core: 64 TMEM columns per row block)
//   A^T Y : lanes = operator columns, k = operator rows;    B = pre-split Y row block  [N = 128 (2 col + term)][K = 128]
//           (the layout tc_gram_prep_kernel writes), D[operator column][2 col + term], flushed with atomics per work item
//
// Tile format (one per orientation, built once per operator by tm_build_kernel): tile t = row block * n_cb + column block;
//   info[t]  u64   : first quad of the tile's RECORD in q_hi / q_lo (bits 0-39, even) | quad count (bits 40-63)
//   record         : 16 B {first quad of the payload (u64), quad count (u32)}, 768 B header = 128 x u32 quad masks +
//                    128 x u16 quad offsets (exclusive prefix of the lanes' quad counts), then the quads: 4 x fp16 each,
//                    lane-major then k-ascending.  One bulk copy brings a whole record into a ring slot.
//   q_hi / q_lo    : the record streams of the two fp16 terms (q_lo only when the values are not exact in fp16)
// HBM bytes per tile at 7 % density: 784 + ~1030 quads x 8 B = 9 KB for 1146 entries (the CSR's 8 B per entry).
//
// Roles inside a CTA (warp-specialised; hand-offs through mbarriers): 16 expander warps = 4 sets x 4 lane quarters (set s
// builds passes p = s mod 4), 4 epilogue warps, 1 MMA issuer, 1 panel-slice loader, 1 tile loader (bulk copies of header
// and quads into a ring).
#include <algorithm>
#include <cuda_fp16.h>

#include "common.cuh"
#define TC_SPIN_LIMIT_ (1u << 15)      // a protocol bug traps within seconds
#include "tcgen05.cuh"

namespace salg {

constexpr int TM_LANES = 128;            // tile lanes (M of the MMA)
constexpr int TM_DEPTH = 128;            // tile depth (k), 8 K-steps of 16
constexpr int TM_HDR = 768;              // header bytes per tile
constexpr int TM_REC = 16 + TM_HDR;      // bytes of a record before its quads
constexpr int TM_REC_Q = TM_REC / 8;     // ... in quads (98)
constexpr int TM_BUILD_WARPS = 16;
constexpr int TM_SETS = TM_BUILD_WARPS / 4;
constexpr int TM_W_EPI = TM_BUILD_WARPS, TM_W_MMA = TM_BUILD_WARPS + 4, TM_W_PLOAD = TM_BUILD_WARPS + 6,
              TM_W_ELOAD = TM_BUILD_WARPS + 7;
constexpr int TM_THREADS = (TM_BUILD_WARPS + 9) * 32;   // + 4 epilogue, 2 MMA issuers, 1 panel loader, 2 tile loaders
#ifndef TM_NA_
#define TM_NA_ 4
#endif
#ifndef TM_NS_
#define TM_NS_ 8
#endif
#ifndef TM_NB_
#define TM_NB_ 3
#endif
constexpr int TM_NA = TM_NA_;            // sparse-operand buffers in TMEM (64 columns = 128 k each)
constexpr int TM_NS = TM_NS_;            // tile ring slots
#ifndef TM_SLOT_QUADS_
#define TM_SLOT_QUADS_ 1152
#endif
constexpr int TM_SLOT_QUADS = TM_SLOT_QUADS_;      // quads a ring slot holds; denser tiles are expanded from global memory
constexpr int TM_SLOT_BYTES = TM_REC + TM_SLOT_QUADS * 8;
constexpr int TM_NB = TM_NB_;            // panel stages
constexpr int TM_STAGE_BYTES = 32768;    // one panel block: 128 (N) x 128 (K) fp16
constexpr int TM_A_COL0 = 256;           // TMEM columns [0, 256): accumulators, [256, 512): sparse-operand buffers
constexpr int TM_EPI_PITCH = 144;         // A X epilogue staging: 32 rows x 128 B per warp, rows 144 B apart (conflict-free both ways)
constexpr int TM_EPI_BYTES = 4 * 32 * TM_EPI_PITCH;
constexpr int TM_SMEM = TM_NB * TM_STAGE_BYTES + TM_NS * TM_SLOT_BYTES + TM_EPI_BYTES + 128;
constexpr int TM_PF = 8;                 // L2 prefetch distance of the tile records, in passes of one loader
constexpr int TM_GC = 2;                 // A^T Y: operator column blocks (accumulators of 128 TMEM columns) per work item
constexpr int TM_MAX_CB = 1 << 20;       // (the builder works chunk by chunk: no width limit of its own)
static_assert(TM_NA * 64 <= 256, "sparse-operand buffers exceed their TMEM half");
static_assert((TM_NA & (TM_NA - 1)) == 0, "A-buffer ring is a power of two (index = pass & (n - 1))");
// Parity waits are only safe on a barrier with ONE waiting role that sees every phase in order (a waiter a phase early would
// alias with the phase before).  Pass p is expanded by set p % 4 into A buffer p % 4 from ring slot p % TM_NS, loaded by
// loader p % 2: with TM_NS a multiple of 4 every slot has one loader and one expander set, every A buffer one set; the issuer
// of A buffer x is x % 2 resp. (x / terms) % 2 within a group, and the issuers re-synchronise between groups (acc_free).
static_assert(TM_NS % 4 == 0 && TM_NA == TM_SETS, "slot -> (loader, expander set) and A buffer -> set must be fixed");
static_assert(TM_SMEM <= 227 * 1024, "shared memory");

struct TmFormat {
    uint64_t* info = nullptr;
    uint2* q_hi = nullptr;
    uint2* q_lo = nullptr;
};
struct TmTiles {
    TmFormat R, T;                       // lanes = rows (A X), lanes = columns (A^T Y)
    int n_rb = 0, n_cb = 0;              // row blocks (padded to a multiple of 4), column blocks of 128
    int a_terms = 1;
    float a_scale = 1.f;
};

static void tm_free_format(salg_ctx* owner, TmFormat& f) {
    dev_free(owner, f.info);
    dev_free(owner, f.q_hi);
    dev_free(owner, f.q_lo);
    f = TmFormat();
}
void tm_free(salg_ctx* owner, void* p) {
    TmTiles* t = (TmTiles*)p;
    if (!t) return;
    tm_free_format(owner, t->R);
    tm_free_format(owner, t->T);
    delete t;
}
bool tm_supported(const salg_csr* c) { return ceil_div(c->ncols, TM_DEPTH) <= TM_MAX_CB; }
int tm_n_rb(const void* tiles) { return ((const TmTiles*)tiles)->n_rb; }
int tm_terms(const void* tiles) { return ((const TmTiles*)tiles)->a_terms; }
float tm_scale(const void* tiles) { return ((const TmTiles*)tiles)->a_scale; }

// ---- format builder --------------------------------------------------------------------------------------------------
__global__ void tm_info_kernel(const int64_t* __restrict__ in_ptr, int in_shift, const int64_t* __restrict__ ptr,
                               const float* __restrict__ val, int64_t nrows, unsigned* __restrict__ info) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    float amax = 0.f;
    for (int64_t r = w; r < nrows; r += nw) {
        const int64_t s = in_ptr[r] >> in_shift;
        const int64_t len = ptr[r + 1] - ptr[r];
        for (int64_t p = lane; p < len; p += 32) {
            const float v = val[s + p];
            bad |= (__half2float(__float2half_rn(v)) != v);
            amax = fmaxf(amax, fabsf(v));
        }
    }
    if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) atomicOr(&info[0], 1u);
#pragma unroll
    for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
    if (lane == 0) atomicMax(&info[1], __float_as_uint(amax));
}

// first quad of row block rb's region in the quad arrays: the block's first entry + (record prefix + two quads of slack) per
// tile before it (a record's start is rounded up to an even quad: 16 B alignment of the bulk copies), rounded up to even.  A block never has
// more quads than entries, so the regions cannot overlap.
__host__ __device__ inline int64_t tm_region(int64_t first_entry, int64_t rb, int n_cb) {
    return (first_entry + rb * (int64_t)n_cb * (TM_REC_Q + 2) + 1) & ~(int64_t)1;
}

// One CTA per 128-row block and orientation; the block's tiles are built CHUNK by chunk (`ch` consecutive tiles):
//   sweep 1  every warp walks its rows from a per-row cursor to the chunk's last column and ORs the quad masks (shared memory)
//   offsets  the lanes' quad counts are prefix-summed per tile, the tiles' record starts continue the block's running total
//   sweep 2  the same entries again (L2 resident): every value lands as one fp16 in an IMAGE of the chunk's records in shared
//            memory, which is then written out with coalesced 16-byte stores
// Scattering the 2-byte stores straight to global memory measured ~55 partial-sector writes per clock on the whole chip —
// 1.2-1.5 ms per orientation at config 3, 26 ms for a 500k x 33k shard — and made the masks of ALL tiles of a row block live
// in shared memory (a 38k-column limit).  A chunk whose records do not fit the image is scattered directly (dense data).
// TRANS: lanes = columns, k = rows (operand of A^T Y).  A warp walks a row in batches of 4 x 32 entries, loads first.
constexpr int TM_BUILD_THREADS = 512;
template <bool TRANS>
__global__ void __launch_bounds__(TM_BUILD_THREADS)
tm_build_kernel(const int64_t* __restrict__ in_ptr, int in_shift, const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                const float* __restrict__ val, int64_t nrows, int n_rb_real, int n_cb, int terms, float a_scale,
                uint64_t* __restrict__ info, uint2* __restrict__ q_hi, uint2* __restrict__ q_lo, int ch, unsigned img_quads) {
    constexpr int THREADS = TM_BUILD_THREADS;
    extern __shared__ __align__(16) uint32_t tm_bsm[];
    uint32_t* s_mask = tm_bsm;                                      // [ch][128]
    uint32_t* s_start = s_mask + (size_t)ch * TM_LANES;             // [ch] tile totals, then record starts (from the region's start)
    unsigned short* s_off = reinterpret_cast<unsigned short*>(s_start + ch);     // [ch][128]
    uint2* s_img = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(tm_bsm) + (((size_t)ch * (TM_LANES * 4 + 4 + TM_LANES * 2) + 15) & ~(size_t)15));
    __shared__ uint32_t s_wsum[THREADS / 32];
    __shared__ uint32_t s_total;
    __shared__ int s_cur0[TM_LANES], s_cur1[TM_LANES];              // per row: first entry of the chunk, first entry after it
    __shared__ long long s_row[TM_LANES];                           // per row: first stored entry
    __shared__ int s_len[TM_LANES];                                 // ... and their number
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    constexpr int BU = 4;
    for (int rb = blockIdx.x; rb < n_rb_real; rb += gridDim.x) {
        const int64_t r0 = (int64_t)rb * TM_LANES;
        const int64_t r1 = r0 + TM_LANES < nrows ? r0 + TM_LANES : nrows;
        const int64_t qbase = tm_region(ptr[r0], rb, n_cb);
        unsigned run = 0;                                           // quads of the block's region used so far (even)
        __syncthreads();
        if (tid < TM_LANES) {
            s_cur0[tid] = 0;
            const int64_t r = r0 + tid;
            s_row[tid] = r < r1 ? (long long)(in_ptr[r] >> in_shift) : 0;
            s_len[tid] = r < r1 ? (int)(ptr[r + 1] - ptr[r]) : 0;
        }
        for (int jc = 0; jc < n_cb; jc += ch) {
            const int nt = jc + ch < n_cb ? ch : n_cb - jc;
            const unsigned c_end = (unsigned)(jc + nt) * TM_DEPTH;
            for (int i = tid; i < nt * TM_LANES; i += THREADS) s_mask[i] = 0;
            __syncthreads();
            // sweep 1: quad masks of the chunk; finds every row's first entry beyond the chunk
            for (int64_t r = r0 + warp; r < r1; r += NW) {
                const unsigned lr = (unsigned)(r - r0);
                const int len = s_len[lr];
                const uint32_t* __restrict__ cr = col + s_row[lr];
                int p0 = s_cur0[lr];
                bool more = true;
                while (more && p0 < len) {
                    unsigned c[BU];
                    const int nbt = (len - p0 + 31) >> 5;                    // sub-batches with entries (warp-uniform)
#pragma unroll
                    for (int u = 0; u < BU; u++) {
                        const int q = p0 + lane + 32 * u;
                        c[u] = (u < nbt && q < len) ? cr[q] : 0xFFFFFFFFu;
                    }
                    int done = 0;
#pragma unroll
                    for (int u = 0; u < BU; u++) {
                        if (u >= nbt) break;
                        const bool in_chunk = c[u] < c_end;                  // (absent entries compare false)
                        done += __popc(__ballot_sync(0xFFFFFFFFu, in_chunk));
                        if (!in_chunk) continue;
                        const unsigned j = (c[u] >> 7) - (unsigned)jc;
                        const unsigned li = TRANS ? (c[u] & 127u) : lr;
                        const unsigned q = TRANS ? (lr >> 2) : ((c[u] & 127u) >> 2);
                        atomicOr(&s_mask[j * TM_LANES + li], 1u << q);
                    }
                    p0 += done;
                    more = done == 32 * BU;                                  // a shorter batch met the chunk's (or the row's) end
                }
                if (lane == 0) s_cur1[lr] = p0;
            }
            __syncthreads();
            // lanes' quad offsets inside every tile (four tiles per trip) and the tiles' quad counts
            for (int j0 = 0; j0 < nt; j0 += THREADS / TM_LANES) {
                const int j = j0 + (tid >> 7), li = tid & 127;
                const unsigned cnt = j < nt ? (unsigned)__popc(s_mask[j * TM_LANES + li]) : 0u;
                unsigned incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (lane == 31) s_wsum[warp] = incl;
                __syncthreads();
                unsigned base = 0;
                for (int w = warp & ~3; w < warp; w++) base += s_wsum[w];
                if (j < nt) {
                    s_off[j * TM_LANES + li] = (unsigned short)(base + incl - cnt);
                    if (li == TM_LANES - 1) s_start[j] = base + incl;
                }
                __syncthreads();
            }
            // record starts (even), continuing the block's region
            if (warp == 0) {
                unsigned r_run = run;
                for (int j0 = 0; j0 < nt; j0 += 32) {
                    const int j = j0 + lane;
                    const unsigned tot = j < nt ? s_start[j] : 0u;
                    const unsigned padded = j < nt ? TM_REC_Q + ((tot + 1u) & ~1u) : 0u;
                    unsigned incl = padded;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const unsigned start = r_run + incl - padded;
                    if (j < nt) {
                        s_start[j] = start;
                        info[(int64_t)rb * n_cb + jc + j] = (uint64_t)(qbase + start) | ((uint64_t)tot << 40);
                    }
                    r_run += __shfl_sync(0xFFFFFFFFu, incl, 31);
                }
                if (lane == 0) s_total = r_run;
            }
            __syncthreads();
            const unsigned c_quads = s_total - run;                  // the chunk's records, quads (even)
            const bool staged = c_quads <= img_quads;                // CTA-uniform
            for (int t = 0; t < terms; t++) {
                uint2* qs = (t ? q_lo : q_hi) + qbase + run;         // the chunk's records in this term's stream
                uint2* img = staged ? s_img : qs;
                for (unsigned i = tid; i < c_quads / 2; i += THREADS) reinterpret_cast<uint4*>(img)[i] = make_uint4(0u, 0u, 0u, 0u);
                __syncthreads();
                // record prefixes: {first payload quad, count} + masks + offsets
                for (int i = tid; i < nt * TM_LANES; i += THREADS) {
                    const int j = i >> 7, li = i & 127;
                    uint8_t* rec = reinterpret_cast<uint8_t*>(img + (s_start[j] - run));
                    reinterpret_cast<uint32_t*>(rec + 16)[li] = s_mask[i];
                    reinterpret_cast<unsigned short*>(rec + 16 + 512)[li] = s_off[i];
                    if (li == TM_LANES - 1) {
                        const unsigned tot = (unsigned)s_off[i] + (unsigned)__popc(s_mask[i]);
                        *reinterpret_cast<unsigned long long*>(rec) = (unsigned long long)(qbase + s_start[j] + TM_REC_Q);
                        *reinterpret_cast<uint2*>(rec + 8) = make_uint2(tot, 0u);
                    }
                }
                // sweep 2: the chunk's values
                unsigned short* h_img = reinterpret_cast<unsigned short*>(img);
                for (int64_t r = r0 + warp; r < r1; r += NW) {
                    const unsigned lr = (unsigned)(r - r0);
                    const long long s = s_row[lr];
                    const uint32_t* __restrict__ cr = col + s;
                    const float* __restrict__ vr = val + s;
                    const int e1 = s_cur1[lr];
                    for (int p0 = s_cur0[lr]; p0 < e1; p0 += 32 * BU) {
                        unsigned c[BU];
                        float x[BU];
                        const int nbt = (e1 - p0 + 31) >> 5;                 // sub-batches with entries (warp-uniform)
#pragma unroll
                        for (int u = 0; u < BU; u++) {
                            const int q = p0 + lane + 32 * u;
                            c[u] = (u < nbt && q < e1) ? cr[q] : 0xFFFFFFFFu;
                            x[u] = (u < nbt && q < e1) ? vr[q] * a_scale : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < BU; u++) {
                            if (u >= nbt) break;
                            if (c[u] == 0xFFFFFFFFu) continue;
                            const unsigned j = (c[u] >> 7) - (unsigned)jc;
                            const unsigned li = TRANS ? (c[u] & 127u) : lr;
                            const unsigned k = TRANS ? lr : (c[u] & 127u);
                            const unsigned q = k >> 2, e = k & 3u;
                            const unsigned m = s_mask[j * TM_LANES + li];
                            const size_t pos = (size_t)(s_start[j] - run) + TM_REC_Q + s_off[j * TM_LANES + li] + __popc(m & ((1u << q) - 1u));
                            const __half hh = __float2half_rn(x[u]);
                            h_img[pos * 4 + e] = __half_as_ushort(t ? __float2half_rn(x[u] - __half2float(hh)) : hh);
                        }
                    }
                }
                if (staged) {
                    __syncthreads();
                    for (unsigned i = tid; i < c_quads / 2; i += THREADS) reinterpret_cast<uint4*>(qs)[i] = reinterpret_cast<const uint4*>(s_img)[i];
                }
                __syncthreads();
            }
            if (tid < TM_LANES) s_cur0[tid] = s_cur1[tid];
            run = s_total;
            __syncthreads();
        }
    }
}

// ---- second builder: BOTH orientations of a row block's tiles in one pass over its entries, through dense tile images ----------
// The first builder turns every entry into a shared-memory atomic (quad masks) and, in a second sweep, a 2-byte store at a
// position looked up from three shared-memory tables, once per orientation: ~250 thread instructions per entry and orientation
// (ncu: 1.08e9 warp instructions per orientation at config 3, 1.4-1.5 ms each).  Here a chunk of TB2_CH column blocks is built
// for both orientations at once:
//   scatter  warps walk their rows' entries of the chunk (coalesced loads, all of a warp's rows in flight) and drop every value
//            into TWO zeroed dense fp16 images of its 128 x 128 tile, [row][k = column] and [column][k = row]
//   scan     one thread per (tile, orientation, lane) reads the lane's 32 quads (8 B each) of its image into registers: the non-zero
//            ones ARE the lane's quads in k order, their pattern the quad mask; a prefix sum over the 128 lanes gives the offsets; the
//            record is put together in shared memory and leaves with coalesced 16-byte stores.  The lane re-zeroes what it took.
// An entry costs two 2-byte stores, a quad one 8-byte load and store; both orientations share the loads of the entries.
// Values whose fp16 image is zero (stored zeros, |x| below 2^-25 of the scale) drop out of the format: they contribute nothing.
constexpr int TB2_THREADS = 512;
constexpr int TB2_CH = 2;                                       // column blocks per chunk: TB2_CH x 2 orientations x 128 lanes = threads
constexpr int TB2_PITCH = 272;                                  // bytes per image row: 128 fp16 + 16 (conflict-free 16-byte reads down a column of lanes)
constexpr int TB2_IMG = TM_LANES * TB2_PITCH;
constexpr int TB2_STAGE_QUADS = 1536;                           // payload quads of a staged record; denser tiles store straight to global
constexpr int TB2_STAGE_BYTES = TM_REC + TB2_STAGE_QUADS * 8;
constexpr int TB2_SMEM = TB2_CH * 2 * (TB2_IMG + TB2_STAGE_BYTES);
constexpr int TB2_RPW = TM_LANES / (TB2_THREADS / 32);          // rows per warp
static_assert(TB2_THREADS == TB2_CH * 2 * TM_LANES, "one scan thread per (tile of the chunk, orientation, lane)");
static_assert(TB2_STAGE_BYTES % 16 == 0 && TB2_IMG % 16 == 0, "alignment of the staging areas");
static_assert(TB2_SMEM <= 220 * 1024, "shared memory");

__global__ void __launch_bounds__(TB2_THREADS, 1)
tm_build2_kernel(const int64_t* __restrict__ in_ptr, int in_shift, const int64_t* __restrict__ ptr, const uint32_t* __restrict__ col,
                 const float* __restrict__ val, int64_t nrows, int n_rb_real, int n_cb, int terms, float a_scale,
                 uint64_t* __restrict__ info_r, uint2* __restrict__ r_hi, uint2* __restrict__ r_lo,
                 uint64_t* __restrict__ info_t, uint2* __restrict__ t_hi, uint2* __restrict__ t_lo) {
    extern __shared__ __align__(16) uint8_t tb2_sm[];
    uint8_t* images = tb2_sm;                                   // [tile j][orientation o] -> images + (2 j + o) * TB2_IMG
    uint8_t* stages = tb2_sm + TB2_CH * 2 * TB2_IMG;            // same indexing, TB2_STAGE_BYTES each
    __shared__ long long s_row[TM_LANES];
    __shared__ int s_len[TM_LANES], s_cur[TM_LANES];
    __shared__ unsigned s_wsum[TB2_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid >> 7, li = tid & 127;                   // scan role: group = 2 j + o (4 warps each), lane li
    const int my_j = grp >> 1, my_o = grp & 1;
    for (int i = tid; i < TB2_CH * 2 * TB2_IMG / 16; i += TB2_THREADS) reinterpret_cast<uint4*>(images)[i] = make_uint4(0u, 0u, 0u, 0u);
    uint8_t* my_img = images + grp * TB2_IMG;
    uint8_t* my_stage = stages + grp * TB2_STAGE_BYTES;
    uint64_t* my_info = my_o ? info_t : info_r;
    for (int rb = blockIdx.x; rb < n_rb_real; rb += gridDim.x) {
        const int64_t r0 = (int64_t)rb * TM_LANES;
        const int64_t qbase = tm_region(ptr[r0], rb, n_cb);
        unsigned run = 0;                                       // quads of the block's region used so far by MY orientation (even)
        __syncthreads();
        if (tid < TM_LANES) {
            const int64_t r = r0 + tid;
            s_row[tid] = r < nrows ? (long long)(in_ptr[r] >> in_shift) : 0;
            s_len[tid] = r < nrows ? (int)(ptr[r + 1] - ptr[r]) : 0;
            s_cur[tid] = 0;
        }
        __syncthreads();
        for (int jc = 0; jc < n_cb; jc += TB2_CH) {
            const int nt = jc + TB2_CH < n_cb ? TB2_CH : n_cb - jc;
            const unsigned c_end = (unsigned)(jc + nt) * TM_DEPTH;
            unsigned mask0 = 0, off0 = 0, tot0 = 0, start0 = 0, adv = 0;
            for (int t = 0; t < terms; t++) {
                // ---- scatter: warp w takes rows w, w + 16, ...; first batches of all its rows in flight together
                {
                    unsigned c[TB2_RPW];
                    float x[TB2_RPW];
#pragma unroll
                    for (int i = 0; i < TB2_RPW; i++) {
                        const int lr = warp + (TB2_THREADS / 32) * i;
                        const int q = s_cur[lr] + lane;
                        const bool ok = q < s_len[lr];
                        c[i] = ok ? col[s_row[lr] + q] : 0xFFFFFFFFu;
                        x[i] = ok ? val[s_row[lr] + q] : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < TB2_RPW; i++) {
                        const int lr = warp + (TB2_THREADS / 32) * i;
                        unsigned cc = c[i];
                        float xx = x[i];
                        int taken = 0, q0 = s_cur[lr];
                        while (true) {
                            const bool in_chunk = cc < c_end;                    // (absent entries compare false)
                            const unsigned bal = __ballot_sync(0xFFFFFFFFu, in_chunk);
                            if (in_chunk) {
                                const float xs = xx * a_scale;
                                const __half hh = __float2half_rn(xs);
                                const unsigned short h = __half_as_ushort(t ? __float2half_rn(xs - __half2float(hh)) : hh);
                                const unsigned j = (cc >> 7) - (unsigned)jc, k = cc & 127u;
                                uint8_t* im = images + (2u * j) * TB2_IMG;
                                *reinterpret_cast<unsigned short*>(im + (unsigned)lr * TB2_PITCH + k * 2u) = h;
                                *reinterpret_cast<unsigned short*>(im + TB2_IMG + k * TB2_PITCH + (unsigned)lr * 2u) = h;
                            }
                            taken += __popc(bal);
                            if (bal != 0xFFFFFFFFu) break;                       // a shorter batch met the chunk's (or the row's) end
                            q0 += 32;
                            const int q = q0 + lane;
                            const bool ok = q < s_len[lr];
                            cc = ok ? col[s_row[lr] + q] : 0xFFFFFFFFu;
                            xx = ok ? val[s_row[lr] + q] : 0.f;
                        }
                        if (t == terms - 1 && lane == 0) s_cur[lr] += taken;   // (an L2 prefetch of the next chunk here measured +5 %)
                    }
                }
                __syncthreads();
                // ---- scan: thread = lane li of (tile my_j, orientation my_o)
                const bool have = my_j < nt;
                uint2 q[32];
                unsigned nz = 0;
                if (have) {
                    // 16-byte loads, and the lane's whole image row zeroed again with 16-byte stores (instruction count, not
                    // shared-memory bandwidth, bounds this phase)
                    uint4* rowp = reinterpret_cast<uint4*>(my_img + li * TB2_PITCH);
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const uint4 v = rowp[i];
                        q[2 * i] = make_uint2(v.x, v.y);
                        q[2 * i + 1] = make_uint2(v.z, v.w);
                    }
#pragma unroll
                    for (int i = 0; i < 16; i++) rowp[i] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int i = 0; i < 32; i++) nz |= ((q[i].x | q[i].y) != 0u ? 1u : 0u) << i;
                }
                if (t == 0) {
                    mask0 = nz;
                    const unsigned cnt = (unsigned)__popc(nz);
                    unsigned incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (lane == 31) s_wsum[warp] = incl;
                    __syncthreads();
                    unsigned base = 0, tot = 0, tot_first = 0;               // tot_first: quads of tile 0 of the chunk in MY orientation
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        const unsigned v = s_wsum[grp * 4 + w];
                        if (grp * 4 + w < warp) base += v;
                        tot += v;
                        tot_first += s_wsum[my_o * 4 + w];
                    }
                    off0 = base + incl - cnt;
                    tot0 = tot;
                    const unsigned rec_first = TM_REC_Q + ((tot_first + 1u) & ~1u);
                    start0 = run + (my_j ? rec_first : 0u);
                    // this chunk's advance of MY orientation's region: tile 0's record, and tile 1's when the chunk has one
                    unsigned tot_second = 0;
#pragma unroll
                    for (int w = 0; w < 4; w++) tot_second += s_wsum[(2 + my_o) * 4 + w];
                    adv = rec_first + (nt > 1 ? TM_REC_Q + ((tot_second + 1u) & ~1u) : 0u);
                    if (have && li == 0) my_info[(int64_t)rb * n_cb + jc + my_j] = (uint64_t)(qbase + start0) | ((uint64_t)tot << 40);
                }
                const bool staged = tot0 <= (unsigned)TB2_STAGE_QUADS;             // uniform over the group's 128 threads
                uint2* g_rec = (my_o ? (t ? t_lo : t_hi) : (t ? r_lo : r_hi)) + qbase + start0;
                if (have) {
                    uint8_t* rec = staged ? my_stage : reinterpret_cast<uint8_t*>(g_rec);
                    reinterpret_cast<uint32_t*>(rec + 16)[li] = mask0;
                    reinterpret_cast<unsigned short*>(rec + 16 + 512)[li] = (unsigned short)off0;
                    if (li == TM_LANES - 1) {
                        *reinterpret_cast<unsigned long long*>(rec) = (unsigned long long)(qbase + start0 + TM_REC_Q);
                        *reinterpret_cast<uint2*>(rec + 8) = make_uint2(tot0, 0u);
                        if (tot0 & 1u) reinterpret_cast<uint2*>(rec + TM_REC)[tot0] = make_uint2(0u, 0u);   // pad to an even quad count
                    }
                    uint2* pay = reinterpret_cast<uint2*>(rec + TM_REC) + off0;
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        if ((mask0 >> i) & 1u) *pay = q[i];
                        pay += (mask0 >> i) & 1u;
                    }
                }
                __syncthreads();
                // ---- staged records leave with coalesced 16-byte stores (each group copies its own)
                if (have && staged) {
                    const unsigned n16 = (TM_REC_Q + ((tot0 + 1u) & ~1u)) / 2;
                    uint4* dst = reinterpret_cast<uint4*>(g_rec);
                    const uint4* src = reinterpret_cast<const uint4*>(my_stage);
                    for (unsigned i = li; i < n16; i += TM_LANES) dst[i] = src[i];
                }
                // (the next scatter only touches the images; the next record is staged after another barrier)
            }
            run += adv;
        }
    }
}

__global__ void tm_fill_u64_kernel(uint64_t* p, int64_t n, uint64_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

void* tm_build(salg_ctx* ctx, const salg_csr* c, const int64_t* in_ptr, const uint32_t* in_col, const float* in_val, int in_shift) {
    cudaStream_t st = ctx->stream;
    if (!in_ptr) {
        in_ptr = c->row_ptr;
        in_col = c->col;
        in_val = (const float*)c->val;
    }
    TmTiles* t = new TmTiles();
    try {
        const int n_rb_real = (int)ceil_div(c->nrows, TM_LANES);
        t->n_rb = (int)(ceil_div(n_rb_real, 4) * 4);
        t->n_cb = (int)ceil_div(c->ncols, TM_DEPTH);
        SALG_REQUIRE(t->n_cb <= TM_MAX_CB, SALG_ERR_UNSUPPORTED, "operator too wide for the TMEM-operand tile format");
        const int64_t n_tiles = (int64_t)t->n_rb * t->n_cb;
        ProfScope ps(ctx, PROF_TRANSPOSE, (double)c->nnz * (4 + 4 + 8 + 8));
        DevBuf<unsigned> info(2, st);
        SALG_CUDA(cudaMemsetAsync(info.get(), 0, 8, st));
        if (c->nnz > 0) {
            int64_t want = ceil_div(c->nrows * 32, 256), cap = (int64_t)ctx->sm_count * 16;
            tm_info_kernel<<<(unsigned)std::max<int64_t>(1, std::min(want, cap)), 256, 0, st>>>(in_ptr, in_shift, c->row_ptr, in_val,
                                                                                                  c->nrows, info.get());
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        unsigned h_info[2] = {0, 0};
        SALG_CUDA(cudaMemcpyAsync(h_info, info.get(), 8, cudaMemcpyDeviceToHost, st));
        SALG_CUDA(cudaStreamSynchronize(st));
        float amax;
        memcpy(&amax, &h_info[1], 4);
        t->a_terms = h_info[0] ? 2 : 1;
        t->a_scale = h_info[0] ? tc_pow2_scale(amax) : 1.f;
        // capacity: every entry its own quad + record prefix and alignment slack per tile + one all-zero record at the end
        // (the tiles of the padding row blocks point at it)
        const size_t q_cap = (size_t)c->nnz + (size_t)n_tiles * (TM_REC_Q + 2) + 2 * TM_REC_Q + 64;
        const uint64_t zero_rec = (uint64_t)((q_cap - TM_REC_Q - 32) & ~(size_t)1);
        // chunk = as many tiles as fit a ~64 KB image at the operator's average density (three CTAs per SM)
        const double avg_tile_q = TM_REC_Q + 2.0 + (double)c->nnz / std::max<double>(1.0, (double)n_rb_real * t->n_cb);
        const unsigned img_quads = 8192;
        int ch = (int)std::max(1.0, std::min<double>(32.0, std::floor(img_quads * 0.9 / avg_tile_q)));
        ch = std::min(ch, t->n_cb);
        const size_t smem_mask = ((size_t)ch * (TM_LANES * 4 + 4 + TM_LANES * 2) + 15) & ~(size_t)15;
        const size_t smem = smem_mask + (size_t)img_quads * 8;
        static const bool first_builder = getenv("SALG_TM_BUILD") && atoi(getenv("SALG_TM_BUILD")) == 1;
        for (int o = 0; o < 2; o++) {
            TmFormat& f = o ? t->T : t->R;
            f.info = (uint64_t*)dev_alloc(ctx, (size_t)(n_tiles + 1) * 8);
            f.q_hi = (uint2*)dev_alloc(ctx, q_cap * 8);
            if (t->a_terms > 1) f.q_lo = (uint2*)dev_alloc(ctx, q_cap * 8);
            SALG_CUDA(cudaMemsetAsync(f.q_hi + zero_rec, 0, (q_cap - zero_rec) * 8, st));
            if (f.q_lo) SALG_CUDA(cudaMemsetAsync(f.q_lo + zero_rec, 0, (q_cap - zero_rec) * 8, st));
            if (t->n_rb > n_rb_real) {
                const int64_t n_pad = (int64_t)(t->n_rb - n_rb_real) * t->n_cb;
                tm_fill_u64_kernel<<<(unsigned)ceil_div(n_pad, 256), 256, 0, st>>>(f.info + (size_t)n_rb_real * t->n_cb, n_pad, zero_rec);
                ctx->n_launch++;
            }
            if (n_rb_real == 0 || !first_builder) continue;
            const int grid = std::min(n_rb_real, ctx->sm_count * 3);
            if (o == 0) {
                set_max_dyn_smem(tm_build_kernel<false>, (int)smem);
                tm_build_kernel<false><<<grid, TM_BUILD_THREADS, smem, st>>>(in_ptr, in_shift, c->row_ptr, in_col, in_val, c->nrows, n_rb_real,
                                                                            t->n_cb, t->a_terms, t->a_scale, f.info, f.q_hi, f.q_lo, ch, img_quads);
            } else {
                set_max_dyn_smem(tm_build_kernel<true>, (int)smem);
                tm_build_kernel<true><<<grid, TM_BUILD_THREADS, smem, st>>>(in_ptr, in_shift, c->row_ptr, in_col, in_val, c->nrows, n_rb_real,
                                                                           t->n_cb, t->a_terms, t->a_scale, f.info, f.q_hi, f.q_lo, ch, img_quads);
            }
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
        if (n_rb_real > 0 && !first_builder) {
            // both orientations in one pass (dense tile images); SALG_TM_BUILD=1 selects the first builder
            set_max_dyn_smem(tm_build2_kernel, TB2_SMEM);
            const int grid = std::min(n_rb_real, ctx->sm_count);
            tm_build2_kernel<<<grid, TB2_THREADS, TB2_SMEM, st>>>(in_ptr, in_shift, c->row_ptr, in_col, in_val, c->nrows, n_rb_real, t->n_cb,
                                                                  t->a_terms, t->a_scale, t->R.info, t->R.q_hi, t->R.q_lo, t->T.info,
                                                                  t->T.q_hi, t->T.q_lo);
            ctx->n_launch++;
            SALG_CUDA(cudaGetLastError());
        }
    } catch (...) {
        cudaStreamSynchronize(st);
        tm_free(ctx, t);
        throw;
    }
    return t;
}

// ---- panel pre-split for A X: block b = K rows [128 b, 128 b + 128), [hi | lo] halves of 16 KB, each the canonical K-major
//      (N = 64 panel columns) x (K = 128) operand with 16 B K-chunks 1024 B apart --------------------------------------------
__global__ void __launch_bounds__(256)
tm_prep_x_kernel(const float* __restrict__ P, int64_t n, const float* __restrict__ scales, uint8_t* __restrict__ out) {
    __shared__ float tile[32][LP + 1];
    const float s = scales[0];
    const int64_t k0 = (int64_t)blockIdx.x * 32;
    for (int i = threadIdx.x; i < 32 * LP; i += 256) {
        const int r = i >> 6, c = i & 63;
        const int64_t k = k0 + r;
        tile[r][c] = (k < n) ? P[k * LP + c] * s : 0.f;
    }
    __syncthreads();
    const int o = threadIdx.x >> 6, c = threadIdx.x & 63;          // K-octet of the 32 rows, panel column
    const int64_t kc0 = k0 + o * 8;
    const int64_t b = kc0 >> 7;
    const uint32_t chunk = (uint32_t)(kc0 & 127) >> 3;
    unsigned short hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float x = tile[o * 8 + j][c];
        const __half h = __float2half_rn(x);
        hi[j] = __half_as_ushort(h);
        lo[j] = __half_as_ushort(__float2half_rn(x - __half2float(h)));
    }
    uint8_t* dst = out + (size_t)b * TM_STAGE_BYTES + chunk * 1024u + (uint32_t)(c >> 3) * 128u + (uint32_t)(c & 7) * 16u;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0] | ((uint32_t)hi[1] << 16), hi[2] | ((uint32_t)hi[3] << 16),
                                                hi[4] | ((uint32_t)hi[5] << 16), hi[6] | ((uint32_t)hi[7] << 16));
    *reinterpret_cast<uint4*>(dst + 16384) = make_uint4(lo[0] | ((uint32_t)lo[1] << 16), lo[2] | ((uint32_t)lo[3] << 16),
                                                        lo[4] | ((uint32_t)lo[5] << 16), lo[6] | ((uint32_t)lo[7] << 16));
}

// ---- the apply half of the fused small-side step (dense.cu: zside_solve_kernel): Z2 = Z0 M for 32 panel rows per CTA, then,
//      from the same shared-memory tile, (i) Z2 itself, (ii) corr += mu^T Z2 (the centring term of the next A X) and (iii) the
//      pre-split operand of the next A X exactly as tm_prep_x_kernel writes it.  In place (Z2 == Z0) allowed. ----------------
__global__ void __launch_bounds__(256)
tm_zside_apply_kernel(const float* Z0, int64_t n, const float* __restrict__ M, const float* __restrict__ mu,
                      const float* __restrict__ scales, float* Z2, uint8_t* __restrict__ out, double* __restrict__ corr) {
    __shared__ __align__(16) float Ms[LP][LP];
    __shared__ __align__(16) float Pt[LP][32 + 2];        // transposed input tile: Pt[k][row]
    __shared__ float tile[32][LP + 1];
    const int tid = threadIdx.x;
    const int64_t k0 = (int64_t)blockIdx.x * 32;
    for (int i = tid; i < LP * LP; i += 256) Ms[i >> 6][i & 63] = M[i];
    for (int i = tid; i < 32 * LP; i += 256) {
        const int r = i >> 6, c = i & 63;
        Pt[c][r] = (k0 + r < n) ? Z0[(k0 + r) * LP + c] : 0.f;
    }
    __syncthreads();
    {
        const int ty = tid >> 4, tx = tid & 15;           // rows 2 ty, 2 ty + 1; columns 4 tx .. 4 tx + 3
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 8
        for (int kk = 0; kk < LP; kk++) {
            const float2 av = *reinterpret_cast<const float2*>(&Pt[kk][2 * ty]);
            const float4 bv = *reinterpret_cast<const float4*>(&Ms[kk][4 * tx]);
            const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int y = 0; y < 4; y++) {
                acc[0][y] = fmaf(av.x, b[y], acc[0][y]);
                acc[1][y] = fmaf(av.y, b[y], acc[1][y]);
            }
        }
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
            for (int y = 0; y < 4; y++) tile[2 * ty + x][4 * tx + y] = acc[x][y];
    }
    __syncthreads();
    for (int i = tid; i < 32 * LP; i += 256) {
        const int r = i >> 6, c = i & 63;
        if (k0 + r < n) Z2[(k0 + r) * LP + c] = tile[r][c];
    }
    if (mu && tid < LP) {
        double sacc = 0.0;
        for (int r = 0; r < 32; r++)
            if (k0 + r < n) sacc = fma((double)mu[k0 + r], (double)tile[r][tid], sacc);
        if (sacc != 0.0) atomicAdd(&corr[tid], sacc);
    }
    const float s = scales[0];
    const int o = tid >> 6, c = tid & 63;                  // K-octet of the 32 rows, panel column
    const int64_t kc0 = k0 + o * 8;
    const int64_t b = kc0 >> 7;
    const uint32_t chunk = (uint32_t)(kc0 & 127) >> 3;
    unsigned short hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float x = fminf(fmaxf(tile[o * 8 + j][c] * s, -60000.f), 60000.f);   // (finite whatever the factorisation did)
        const __half h = __float2half_rn(x);
        hi[j] = __half_as_ushort(h);
        lo[j] = __half_as_ushort(__float2half_rn(x - __half2float(h)));
    }
    uint8_t* dst = out + (size_t)b * TM_STAGE_BYTES + chunk * 1024u + (uint32_t)(c >> 3) * 128u + (uint32_t)(c & 7) * 16u;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0] | ((uint32_t)hi[1] << 16), hi[2] | ((uint32_t)hi[3] << 16),
                                                hi[4] | ((uint32_t)hi[5] << 16), hi[6] | ((uint32_t)hi[7] << 16));
    *reinterpret_cast<uint4*>(dst + 16384) = make_uint4(lo[0] | ((uint32_t)lo[1] << 16), lo[2] | ((uint32_t)lo[3] << 16),
                                                        lo[4] | ((uint32_t)lo[5] << 16), lo[6] | ((uint32_t)lo[7] << 16));
}

// ---- PTX: TMEM store, MMA with the A operand in TMEM ---------------------------------------------------------------------
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem_d] (+)= A[tmem_a] (128 lanes x 8 columns = K 16 fp16) * B[desc_b]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// eight K-steps in one asm block (the issuing thread is a serial resource of the CTA): A advances 8 TMEM columns, B b_step
// (16 B units) per step; only the first MMA may overwrite the accumulator
__device__ __forceinline__ void umma_ts_run8(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate_first,
                                             uint64_t b_step) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t.reg .b32 ta;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "mov.b64 db, %2;\n\tmov.b32 ta, %1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, p;\n\t"
        "add.u64 db, db, %5;\n\tadd.u32 ta, ta, 8;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, q;\n\t"
        "add.u64 db, db, %5;\n\tadd.u32 ta, ta, 8;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, q;\n\t"
        "add.u64 db, db, %5;\n\tadd.u32 ta, ta, 8;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, q;\n\t"
        "add.u64 db, db, %5;\n\tadd.u32 ta, ta, 8;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, q;\n\t"
        "add.u64 db, db, %5;\n\tadd.u32 ta, ta, 8;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, q;\n\t"
        "add.u64 db, db, %5;\n\tadd.u32 ta, ta, 8;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, q;\n\t"
        "add.u64 db, db, %5;\n\tadd.u32 ta, ta, 8;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, q;\n\t"
        "}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "l"(b_step)
        : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// 16 quads of a lane -> 32 registers: quad i (mask bit) is loaded from a running cursor, absent quads are zero.  Two cursors
// (quads 0-7 and 8-15, the second starts popc(low byte) quads later) halve the serial address chain.
template <bool GLOBAL>
__device__ __forceinline__ void tm_expand16(uint32_t m16, uint32_t& cur_s, const uint2*& cur_g, uint32_t (&r)[32]) {
    const uint32_t n_lo = (uint32_t)__popc(m16 & 0xFFu), n_all = (uint32_t)__popc(m16 & 0xFFFFu);
    uint32_t cs[2] = {cur_s, cur_s + n_lo * 8u};
    const uint2* cg[2] = {cur_g, cur_g + n_lo};
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int q = i + 8 * h;
            const uint32_t bit = m16 & (1u << q);
            if (!GLOBAL) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %3, 0;\n\t"
                    "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
                    "@p ld.shared.v2.b32 {%0, %1}, [%2];\n\t"
                    "@p add.u32 %2, %2, 8;\n\t}"
                    : "=r"(r[2 * q]), "=r"(r[2 * q + 1]), "+r"(cs[h])
                    : "r"(bit));
            } else {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %3, 0;\n\t"
                    "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
                    "@p ld.global.nc.v2.b32 {%0, %1}, [%2];\n\t"
                    "@p add.u64 %2, %2, 8;\n\t}"
                    : "=r"(r[2 * q]), "=r"(r[2 * q + 1]), "+l"(cg[h])
                    : "r"(bit));
            }
        }
    }
    cur_s += n_all * 8u;
    cur_g += n_all;
}

// ---- pass sequences ---------------------------------------------------------------------------------------------------------
// Every role of a CTA walks the same private sequence of PASSES (tile, operator term):
//   A X   : groups = row-block pairs g = blockIdx.x, + gridDim.x, ...; order (pair, column block, row block of the pair, term);
//           the two tiles of a pair share the panel stage of their column block; accumulator sets alternate between pairs
//   A^T Y : groups = work items it = blockIdx.x, + gridDim.x, ... -> (column group g = it mod n_groups, row range);
//           order (row block, column block of the group, term); the tiles of a row block share its Y stage
struct TmPass {
    int64_t tile, block;        // tile in the directory, panel block of the stage
    uint32_t stage, grp;        // running stage index, running group index of this CTA
    int term, acc, set;
    bool stage_first, stage_last, acc_first, grp_first, grp_last, a_last, solo;
};
// (state is advanced incrementally: every role walks the sequence once per pass, two of them are serial resources of the CTA)
template <bool ATY>
struct TmSeq {
    int n_cb, terms, n_mine;
    int n_groups, rb_per_range, n_rb;          // A^T Y
    int i = 0, a = 0, b = 0, term = 0, na = 0, nb = 0;
    uint32_t stage = 0;
    int64_t tile_a = 0;                        // tile of (a, b = 0)
    int rb0 = 0, cb_lo = 0;
    __device__ __forceinline__ void load_item() {
        if (i >= n_mine) return;
        if (!ATY) {
            const int64_t g = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
            tile_a = 2 * g * n_cb;
            na = n_cb;
            nb = 2;
        } else {
            const int it = (int)blockIdx.x + i * (int)gridDim.x;
            const int g = it % n_groups, rg = it / n_groups;
            rb0 = rg * rb_per_range;
            const int rb1 = rb0 + rb_per_range < n_rb ? rb0 + rb_per_range : n_rb;
            cb_lo = g * TM_GC;
            nb = n_cb - cb_lo < TM_GC ? n_cb - cb_lo : TM_GC;
            na = rb1 - rb0;
            tile_a = (int64_t)rb0 * n_cb + cb_lo;
        }
    }
    __device__ __forceinline__ bool valid() const { return i < n_mine; }
    __device__ __forceinline__ int gc() const { return nb; }
    __device__ __forceinline__ TmPass cur() const {
        TmPass p;
        p.term = term;
        p.grp = (uint32_t)i;
        p.stage = stage;
        p.tile = tile_a + (ATY ? b : b * n_cb);
        p.block = ATY ? rb0 + a : a;
        p.acc = ATY ? b : (i & 1) * 2 + b;
        p.set = ATY ? 0 : (i & 1);
        p.stage_first = b == 0 && term == 0;
        p.stage_last = b == nb - 1 && term == terms - 1;
        p.acc_first = a == 0 && term == 0;
        p.grp_first = a == 0 && p.stage_first;
        p.grp_last = a == na - 1 && p.stage_last;
        p.a_last = a == na - 1;
        p.solo = nb == 1;
        return p;
    }
    __device__ __forceinline__ int64_t tile() const { return tile_a + (ATY ? b : b * n_cb); }
    __device__ __forceinline__ int64_t block() const { return ATY ? rb0 + a : a; }
    __device__ __forceinline__ void next_stage() {      // first pass of the next panel stage
        term = 0;
        b = 0;
        stage++;
        tile_a += ATY ? n_cb : 1;
        if (++a < na) return;
        a = 0;
        i++;
        load_item();
    }
    __device__ __forceinline__ void next_group() {      // first pass of the next group
        stage += (uint32_t)(na - a);
        term = 0;
        b = 0;
        a = 0;
        i++;
        load_item();
    }
    __device__ __forceinline__ void next() {
        if (++term < terms) return;
        term = 0;
        if (++b < nb) return;
        b = 0;
        stage++;
        tile_a += ATY ? n_cb : 1;
        if (++a < na) return;
        a = 0;
        i++;
        load_item();
    }
};

#ifndef TM_DBG_
#define TM_DBG_ 0
#endif
__device__ unsigned long long g_tm_dbg[32];      // cycle counters of CTA 0 (experiment builds: -DTM_DBG_=1, SALG_TM_DBG=1)
#define TM_T(acc) do { if (TM_DBG_) { long long _t = clock64(); acc += _t - t_prev; t_prev = _t; } } while (0)

// Every wait of the product kernel goes through tm_wait.  Experiment builds (-DTM_DBG_=1) give up after 2^16 polls, record
// (tag, pass, CTA) of the first stuck waits in g_tm_stuck and let the kernel drain with garbage results: a protocol bug can
// then be read off the host instead of ending in a trap.
__device__ unsigned long long g_tm_stuck[64];
__device__ unsigned g_tm_nstuck;
__device__ __forceinline__ void tm_wait(uint32_t addr, uint32_t parity, uint32_t tag, uint32_t p, volatile int* s_abort) {
#if TM_DBG_
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok) {
            if (*s_abort) return;
            if (++spins > (1u << 16)) {
                *s_abort = 1;
                if ((threadIdx.x & 31) == 0) {
                    const unsigned k = atomicAdd(&g_tm_nstuck, 1u);
                    if (k < 64) g_tm_stuck[k] = ((unsigned long long)tag << 56) | ((unsigned long long)(blockIdx.x & 0xFFFF) << 40) |
                                                ((unsigned long long)(threadIdx.x >> 5) << 32) | p;
                }
                return;
            }
        }
    } while (!ok);
#else
    mbar_wait_a(addr, parity);
#endif
}

struct TmArgs {
    const uint64_t* info;
    const uint2* q_hi;
    const uint2* q_lo;
    int n_rb, n_cb, terms;
    int n_units;              // A X: row-block pairs; A^T Y: work items
    int n_groups, rb_per_range, n_rb_real;
    int64_t n_out;            // A X: rows of Y; A^T Y: rows of Z (operator columns)
    const uint8_t* prep;      // pre-split panel blocks (32 KB each)
    const float* scales;
    float* out;
    const double* corr;       // A X
    unsigned* amax_out;       // A X
    int dbg;                  // timing experiments (SALG_TM_DBG bits; results are wrong with a bit >= 2 set)
    int b_terms;              // A X: fp16 terms of the panel that are multiplied (2 = 22 significant bits, 1 = 11: experiment)
};

// The serial roles (MMA issuers, loaders) run as CONVERGED warps: every lane walks the same loop, waits on the same barrier,
// and one elected lane issues the tcgen05 / bulk-copy instruction.  Measured on B200: as `if (lane == 0)` single threads they
// ran at ~6 clk per instruction with an ELECT + R2UR loop in front of every tcgen05.mma (the compiler cannot keep a divergent
// thread's descriptors in uniform registers) and were the bottleneck of the whole kernel.  Their loops therefore use hoisted
// shared-memory addresses, power-of-two rings and incremental counters: every instruction of these warps is on the critical path.
template <bool ATY>
__global__ void __launch_bounds__(TM_THREADS, 1)
tm_product_kernel(const TmArgs g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sStage = smem_raw;
    uint8_t* sRing = sStage + TM_NB * TM_STAGE_BYTES;
    uint8_t* sEpi = sRing + TM_NS * TM_SLOT_BYTES;
    __shared__ uint64_t e_full[TM_NS], e_free[TM_NS], a_full[TM_NA], a_free[TM_NA], d_full[TM_NB], d_free[TM_NB], acc_full[2], acc_free[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_corr[LP];
    __shared__ int s_abort_flag;
    volatile int* s_abort = &s_abort_flag;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);       // provably warp-uniform
    if (tid == 0) s_abort_flag = 0;
    TmSeq<ATY> seq;
    seq.n_cb = g.n_cb;
    seq.terms = g.terms;
    seq.n_mine = ((int)blockIdx.x < g.n_units) ? (g.n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    seq.n_groups = g.n_groups;
    seq.rb_per_range = g.rb_per_range;
    seq.n_rb = g.n_rb_real;
    seq.load_item();

    if (tid == 0) {
        for (int i = 0; i < TM_NS; i++) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_free[i], 4);
        }
        for (int i = 0; i < TM_NA; i++) {
            mbar_init(&a_full[i], 4);
            mbar_init(&a_free[i], 1);
        }
        for (int i = 0; i < TM_NB; i++) {
            mbar_init(&d_full[i], 1);
            mbar_init(&d_free[i], 2);        // one commit per MMA issuer
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&acc_full[i], 2);      // one commit per MMA issuer
            mbar_init(&acc_free[i], 4);
        }
        fence_barrier_init();
    }
    if (tid < LP) s_corr[tid] = (!ATY && g.corr) ? (float)g.corr[tid] : 0.f;
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t ring0 = smem_u32(sRing), stage0 = smem_u32(sStage);
    const uint32_t b_efull = smem_u32(e_full), b_efree = smem_u32(e_free), b_afull = smem_u32(a_full), b_afree = smem_u32(a_free),
                   b_dfull = smem_u32(d_full), b_dfree = smem_u32(d_free), b_accfull = smem_u32(acc_full), b_accfree = smem_u32(acc_free);

    if (warp < TM_BUILD_WARPS) {
        // ================= expanders: compressed lane -> registers -> TMEM =================
        const int set = warp >> 2, qd = warp & 3;
        const int row = qd * 32 + lane;
        uint32_t p = 0;
        long long c_ef = 0, c_x0 = 0, c_af = 0, c_x1 = 0, c_sw = 0, c_n = 0, t_prev = TM_DBG_ ? clock64() : 0, t_begin = t_prev;
        for (; seq.valid(); seq.next(), p++) {
            if ((int)(p & (TM_SETS - 1)) != set) continue;
            const uint32_t slot = p % TM_NS, ab = p & (TM_NA - 1);
            const bool lo_term = seq.term != 0;
            tm_wait(b_efull + slot * 8u, (p / TM_NS) & 1, 1, p, s_abort);
            if (TM_DBG_ && *s_abort) break;
            TM_T(c_ef);
            const uint8_t* sl = sRing + (size_t)slot * TM_SLOT_BYTES;
            const uint4 meta = *reinterpret_cast<const uint4*>(sl);                 // first payload quad (u64), quad count
            const uint32_t mask = reinterpret_cast<const uint32_t*>(sl + 16)[row];
            const uint32_t off = reinterpret_cast<const unsigned short*>(sl + 16 + 512)[row];
            const bool in_smem = meta.z <= (uint32_t)TM_SLOT_QUADS;
            const uint32_t taddr = tmem_base + TM_A_COL0 + ab * 64u + ((uint32_t)(qd * 32) << 16);
            uint32_t r[32];
            uint32_t cur_s = ring0 + slot * (uint32_t)TM_SLOT_BYTES + TM_REC + off * 8u;
            const uint2* cur_g = (lo_term ? g.q_lo : g.q_hi) + (((uint64_t)meta.y << 32) | meta.x) + off;
            const uint32_t xmask = (g.dbg & 16) ? 0u : mask;                  // (timing experiment: nothing to expand)
            if (in_smem) tm_expand16<false>(xmask & 0xFFFFu, cur_s, cur_g, r);
            else tm_expand16<true>(xmask & 0xFFFFu, cur_s, cur_g, r);
            TM_T(c_x0);
            if (p >= TM_NA) tm_wait(b_afree + ab * 8u, ((p / TM_NA) - 1) & 1, 2, p, s_abort);
            if (TM_DBG_ && *s_abort) break;     // the MMAs that read this buffer have retired
            TM_T(c_af);
            tc_fence_after();
            tmem_st32(taddr, r);
            if (in_smem) tm_expand16<false>(xmask >> 16, cur_s, cur_g, r);
            else tm_expand16<true>(xmask >> 16, cur_s, cur_g, r);
            tmem_st32(taddr + 32u, r);
            TM_T(c_x1);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_a(b_afull + ab * 8u);
                mbar_arrive_a(b_efree + slot * 8u);
            }
            TM_T(c_sw);
            c_n++;
        }
        if (TM_DBG_ && blockIdx.x == 0 && warp == 0 && lane == 0) {
            g_tm_dbg[0] = clock64() - t_begin; g_tm_dbg[1] = c_ef; g_tm_dbg[2] = c_x0; g_tm_dbg[3] = c_af; g_tm_dbg[4] = c_x1;
            g_tm_dbg[5] = c_sw; g_tm_dbg[6] = c_n;
        }
    } else if (warp >= TM_W_ELOAD) {
        // ================= tile loaders: loader w serves the passes p = w (mod 2), one bulk copy per pass =================
        if (seq.valid()) {
            const uint32_t me = (uint32_t)(warp - TM_W_ELOAD);
            // directory entries of this loader's passes are gathered 32 at a time, one per lane, a batch ahead (a second
            // iterator): one global-memory latency per 32 passes instead of one per pass
            TmSeq<ATY> pre = seq;
            if (me) pre.next();
            auto gather = [&]() -> uint64_t {
                int64_t my_tile = -1;
                for (int j = 0; j < 32 && pre.valid(); j++) {
                    const int64_t t = pre.tile();
                    if (lane == j) my_tile = t;
                    pre.next();
                    if (pre.valid()) pre.next();
                }
                return my_tile >= 0 ? g.info[my_tile] : 0ull;
            };
            uint64_t inf_cur = gather(), inf_nxt = gather();
            uint32_t p = 0, own = 0;
            long long c_w = 0, c_r = 0, t_prev = TM_DBG_ ? clock64() : 0;
            for (; seq.valid(); seq.next(), p++) {
                if ((p & 1u) != me) continue;
                const uint2* qs = seq.term ? g.q_lo : g.q_hi;
                if (own && (own & 31u) == 0) {
                    inf_cur = inf_nxt;
                    inf_nxt = gather();
                }
                const uint64_t inf = __shfl_sync(0xFFFFFFFFu, inf_cur, (int)(own & 31u));
                own++;
                const uint32_t slot = p % TM_NS;
                TM_T(c_r);
                if (p >= TM_NS) tm_wait(b_efree + slot * 8u, ((p / TM_NS) - 1) & 1, 3, p, s_abort);
                if (TM_DBG_ && *s_abort) break;
                TM_T(c_w);
                const uint32_t cnt = (uint32_t)(inf >> 40);
                uint32_t bytes = TM_REC + (cnt <= (uint32_t)TM_SLOT_QUADS ? ((cnt + 1u) & ~1u) * 8u : 0u);
                if (g.dbg & 4) bytes = TM_REC;                                  // (timing experiment: record prefixes only)
                // L2 prefetch of the record TM_PF of this loader's passes ahead: a ring slot is then held for an L2 latency
                // instead of a DRAM latency (the ring holds 8 passes; measured slot lifetime ~5600 clk without the prefetch)
                const uint32_t pf_idx = ((own - 1u) & 31u) + TM_PF;
                const uint64_t pf_inf = __shfl_sync(0xFFFFFFFFu, pf_idx < 32u ? inf_cur : inf_nxt, (int)(pf_idx & 31u));
                const uint32_t pf_cnt = (uint32_t)(pf_inf >> 40);
                if (elect_one()) {
                    mbar_expect_tx_a(b_efull + slot * 8u, bytes);
                    bulk_g2s_a(ring0 + slot * (uint32_t)TM_SLOT_BYTES, qs + (inf & ((1ull << 40) - 1)), bytes, b_efull + slot * 8u);
                    if (pf_inf && !(g.dbg & 32))
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(qs + (pf_inf & ((1ull << 40) - 1))),
                                     "r"(TM_REC + ((pf_cnt + 1u) & ~1u) * 8u)
                                     : "memory");
                }
                __syncwarp();
            }
            if (TM_DBG_ && blockIdx.x == 0 && me == 0 && lane == 0) { g_tm_dbg[8] = c_w; g_tm_dbg[9] = c_r; g_tm_dbg[10] = p; }
        }
    } else if (warp == TM_W_PLOAD) {
        // ================= panel-stage loader =================
        {
            long long c_w = 0, c_r = 0, t_prev = TM_DBG_ ? clock64() : 0;
            uint32_t bb = 0, use = 0;                      // stage ring position, completed laps
            for (; seq.valid(); seq.next_stage()) {
                TM_T(c_r);
                if (use) tm_wait(b_dfree + bb * 8u, (use - 1) & 1, 4, seq.stage, s_abort);
                if (TM_DBG_ && *s_abort) break;
                TM_T(c_w);
                if (elect_one()) {
                    if (g.dbg & 2) {                                      // (timing experiment: no panel loads)
                        mbar_arrive_a(b_dfull + bb * 8u);
                    } else {
                        mbar_expect_tx_a(b_dfull + bb * 8u, TM_STAGE_BYTES);
                        bulk_g2s_a(stage0 + bb * (uint32_t)TM_STAGE_BYTES, g.prep + (size_t)seq.block() * TM_STAGE_BYTES, TM_STAGE_BYTES,
                                   b_dfull + bb * 8u);
                        // the Y row blocks of a range are consecutive: prefetch the block 4 stages ahead into L2
                        if (ATY && seq.a + 4 < seq.na && !(g.dbg & 32))
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g.prep + (size_t)(seq.block() + 4) * TM_STAGE_BYTES),
                                         "r"((uint32_t)TM_STAGE_BYTES)
                                         : "memory");
                    }
                }
                __syncwarp();
                if (++bb == TM_NB) {
                    bb = 0;
                    use++;
                }
            }
            if (TM_DBG_ && blockIdx.x == 0 && lane == 0) { g_tm_dbg[12] = c_w; g_tm_dbg[13] = c_r; }
        }
    } else if (warp == TM_W_MMA || warp == TM_W_MMA + 1) {
        // ================= MMA issuers: issuer w owns the accumulators acc = w (mod 2) =================
        // (A X: the row block of a pair, A^T Y: the column block of a group.)  Each walks ITS passes with plain nested loops:
        //   A X   pass p = ((stage * 2) + w) * terms + term,                 stage = group * n_cb + column block
        //   A^T Y pass p = p0(item) + ((row block * gc) + w) * terms + term, stage = s0(item) + row block
        {
            const uint32_t me = (uint32_t)(warp - TM_W_MMA);
            constexpr uint32_t NSET = ATY ? 1 : 2;
            constexpr uint32_t idesc = tc_idesc(ATY ? 128 : 64);
            const uint64_t b_desc0 = ATY ? umma_desc(stage0, 2048, 128) : umma_desc(stage0, 1024, 128);
            const uint32_t terms = (uint32_t)g.terms;
            const bool no_mma = (g.dbg & 8) != 0;                                 // (timing experiment: plain arrivals)
            long long c_acc = 0, c_d = 0, c_a = 0, c_is = 0, c_cm = 0, c_sq = 0, t_prev = TM_DBG_ ? clock64() : 0, t_begin = t_prev;
            uint32_t p0 = 0, bb = 0, dlap = 0, n_own = 0;          // passes before this group, stage ring position / laps
            for (int i = 0; i < seq.n_mine && !(TM_DBG_ && *s_abort); i++) {
                int na, nb;
                if (!ATY) {
                    na = g.n_cb;
                    nb = 2;
                } else {
                    const int it = (int)blockIdx.x + i * (int)gridDim.x;
                    const int gi = it % g.n_groups, rg = it / g.n_groups;
                    const int rb0 = rg * g.rb_per_range;
                    const int rb1 = rb0 + g.rb_per_range < g.n_rb_real ? rb0 + g.rb_per_range : g.n_rb_real;
                    na = rb1 - rb0;
                    nb = g.n_cb - gi * TM_GC < TM_GC ? g.n_cb - gi * TM_GC : TM_GC;
                }
                const bool mine = me < (uint32_t)nb, solo = nb == 1;
                const uint32_t set = ATY ? 0u : ((uint32_t)i & 1u);
                const uint32_t acc = ATY ? me : set * 2u + me;
                const uint32_t d_tmem = tmem_base + acc * (ATY ? 128u : 64u);
                // the epilogue has drained the accumulators of the group NSET before this one.  EVERY issuer waits, also one
                // without passes in this group (a column group with a single block): skipping would let it run a whole group
                // ahead, where the parity of its next wait aliases with the phase before
                if ((uint32_t)i >= NSET) {
                    tm_wait(b_accfree + set * 8u, (((uint32_t)i / NSET) - 1) & 1, 5, (uint32_t)i, s_abort);
                    tc_fence_after();
                }
                TM_T(c_acc);
                for (int a = 0; a < na && !(TM_DBG_ && *s_abort); a++) {
                    if (mine) {
                        tm_wait(b_dfull + bb * 8u, dlap & 1, 6, p0 + (uint32_t)a, s_abort);
                        TM_T(c_d);
                        const uint64_t bd = b_desc0 + (uint64_t)bb * (TM_STAGE_BYTES >> 4);
                        uint32_t p = p0 + ((uint32_t)a * (uint32_t)nb + me) * terms;
                        for (uint32_t term = 0; term < terms && !(TM_DBG_ && *s_abort); term++, p++) {
                            const uint32_t ab = p & (TM_NA - 1);
                            tm_wait(b_afull + ab * 8u, (p / TM_NA) & 1, 7, p, s_abort);
                            tc_fence_after();
                            TM_T(c_a);
                            const uint32_t a_tmem = tmem_base + TM_A_COL0 + ab * 64u;
                            const uint32_t first = (a == 0 && term == 0) ? 0u : 1u;
                            if (elect_one()) {
                                if (no_mma) {
                                    mbar_arrive_a(b_afree + ab * 8u);
                                } else {
                                    if (!ATY) {
                                        umma_ts_run8(d_tmem, a_tmem, bd, idesc, first, 128);
                                        if (g.b_terms > 1) umma_ts_run8(d_tmem, a_tmem, bd + 1024u, idesc, 1u, 128);
                                    } else {
                                        umma_ts_run8(d_tmem, a_tmem, bd, idesc, first, 256);
                                    }
                                    umma_commit_a(b_afree + ab * 8u);
                                }
                            }
                            __syncwarp();
                            TM_T(c_is);
                            n_own++;
                        }
                        // this issuer's last pass of the stage / of the group (a group with one accumulator: both arrivals)
                        if (elect_one()) {
                            if (no_mma) {
                                mbar_arrive_a(b_dfree + bb * 8u);
                                if (solo) mbar_arrive_a(b_dfree + bb * 8u);
                                if (a == na - 1) {
                                    mbar_arrive_a(b_accfull + set * 8u);
                                    if (solo) mbar_arrive_a(b_accfull + set * 8u);
                                }
                            } else {
                                umma_commit_a(b_dfree + bb * 8u);
                                if (solo) umma_commit_a(b_dfree + bb * 8u);
                                if (a == na - 1) {
                                    umma_commit_a(b_accfull + set * 8u);
                                    if (solo) umma_commit_a(b_accfull + set * 8u);
                                }
                            }
                        }
                        __syncwarp();
                        TM_T(c_cm);
                    }
                    if (++bb == TM_NB) {
                        bb = 0;
                        dlap++;
                    }
                }
                p0 += (uint32_t)na * (uint32_t)nb * terms;
            }
            if (TM_DBG_ && blockIdx.x == 0 && me == 0 && lane == 0) {
                g_tm_dbg[16] = clock64() - t_begin; g_tm_dbg[17] = c_acc; g_tm_dbg[18] = c_d; g_tm_dbg[19] = c_a; g_tm_dbg[20] = c_is;
                g_tm_dbg[21] = c_cm; g_tm_dbg[22] = c_sq; g_tm_dbg[23] = n_own;
            }
        }
    } else if (warp >= TM_W_EPI && warp < TM_W_EPI + 4) {
        // ================= epilogue =================
        const int qd = warp & 3;
        const float inv = g.scales[1];
        float amax = 0.f;
        for (; seq.valid(); seq.next_group()) {
            const uint32_t grp = (uint32_t)seq.i;
            const int set = ATY ? 0 : (int)(grp & 1u);
            constexpr uint32_t NSET = ATY ? 1 : 2;
            const int cb_lo = seq.cb_lo, gc = seq.gc();
            tm_wait(b_accfull + set * 8u, (grp / NSET) & 1, 8, grp, s_abort);
            if (TM_DBG_ && *s_abort) break;
            tc_fence_after();
            if (!ATY) {
                const int64_t pair = (int64_t)blockIdx.x + (int64_t)grp * gridDim.x;
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    const int64_t rowi = (2 * pair + h) * TM_LANES + qd * 32 + lane;
                    const uint32_t t0 = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((grp & 1u) * 2u + h) * 64u;
#pragma unroll 1
                    for (int c = 0; c < 2; c++) {
                        // lane = row: 32 accumulator columns -> staging (row pitch 144 B), then the warp stores four rows x
                        // 128 contiguous bytes per instruction (a lane writing its own row touched 32 lines per instruction:
                        // measured 13 % of the kernel)
                        uint32_t v[32];
                        tmem_ld32(t0 + c * 32, v);
                        uint8_t* stg = sEpi + (size_t)qd * (32 * TM_EPI_PITCH);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 y;
                            y.x = __uint_as_float(v[4 * j]) * inv - s_corr[c * 32 + 4 * j];
                            y.y = __uint_as_float(v[4 * j + 1]) * inv - s_corr[c * 32 + 4 * j + 1];
                            y.z = __uint_as_float(v[4 * j + 2]) * inv - s_corr[c * 32 + 4 * j + 2];
                            y.w = __uint_as_float(v[4 * j + 3]) * inv - s_corr[c * 32 + 4 * j + 3];
                            if (rowi < g.n_out) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w))));
                            *reinterpret_cast<float4*>(stg + lane * TM_EPI_PITCH + j * 16) = y;
                        }
                        __syncwarp();
                        const int64_t row_w = rowi - lane;                       // first row of this warp
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int rr = i * 4 + (lane >> 3);
                            const float4 y = *reinterpret_cast<const float4*>(stg + rr * TM_EPI_PITCH + (lane & 7) * 16);
                            if (row_w + rr < g.n_out && !(g.dbg & 64))
                                *reinterpret_cast<float4*>(g.out + (row_w + rr) * LP + c * 32 + (lane & 7) * 4) = y;
                        }
                        __syncwarp();
                    }
                }
            } else {
#pragma unroll 1
                for (int lc = 0; lc < gc; lc++) {
                    const int64_t cA = (int64_t)(cb_lo + lc) * TM_LANES + qd * 32 + lane;       // operator column of this lane
                    const uint32_t t0 = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)lc * 128u;
#pragma unroll 1
                    for (int c = 0; c < 4; c++) {
                        uint32_t v[32];
                        tmem_ld32(t0 + c * 32, v);
                        if (cA < g.n_out) {
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                const float x = (__uint_as_float(v[2 * j]) + __uint_as_float(v[2 * j + 1])) * inv;
                                if (x != 0.f) atomicAdd(g.out + cA * LP + c * 16 + j, x);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(b_accfree + set * 8u);
        }
        if (!ATY && g.amax_out) {
#pragma unroll
            for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
            if (lane == 0 && amax > 0.f) atomicMax(g.amax_out, __float_as_uint(amax));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
void tc_panel_scales(salg_ctx* ctx, const float* P, int64_t n, float a_scale, float* d_scales, unsigned* d_amax);   // tc.cu

static void tm_dbg_print(salg_ctx* ctx, const char* what) {
    if (!TM_DBG_ || !getenv("SALG_TM_DBG")) return;
    unsigned long long h[32];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpyFromSymbol(h, g_tm_dbg, sizeof(h));
    fprintf(stderr, "[tm %s] expander(set 0) total %llu passes %llu: wait e_full %llu expand0 %llu wait a_free %llu st+expand1 %llu st-wait+arrive %llu\n",
            what, h[0], h[6], h[1], h[2], h[3], h[4], h[5]);
    fprintf(stderr, "[tm %s] mma total %llu passes %llu: seq %llu acc_free %llu d_full %llu a_full %llu issue %llu commit %llu | loader wait e_free %llu rest %llu | panel wait d_free %llu rest %llu\n",
            what, h[16], h[23], h[22], h[17], h[18], h[19], h[20], h[21], h[8], h[9], h[12], h[13]);
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_tm_dbg, z, sizeof(z));
    unsigned ns = 0;
    unsigned long long st[64];
    cudaMemcpyFromSymbol(&ns, g_tm_nstuck, 4);
    cudaMemcpyFromSymbol(st, g_tm_stuck, sizeof(st));
    for (unsigned k = 0; k < ns && k < 64; k++)
        fprintf(stderr, "[tm %s] STUCK wait tag %llu (1 e_full 2 a_free 3 e_free 4 d_free 5 acc_free 6 d_full 7 a_full 8 acc_full) cta %llu warp %llu pass/idx %llu\n",
                what, st[k] >> 56, (st[k] >> 40) & 0xFFFF, (st[k] >> 32) & 0xFF, st[k] & 0xFFFFFFFFull);
    ns = 0;
    cudaMemcpyToSymbol(g_tm_nstuck, &ns, 4);
}

size_t tm_xprep_bytes(const void* tiles) { return (size_t)((const TmTiles*)tiles)->n_cb * TM_STAGE_BYTES; }

// Y (nrows x 64) = A X - 1 corr^T with X given pre-split (tm_prep_x_kernel / tm_zside_apply_kernel layout, scales = {s, 1 / (s a_scale)});
// d_amax (optional, zeroed here) receives the bits of max |Y|
void tm_spmm_A_prepped(salg_ctx* ctx, const salg_csr* c, void* tiles, const uint8_t* Xprep, const float* scales, float* Y,
                       const double* corr, unsigned* d_amax, int b_terms) {
    cudaStream_t st = ctx->stream;
    TmTiles* t = (TmTiles*)tiles;
    if (d_amax) SALG_CUDA(cudaMemsetAsync(d_amax, 0, 4, st));
    TmArgs a{};
    a.info = t->R.info;
    a.q_hi = t->R.q_hi;
    a.q_lo = t->R.q_lo;
    a.n_rb = t->n_rb;
    a.n_cb = t->n_cb;
    a.terms = t->a_terms;
    const int n_pairs = (int)ceil_div(ceil_div(c->nrows, TM_LANES), 2);
    a.n_units = n_pairs;
    a.n_out = c->nrows;
    a.prep = Xprep;
    a.scales = scales;
    a.out = Y;
    a.corr = corr;
    a.amax_out = d_amax;
    a.b_terms = b_terms;
    a.dbg = getenv("SALG_TM_DBG") ? atoi(getenv("SALG_TM_DBG")) : 0;
    set_max_dyn_smem(tm_product_kernel<false>, TM_SMEM);
    const int grid = std::min(n_pairs, std::max(1, ctx->sm_count - ctx->sm_reserve));
    tm_product_kernel<false><<<grid, TM_THREADS, TM_SMEM, st>>>(a);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    tm_dbg_print(ctx, "ax");
}

// Y (nrows x 64) = A X - 1 corr^T; d_amax (optional, zeroed here) receives the bits of max |Y|
void tm_spmm_A(salg_ctx* ctx, const salg_csr* c, void* tiles, const float* X, float* Y, const double* corr, unsigned* d_amax, int b_terms) {
    cudaStream_t st = ctx->stream;
    TmTiles* t = (TmTiles*)tiles;
    DevBuf<uint8_t> Xprep(tm_xprep_bytes(t), st);
    DevBuf<float> scales(2, st);
    DevBuf<unsigned> amax(1, st);
    tc_panel_scales(ctx, X, c->ncols, t->a_scale, scales.get(), amax.get());
    tm_prep_x_kernel<<<(unsigned)(t->n_cb * (TM_DEPTH / 32)), 256, 0, st>>>(X, c->ncols, scales.get(), Xprep.get());
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    tm_spmm_A_prepped(ctx, c, tiles, Xprep.get(), scales.get(), Y, corr, d_amax, b_terms);
}

// the apply half of the fused small-side step: Z (n_eff x 64) <- Z M, corr (zeroed by zside_solve) += mu^T Z, Xprep = pre-split Z
void tm_zside_apply(salg_ctx* ctx, const salg_csr* c, void* tiles, float* Z, const float* d_M, const float* mu, const float* scales,
                    uint8_t* Xprep, double* corr) {
    TmTiles* t = (TmTiles*)tiles;
    if (t->n_cb == 0) return;
    tm_zside_apply_kernel<<<(unsigned)(t->n_cb * (TM_DEPTH / 32)), 256, 0, ctx->stream>>>(Z, c->ncols, d_M, mu, scales, Z, Xprep, corr);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
}

// Z += A^T Y (Z pre-initialised with the centring term), Y given pre-split as tc_gram_prep / tc_prep_kernel<128> write it
void tm_aty_launch(salg_ctx* ctx, const salg_csr* c, void* tiles, const uint8_t* Yprep, const float* scales, float* Z) {
    cudaStream_t st = ctx->stream;
    TmTiles* t = (TmTiles*)tiles;
    const int n_rb_real = (int)ceil_div(c->nrows, TM_LANES);
    if (n_rb_real == 0 || t->n_cb == 0) return;
    const int n_groups = (int)ceil_div(t->n_cb, TM_GC);
    // work items (column group, row range): enough of them to balance the SMs, each long enough to amortise its flush
    const int sms = ctx->sm_count;
    int best_ranges = 1;
    double best_cost = 1e300;
    const int max_ranges = std::max(1, std::min(n_rb_real, std::max(1, n_rb_real / 4)));
    for (int ranges = 1; ranges <= max_ranges; ranges++) {
        const int rpr = (int)ceil_div(n_rb_real, ranges);
        const int rr = (int)ceil_div(n_rb_real, rpr);
        if (rr != ranges) continue;
        const int64_t items = (int64_t)n_groups * rr;
        const int64_t waves = ceil_div(items, sms);
        const double cost = (double)waves * (rpr + 2.0);            // (+2: flush and pipeline fill of an item, in row blocks)
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best_ranges = rr;
        }
        if (items > (int64_t)sms * 64) break;
    }
    const int rb_per_range = (int)ceil_div(n_rb_real, best_ranges);
    const int ranges = (int)ceil_div(n_rb_real, rb_per_range);
    TmArgs a{};
    a.info = t->T.info;
    a.q_hi = t->T.q_hi;
    a.q_lo = t->T.q_lo;
    a.n_rb = t->n_rb;
    a.n_cb = t->n_cb;
    a.terms = t->a_terms;
    a.n_units = n_groups * ranges;
    a.n_groups = n_groups;
    a.rb_per_range = rb_per_range;
    a.n_rb_real = n_rb_real;
    a.n_out = c->ncols;
    a.prep = Yprep;
    a.scales = scales;
    a.out = Z;
    a.dbg = getenv("SALG_TM_DBG") ? atoi(getenv("SALG_TM_DBG")) : 0;
    set_max_dyn_smem(tm_product_kernel<true>, TM_SMEM);
    const int grid = std::min(a.n_units, sms);
    tm_product_kernel<true><<<grid, TM_THREADS, TM_SMEM, st>>>(a);
    ctx->n_launch++;
    SALG_CUDA(cudaGetLastError());
    tm_dbg_print(ctx, "aty");
}

}  // namespace salg
