// api.cu — context lifecycle, error reporting, profiling records, NCCL plumbing.
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include "common.cuh"

namespace salg {

static thread_local std::string g_last_error;
static std::mutex g_ctx_mutex;
static std::set<const salg_ctx*> g_live_ctx;

bool ctx_alive(const salg_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    return ctx && g_live_ctx.count(ctx) != 0;
}

void set_last_error(const std::string& msg) { g_last_error = msg; }

static const char* kProfNames[PROF_NCLS] = {
    "spmm", "spmm_t", "gram", "chol", "panel_mul", "jacobi", "stats", "transpose",
    "compact", "elementwise", "allreduce", "h2d", "spmv", "other"};

ProfScope::ProfScope(salg_ctx* c, int cls, double bytes) : ctx(c) {
    if (!ctx->prof_on) return;
    if (ctx->prof_products_only && cls != PROF_SPMM && cls != PROF_SPMMT) return;
    ProfRecord r;
    r.cls = cls;
    r.bytes = bytes;
    auto get_event = [&]() {
        cudaEvent_t e;
        if (!ctx->event_pool.empty()) {
            e = ctx->event_pool.back();
            ctx->event_pool.pop_back();
        } else {
            if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;
        }
        return e;
    };
    r.e0 = get_event();
    r.e1 = get_event();
    if (!r.e0 || !r.e1) return;
    cudaEventRecord(r.e0, ctx->stream);
    ctx->prof_pending.push_back(r);
    idx = (int)ctx->prof_pending.size() - 1;
}

ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(ctx->prof_pending[idx].e1, ctx->stream);
}

void prof_collect(salg_ctx* ctx) {
    for (auto& r : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess &&
            cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            ctx->prof_ms[r.cls] += ms;
            ctx->prof_launches[r.cls] += 1;
            ctx->prof_bytes[r.cls] += r.bytes;
        }
        ctx->event_pool.push_back(r.e0);
        ctx->event_pool.push_back(r.e1);
    }
    ctx->prof_pending.clear();
}

namespace {
constexpr size_t SMALL_MAX = (size_t)1 << 20;
constexpr int SMALL_CLASSES = 21;                                  // 2^0 .. 2^20 bytes
struct SmallCache { std::vector<void*> free_list[SMALL_CLASSES]; };
std::mutex g_small_mu;
std::map<cudaStream_t, SmallCache> g_small;
inline int small_class(size_t bytes) {
    int c = 8;                                                     // at least 256 B
    while (((size_t)1 << c) < bytes) c++;
    return c;
}
}  // namespace

void* small_alloc(cudaStream_t s, size_t bytes) {
    void* p = nullptr;
    if (bytes <= SMALL_MAX) {
        const int c = small_class(bytes);
        {
            std::lock_guard<std::mutex> lk(g_small_mu);
            auto& fl = g_small[s].free_list[c];
            if (!fl.empty()) {
                p = fl.back();
                fl.pop_back();
                return p;
            }
        }
        SALG_CUDA(cudaMallocAsync(&p, (size_t)1 << c, s));
        return p;
    }
    SALG_CUDA(cudaMallocAsync(&p, bytes, s));
    return p;
}

void small_free(cudaStream_t s, void* p, size_t bytes) {
    if (!p) return;
    if (bytes <= SMALL_MAX) {
        std::lock_guard<std::mutex> lk(g_small_mu);
        auto it = g_small.find(s);
        if (it != g_small.end()) {                                  // (stream already torn down: fall through)
            it->second.free_list[small_class(bytes)].push_back(p);
            return;
        }
    }
    cudaFreeAsync(p, s);
}

void small_cache_release(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_small_mu);
    auto it = g_small.find(s);
    if (it == g_small.end()) return;
    for (auto& fl : it->second.free_list)
        for (void* p : fl) cudaFreeAsync(p, s);
    g_small.erase(it);
}

void set_max_dyn_smem_impl(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> done;      // (kernel, device) -> bytes already granted
    int dev = 0;
    SALG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    int& have = done[{kernel, dev}];
    if (have < bytes) {
        SALG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        have = bytes;
    }
}

void allreduce_f64(salg_ctx* ctx, double* buf, size_t n) {
    if (ctx->nranks <= 1 || n == 0) return;
    ProfScope ps(ctx, PROF_ALLREDUCE, (double)n * 8);
    if (p2p_allreduce(ctx, buf, n, nullptr, 0)) return;
    SALG_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, ctx->comm, ctx->stream));
}

// one NCCL launch for the pair (Gram + column sums in f64, partial A^T Y panel in f32) of a power-iteration half step
void allreduce_gram_and_panel(salg_ctx* ctx, double* gram, size_t n_gram, float* panel, size_t n_panel) {
    if (ctx->nranks <= 1) return;
    ProfScope ps(ctx, PROF_ALLREDUCE, (double)n_gram * 8 + (double)n_panel * 4);
    if (p2p_allreduce(ctx, gram, n_gram, panel, n_panel)) return;
    SALG_NCCL(ncclGroupStart());
    if (n_gram) SALG_NCCL(ncclAllReduce(gram, gram, n_gram, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    if (n_panel) SALG_NCCL(ncclAllReduce(panel, panel, n_panel, ncclFloat, ncclSum, ctx->comm, ctx->stream));
    SALG_NCCL(ncclGroupEnd());
}

template <>
void allreduce_T<float>(salg_ctx* ctx, float* buf, size_t n) {
    if (ctx->nranks <= 1 || n == 0) return;
    ProfScope ps(ctx, PROF_ALLREDUCE, (double)n * 4);
    if (p2p_allreduce(ctx, nullptr, 0, buf, n)) return;
    SALG_NCCL(ncclAllReduce(buf, buf, n, ncclFloat, ncclSum, ctx->comm, ctx->stream));
}
template <>
void allreduce_T<double>(salg_ctx* ctx, double* buf, size_t n) {
    allreduce_f64(ctx, buf, n);
}

static salg_ctx* ctx_new(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw Error(SALG_ERR_CUDA,
                    std::string("no CUDA device available (libsalg_b200 has no CPU fallback): ") +
                        cudaGetErrorString(e));
    SALG_REQUIRE(device >= 0 && device < ndev, SALG_ERR_BAD_ARG, "device index out of range");
    SALG_CUDA(cudaSetDevice(device));
    salg_ctx* c = new salg_ctx();
    c->device = device;
    cudaDeviceProp prop;
    SALG_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    {
        const char* e = getenv("SALG_SPMM_IMPL");
        // default: tcgen05 products with the sparse operand expanded into TMEM (tm.cu); "tc" = the dense-tile generation
        // (tc.cu, also the fallback for operators wider than the TMEM-operand builder supports), "chunk" = CUDA cores
        c->spmm_impl = (e && strcmp(e, "chunk") == 0) ? 1 : (e && strcmp(e, "tc") == 0) ? 0 : 2;
    }
    SALG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SALG_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    // keep freed temporaries cached in the stream-ordered pool
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    {
        std::lock_guard<std::mutex> lk(g_ctx_mutex);
        g_live_ctx.insert(c);
    }
    return c;
}

}  // namespace salg

using namespace salg;

extern "C" {

const char* salg_last_error(void) { return g_last_error.c_str(); }

int salg_version(void) { return SALG_VERSION; }

int salg_device_count(int* out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        *out = n;
    });
}

int salg_ctx_create(int device, salg_ctx** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        *out = ctx_new(device);
    });
}

int salg_nccl_unique_id(void* out128) {
    return guarded([&] {
        SALG_REQUIRE(out128, SALG_ERR_BAD_ARG, "out is NULL");
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
        ncclUniqueId id;
        SALG_NCCL(ncclGetUniqueId(&id));
        memcpy(out128, &id, sizeof(id));
    });
}

int salg_ctx_create_dist(int device, int rank, int nranks, const void* uid, salg_ctx** out) {
    return guarded([&] {
        SALG_REQUIRE(out, SALG_ERR_BAD_ARG, "out is NULL");
        SALG_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, SALG_ERR_BAD_ARG, "bad rank/nranks");
        salg_ctx* c = ctx_new(device);
        c->rank = rank;
        c->nranks = nranks;
        if (nranks > 1) {
            SALG_REQUIRE(uid, SALG_ERR_BAD_ARG, "nccl_unique_id is NULL");
            ncclUniqueId id;
            memcpy(&id, uid, sizeof(id));
            ncclResult_t r = ncclCommInitRank(&c->comm, nranks, id, rank);
            if (r != ncclSuccess) {
                std::string m = std::string("ncclCommInitRank failed: ") + ncclGetErrorString(r);
                salg_ctx_destroy(c);
                throw Error(SALG_ERR_NCCL, m);
            }
            try {
                p2p_init(c);            // collective; leaves NCCL as the only path if any rank cannot map its peers
            } catch (...) {
                salg_ctx_destroy(c);
                throw;
            }
        }
        *out = c;
    });
}

int salg_ctx_destroy(salg_ctx* c) {
    return guarded([&] {
        if (!c) return;
        {
            std::lock_guard<std::mutex> lk(g_ctx_mutex);
            g_live_ctx.erase(c);
        }
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        prof_collect(c);
        for (auto e : c->event_pool) cudaEventDestroy(e);
        if (c->ev_fork) cudaEventDestroy(c->ev_fork);
        if (c->ev_join) cudaEventDestroy(c->ev_join);
        if (c->timer0) cudaEventDestroy(c->timer0);
        if (c->timer1) cudaEventDestroy(c->timer1);
        for (int i = 0; i < salg_ctx::N_STAGE; i++) {
            if (c->stage[i]) cudaFreeHost(c->stage[i]);
            if (i == 0 && c->scratch) cudaFree(c->scratch);
            if (i == 0 && c->stream) small_cache_release(c->stream);
            if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
        }
        p2p_destroy(c);
        if (c->comm) ncclCommDestroy(c->comm);
        if (c->stream) cudaStreamDestroy(c->stream);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        delete c;
    });
}

int salg_ctx_sync(salg_ctx* c) {
    return guarded([&] {
        SALG_REQUIRE(c, SALG_ERR_BAD_ARG, "ctx is NULL");
        SALG_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int salg_ctx_rank(const salg_ctx* c, int* rank, int* nranks) {
    return guarded([&] {
        SALG_REQUIRE(c, SALG_ERR_BAD_ARG, "ctx is NULL");
        if (rank) *rank = c->rank;
        if (nranks) *nranks = c->nranks;
    });
}

int salg_prof_enable(salg_ctx* c, int on) {
    return guarded([&] {
        SALG_REQUIRE(c, SALG_ERR_BAD_ARG, "ctx is NULL");
        c->prof_on = on != 0;
        c->prof_products_only = on == 2;     // 2: time only the two sparse-product classes (fewer events in a timed region)
    });
}

int salg_prof_reset(salg_ctx* c) {
    return guarded([&] {
        SALG_REQUIRE(c, SALG_ERR_BAD_ARG, "ctx is NULL");
        prof_collect(c);
        for (int i = 0; i < PROF_NCLS; i++) {
            c->prof_ms[i] = 0;
            c->prof_launches[i] = 0;
            c->prof_bytes[i] = 0;
        }
    });
}

int salg_timer_start(salg_ctx* c) {
    return guarded([&] {
        SALG_REQUIRE(c, SALG_ERR_BAD_ARG, "ctx is NULL");
        SALG_CUDA(cudaSetDevice(c->device));
        if (!c->timer0) {
            SALG_CUDA(cudaEventCreate(&c->timer0));
            SALG_CUDA(cudaEventCreate(&c->timer1));
        }
        SALG_CUDA(cudaStreamSynchronize(c->stream));
        SALG_CUDA(cudaEventRecord(c->timer0, c->stream));
    });
}

int salg_timer_stop(salg_ctx* c, double* ms) {
    return guarded([&] {
        SALG_REQUIRE(c && ms && c->timer0, SALG_ERR_BAD_ARG, "timer not started");
        SALG_CUDA(cudaEventRecord(c->timer1, c->stream));
        SALG_CUDA(cudaEventSynchronize(c->timer1));
        float f = 0.f;
        SALG_CUDA(cudaEventElapsedTime(&f, c->timer0, c->timer1));
        *ms = (double)f;
    });
}

int salg_ctx_set_spmm_impl(salg_ctx* c, int impl) {
    return guarded([&] {
        SALG_REQUIRE(c, SALG_ERR_BAD_ARG, "ctx is NULL");
        SALG_REQUIRE(impl >= 0 && impl <= 2, SALG_ERR_BAD_ARG,
                     "impl must be 0 (tcgen05, dense tile in shared memory), 1 (chunk) or 2 (tcgen05, sparse operand in TMEM)");
        c->spmm_impl = impl;
    });
}

int salg_launch_count(salg_ctx* c, int64_t* out) {
    return guarded([&] {
        SALG_REQUIRE(c && out, SALG_ERR_BAD_ARG, "NULL argument");
        *out = c->n_launch;
    });
}

int salg_prof_count(void) { return PROF_NCLS; }

const char* salg_prof_name(int cls) {
    if (cls < 0 || cls >= PROF_NCLS) return "";
    return kProfNames[cls];
}

int salg_prof_get(salg_ctx* c, int cls, double* ms, int64_t* launches, double* bytes) {
    return guarded([&] {
        SALG_REQUIRE(c, SALG_ERR_BAD_ARG, "ctx is NULL");
        SALG_REQUIRE(cls >= 0 && cls < PROF_NCLS, SALG_ERR_BAD_ARG, "bad profiling class");
        prof_collect(c);
        if (ms) *ms = c->prof_ms[cls];
        if (launches) *launches = c->prof_launches[cls];
        if (bytes) *bytes = c->prof_bytes[cls];
    });
}

}  // extern "C"
