// tcgen05.cuh — PTX wrappers shared by the tensor-core products (tc.cu: tile-densified operand in shared memory,
// tm.cu: sparse operand expanded into TMEM): mbarriers, bulk copies, tcgen05 alloc / mma / commit / ld, UMMA descriptors.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace salg {

#ifndef TC_SPIN_LIMIT_
#define TC_SPIN_LIMIT_ (1u << 24)
#endif
constexpr uint32_t TC_SPIN_LIMIT = TC_SPIN_LIMIT_;     // polls before a waiting thread traps (each poll suspends up to TC_WAIT_HINT_NS)
// suspend-time hint of mbarrier.try_wait: a waiting thread sleeps in hardware until the phase completes (or this many
// ns pass) instead of re-issuing the poll; with the default hint the ~25 waiting lanes of a CTA were measured to take
// most of the issue slots of the SM (ncu: 62 % issue utilisation, three quarters of it poll loops)
#ifndef TC_WAIT_HINT_NS_
#define TC_WAIT_HINT_NS_ 200000u
#endif
constexpr uint32_t TC_WAIT_HINT_NS = TC_WAIT_HINT_NS_;
#ifndef TC_FAST_WAITS_
#define TC_FAST_WAITS_ 0
#endif

// ---- PTX helpers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(TC_WAIT_HINT_NS)
            : "memory");
        if (!ok && ++spins > TC_SPIN_LIMIT) __trap();   // never hang the GPU on a protocol bug
    } while (!ok);
}
// the same wait without the suspend hint, for the two hand-offs that sit on the critical chain of a unit (scatter group
// <- MMAs retired, MMA issuer <- operand built): waking from a long suspend was measured against polling
#ifndef TC_FAST_WAITS_
#define TC_FAST_WAITS_ 0
#endif
__device__ __forceinline__ void mbar_wait_crit(uint64_t* bar, uint32_t parity) {
    if (!TC_FAST_WAITS_) { mbar_wait(bar, parity); return; }
    uint32_t addr = smem_u32(bar), ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && ++spins > TC_SPIN_LIMIT) __trap();
    } while (!ok);
}
// one lane polls, the warp follows
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same operations on raw shared-memory addresses (the serial roles of tm.cu compute them once, outside their loops)
#ifndef TM_WAIT_HINT_
#define TM_WAIT_HINT_ 0
#endif
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        if (TM_WAIT_HINT_) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(addr), "r"(parity), "r"(TC_WAIT_HINT_NS)
                : "memory");
        } else {
            // no suspend-time hint: the hardware-default suspension of try_wait (whole warps wait here, one poll per warp)
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(addr), "r"(parity)
                : "memory");
        }
        if (!ok && ++spins > (TM_WAIT_HINT_ ? TC_SPIN_LIMIT : (1u << 21))) __trap();
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gsrc),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Four consecutive K-steps (K = 16 fp16 each) of one product in ONE asm block: the issuing thread is the serial
// resource of the CTA, so per-MMA overhead is two 64-bit adds.  Descriptor start addresses advance by a_step /
// b_step (16 B units) per K-step; only the first MMA may overwrite the accumulator.
__device__ __forceinline__ void umma_f16_run4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate_first, uint64_t a_step, uint64_t b_step) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "mov.b64 da, %1;\n\tmov.b64 db, %2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "add.u64 da, da, %5;\n\tadd.u64 db, db, %6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
        "add.u64 da, da, %5;\n\tadd.u64 db, db, %6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
        "add.u64 da, da, %5;\n\tadd.u64 db, db, %6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t"
        "}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate_first), "l"(a_step), "l"(b_step)
        : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B (128 B contiguous);
// LBO = byte distance between the two 16 B K-chunks of one instruction, SBO = distance between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D f32 (bit 4), A/B fp16 (format 0), both K-major, M = 128, N as given
__host__ __device__ constexpr uint32_t tc_idesc(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp16 term `t` (0, 1) of a scaled float: x ~ h0 + h1 with 11 significant bits each
__device__ __forceinline__ unsigned short f16_term(float x, int t) {
    __half h = __float2half_rn(x);
    if (t > 0) h = __float2half_rn(x - __half2float(h));
    return __half_as_ushort(h);
}

// byte offset of element (mn, k) inside a canonical K-major no-swizzle operand whose 16 B K-chunks are
// `chunk_stride` bytes apart (8-row groups are 128 B apart)
__device__ __forceinline__ uint32_t canon_off(uint32_t mn, uint32_t k, uint32_t chunk_stride) {
    return (k >> 3) * chunk_stride + (mn >> 3) * 128u + (mn & 7u) * 16u + (k & 7u) * 2u;
}

// power-of-two scale that puts `amax` just below 2^14 (fp16 max is 65504; the second term is 2^-11 smaller)
__host__ __device__ inline float tc_pow2_scale(float amax) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int e;
    frexpf(amax, &e);                 // amax = f * 2^e, f in [0.5, 1)
    return ldexpf(1.f, 14 - e);
}

}  // namespace salg
