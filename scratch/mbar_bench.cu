// microbenchmark: mbarrier ping-pong latency between two warps of one CTA, plus primitive costs
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait_try(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok = 0;
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void mbar_wait_test(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok = 0;
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// mode: 0 try_wait/arrive, 1 test_wait/arrive, 2 try_wait with B side using tcgen05.commit
__global__ void pingpong(int iters, int mode, int extra_warps, long long* out) {
    __shared__ uint64_t a2b, b2a;
    __shared__ uint64_t dummy;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&a2b, 1); mbar_init(&b2a, 1); mbar_init(&dummy, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    long long t0 = clock64();
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < iters; i++) {
            mbar_arrive(&a2b);
            if (mode == 1) mbar_wait_test(&b2a, i & 1); else mbar_wait_try(&b2a, i & 1);
        }
        out[blockIdx.x] = clock64() - t0;
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < iters; i++) {
            if (mode == 1) mbar_wait_test(&a2b, i & 1); else mbar_wait_try(&a2b, i & 1);
            if (mode == 2) commit(&b2a); else mbar_arrive(&b2a);
        }
    } else if (warp >= 2 && warp < 2 + extra_warps) {
        // bystanders spinning on a barrier that never completes until the end (contention test)
        if (lane == 0) { while (true) { uint32_t ok; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&dummy)), "r"(0) : "memory"); if (ok) break; } }
    }
    if (warp == 0 && lane == 0) { __threadfence_block(); mbar_arrive(&dummy); }
}

// cost of primitives executed by all threads of a 512-thread group
__global__ void prims(int iters, long long* out) {
    __shared__ uint4 buf[1024];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) asm volatile("bar.sync 1, 512;" ::: "memory");
    long long t1 = clock64();
    for (int i = 0; i < iters; i++) { buf[threadIdx.x] = make_uint4(i, 0, 0, 0); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    long long t2 = clock64();
    for (int i = 0; i < iters; i++) { buf[threadIdx.x] = make_uint4(i, 0, 0, 0); __syncwarp(); }
    long long t3 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; }
}

int main() {
    long long* d; cudaMalloc(&d, 1024 * 8); long long h[8];
    const int iters = 20000;
    for (int mode = 0; mode < 3; mode++) for (int extra = 0; extra <= 12; extra += 12) {
        pingpong<<<1, 32 * 16, 0>>>(iters, mode, extra, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        printf("pingpong mode %d (0 try_wait, 1 test_wait, 2 try_wait + tcgen05.commit) bystanders %2d: %.1f cycles per round trip (%s)\n", mode, extra, (double)h[0] / iters, cudaGetErrorString(e));
    }
    prims<<<1, 512>>>(iters, d);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("bar.sync(512): %.1f cyc   STS+fence.proxy.async: %.1f cyc   STS+syncwarp: %.1f cyc\n", (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters);
    return 0;
}
