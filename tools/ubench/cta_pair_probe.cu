// cta_pair_probe.cu -- does a CTA pair (cluster of 2, tcgen05 cta_group::2) compute D[256 x N] = [A0; A1] x [B0; B1]^T the way the
// next round's conv kernels want to use it?  Each CTA writes ITS 128 rows of A and ITS N/2 rows of B (K = 64, SWIZZLE_128B
// K-major) into its own shared memory at the same offsets; the peer tells the leader its operands are in place through a
// remote mbarrier arrive; the leader issues 4 x tcgen05.mma.cta_group::2 (M = 256) and a multicast commit; each CTA reads its
// own 128 TMEM lanes.  Host checks both halves exactly (small integers).  All waits are bounded.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cta_pair_probe cta_pair_probe.cu
#include "../../text_segmentation_image_inpainting_b200/csrc/pcb_ptx.cuh"
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ int g_flag = 0;
__device__ __forceinline__ bool spin(uint32_t bar, uint32_t ph, int code) {
    const long long t0 = clock64();
    while (!ptx::mbar_try_wait(bar, ph)) {
        if (clock64() - t0 > 200000000ll) { atomicCAS(&g_flag, 0, code); return false; }
    }
    return true;
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    uint8_t *gen = smem + (base - ptx::smem_u32(smem));
    const uint32_t sA = base, sB = base + 16384, bar_ready = sB + 16384, bar_done = bar_ready + 8, tptr = bar_done + 8;
    const uint32_t rank = cta_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA's operands: A rows [128*rank, +128), B rows [N/2*rank, +N/2)
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int r = i / 8, ch = i % 8;
        *reinterpret_cast<uint4 *>(gen + r * 128 + ((ch ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(A + (128 * rank + r) * 64 + ch * 8);
    }
    for (int i = threadIdx.x; i < (N / 2) * 8; i += blockDim.x) {
        const int r = i / 8, ch = i % 8;
        *reinterpret_cast<uint4 *>(gen + 16384 + r * 128 + ((ch ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(B + ((N / 2) * rank + r) * 64 + ch * 8);
    }
    if (threadIdx.x == 0) { ptx::mbar_init(bar_ready, 1); ptx::mbar_init(bar_done, 1); ptx::fence_mbar_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "n"(N < 32 ? 32 : N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync();                       // both CTAs: operands written, barriers initialised, TMEM allocated
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t *>(gen + (tptr - base));
    bool ok = true;
    if (rank == 1 && threadIdx.x == 0) remote_arrive(mapa(bar_ready, 0));      // "peer operands ready" -> leader's barrier
    if (rank == 0 && threadIdx.x == 0) {
        ok = spin(bar_ready, 0, 1);
        if (ok) {
            ptx::tc_fence_after();
            constexpr uint32_t idesc = ptx::make_idesc_bf16(256, N, 0, 0);
            const uint64_t da = ptx::make_smem_desc(sA, 16, 1024), db = ptx::make_smem_desc(sB, 16, 1024);
            for (int k = 0; k < 4; ++k) {
                const uint32_t acc = k != 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
            }
            const uint16_t mask = 3;      // arrive on bar_done of both CTAs
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_done), "h"(mask) : "memory");
        }
    }
    ok = spin(bar_done, 0, 2 + rank);
    ptx::tc_fence_after();
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(tmem + ((warp * 32u) << 16) + c0, r);
            ptx::tmem_ld_wait();
            for (int j = 0; j < 32; ++j) D[(128 * rank + warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(N < 32 ? 32 : N) : "memory");
}

template <int N>
static void run() {
    std::vector<__nv_bfloat16> a(256 * 64), b(N * 64);
    std::vector<float> af(a.size()), bf(b.size());
    srand(7);
    for (size_t i = 0; i < a.size(); ++i) { af[i] = (float)(rand() % 7 - 3); a[i] = __float2bfloat16(af[i]); }
    for (size_t i = 0; i < b.size(); ++i) { bf[i] = (float)(rand() % 5 - 2); b[i] = __float2bfloat16(bf[i]); }
    __nv_bfloat16 *da, *db; float *dd;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, 256 * N * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, 256 * N * 4);
    cudaFuncSetAttribute(pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    pair_kernel<N><<<2, 128, 40000>>>(da, db, dd);
    cudaError_t e = cudaDeviceSynchronize();
    int flag = 0; cudaMemcpyFromSymbol(&flag, g_flag, sizeof(int));
    printf("N=%3d: launch %s, timeout code %d (1 leader waiting for peer, 2/3 rank waiting for commit)\n", N, cudaGetErrorString(e), flag);
    if (e != cudaSuccess) return;
    std::vector<float> d(256 * N);
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    int bad[2] = {0, 0};
    for (int m = 0; m < 256; ++m) for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += af[m * 64 + k] * bf[n * 64 + k];
        if (ref != d[m * N + n]) ++bad[m / 128];
    }
    printf("       rows 0-127 (leader CTA): %d wrong of %d; rows 128-255 (peer CTA): %d wrong of %d\n", bad[0], 128 * N, bad[1], 128 * N);
    int zero = 0; cudaMemcpyToSymbol(g_flag, &zero, sizeof(int));
    cudaFree(da); cudaFree(db); cudaFree(dd);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    run<64>();
    run<128>();
    printf("done\n");
    return 0;
}
