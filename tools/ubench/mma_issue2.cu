// mma_issue2.cu -- what slows the MMA-issuing thread down inside a pipelined kernel?  N=128, items of 12 MMAs (3 taps x 4 K steps).
//   flags bit0: per item  try_wait on a ready barrier + tcgen05.fence::after + tcgen05.commit to a ring barrier
//         bit1: a warp on the same scheduler (warp 5) polls an mbarrier that never completes (fixer / epilogue style wait)
//         bit2: four more warps (2,3,4,6) poll as well
//         bit3: a TMA thread streams 64 KB per item into a separate smem ring concurrently (bulk copies from global)
#include "../../text_segmentation_image_inpainting_b200/csrc/pcb_ptx.cuh"
#include <stdio.h>

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(320, 1) k(int items, int flags, const uint8_t *gsrc, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 32768, sT = base + 65536;            // sT: 2 x 64 KB TMA ring
    const uint32_t bars = sT + 2 * 65536, tptr = bars + 128;
    uint8_t *gen = smem + (base - ptx::smem_u32(smem));
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(gen)[i] = 0;
    // bars: [0] ready (completed once), [1] never, [2..5] commit ring, [6,7] tma full, [8,9] tma empty, [10] final
    if (threadIdx.x == 0) { for (int i = 0; i < 12; ++i) ptx::mbar_init(bars + 8 * i, 1); ptx::fence_mbar_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc<256>(tptr); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t *>(gen + (tptr - base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ volatile int done;
    if (threadIdx.x == 0) { done = 0; ptx::mbar_arrive(bars); }
    __syncthreads();
    if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 128, 0, 0);
            const uint64_t da0 = ptx::make_smem_desc(sA, 16, 1024), db0 = ptx::make_smem_desc(sB, 16, 1024);
            int s = 0;
            const long long t0 = clock64();
            for (int it = 0; it < items; ++it) {
                if (flags & 1) { while (!ptx::mbar_try_wait(bars, 0)) {} ptx::tc_fence_after(); }
                uint64_t da = da0 + (it & 1) * 1024, db = db0;
                uint32_t accum = it;
                for (int tc = 0; tc < 3; ++tc, da += 8, db += 16) {
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) { ptx::umma_bf16(tmem, da + 2 * k2, db + 2 * k2, idesc, accum != 0); accum = 1; }
                }
                if (flags & 1) { ptx::umma_commit(bars + 8 * (2 + s)); s = (s + 1) & 3; }
            }
            ptx::umma_commit(bars + 8 * 10);
            while (!ptx::mbar_try_wait(bars + 8 * 10, 0)) {}
            out[blockIdx.x] = clock64() - t0;
            done = 1;
        }
    } else if (warp == 0) {
        if ((flags & 8) && lane == 0) {
            int s = 0; uint32_t ph = 0;
            const uint8_t *src = gsrc + (size_t)blockIdx.x * (1 << 20);
            for (int it = 0; !done && it < items; ++it) {
                ptx::mbar_arrive_expect_tx(bars + 8 * (6 + s), 65536);
                for (int c = 0; c < 4; ++c) bulk_g2s(sT + s * 65536 + c * 16384, src + ((it * 4 + c) & 63) * 16384, 16384, bars + 8 * (6 + s));
                while (!ptx::mbar_try_wait(bars + 8 * (6 + s), ph)) {}
                if (s == 1) ph ^= 1;
                s ^= 1;
            }
        }
    } else if ((warp == 5 && (flags & 2)) || ((warp == 2 || warp == 3 || warp == 4 || warp == 6) && (flags & 4))) {
        while (!done) { ptx::mbar_try_wait(bars + 8, 0); }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<256>(tmem);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    long long *d; uint8_t *g;
    cudaMalloc(&d, 148 * 8); cudaMalloc(&g, 148ull << 20); cudaMemset(g, 0, 148ull << 20);
    const size_t smem = 65536 + 2 * 65536 + 1024 + 256;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int items = 4096;
    for (int flags : {0, 1, 3, 7, 9, 15}) {
        k<<<148, 320, smem>>>(256, flags, g, d);
        k<<<148, 320, smem>>>(items, flags, g, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("flags=%2d (%s%s%s%s) : %7.1f cycles per item of 12 MMAs = %6.1f per MMA (pipe floor 64)  %s\n", flags, flags & 1 ? "wait+fence+commit " : "", flags & 2 ? "poller@same-SMSP " : "",
               flags & 4 ? "4 more pollers " : "", flags & 8 ? "TMA 64KB/item" : "", avg / items, avg / items / 12, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
