// mma_issue.cu -- how fast can ONE thread issue tcgen05.mma (cta_group::1, kind::f16, M=128, K=16)?
// Back-to-back UTCHMMA on fixed shared-memory operands, no data dependencies other than the accumulator:
//   variant 0: constant descriptors, constant accumulate flag (the leanest possible stream)
//   variant 1: descriptors advanced by 64-bit adds per MMA, runtime accumulate flag (what the conv kernel does)
// for N = 64 / 128 / 256.  Reports SM cycles per MMA; the tensor pipe needs N/2 cycles (32 / 64 / 128).
#include "../../text_segmentation_image_inpainting_b200/csrc/pcb_ptx.cuh"
#include <stdio.h>

template <int N, int VARIANT>
__global__ void issue_kernel(int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 4 * 16384, bar = sB + 4 * 32768, tptr = bar + 8;
    uint8_t *gen = smem + (base - ptx::smem_u32(smem));
    for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(gen)[i] = 0;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc<256>(tptr); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t *>(gen + (tptr - base));
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
        const uint64_t da0 = ptx::make_smem_desc(sA, 16, 1024), db0 = ptx::make_smem_desc(sB, 16, 1024);
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (VARIANT == 0) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(da0 + 2 * (k & 3)), "l"(db0 + 2 * (k & 3)), "r"(idesc) : "memory");
                }
            } else {
                uint64_t da = da0 + (it & 3) * 1024, db = db0 + (it & 3) * 2048;
                uint32_t accum = it;
                for (int tc = 0; tc < 4; ++tc, da += 8, db += 16) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { ptx::umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, accum != 0); accum = 1; }
                }
            }
            if ((it & 7) == 7) {                      // drain every 128 MMAs so the queue depth stays bounded
                ptx::umma_commit(bar);
                while (!ptx::mbar_try_wait(bar, ph)) {}
                ph ^= 1;
            }
        }
        ptx::umma_commit(bar);
        while (!ptx::mbar_try_wait(bar, ph)) {}
        out[blockIdx.x] = clock64() - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<256>(tmem);
}

template <int N, int V>
void run(long long *d) {
    const int iters = 2048;
    const size_t smem = 4 * 16384 + 4 * 32768 + 1024 + 64;
    cudaFuncSetAttribute(issue_kernel<N, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    issue_kernel<N, V><<<148, 128, smem>>>(64, d);
    issue_kernel<N, V><<<148, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("N=%3d variant %d : %7.1f cycles per MMA (pipe floor %d)  %s\n", N, V, avg / (iters * 16.0), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    long long *d;
    cudaMalloc(&d, 148 * 8);
    run<64, 0>(d); run<128, 0>(d); run<256, 0>(d);
    run<64, 1>(d); run<128, 1>(d); run<256, 1>(d);
    return 0;
}
