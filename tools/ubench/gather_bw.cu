// gather_bw.cu -- microbenchmark of the A-operand feeding path: 16-byte cp.async (LDGSTS) gathers of 128-byte
// pixel rows into a shared-memory ring with mbarrier completion, consumer = one thread that waits for the stage
// and releases it immediately (no MMA).  Sweeps producer warps, ring depth, CTAs/SM and row pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu ; run on one B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t b, uint32_t ph) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) { while (!mbar_try(b, ph)) {} }
__device__ __forceinline__ void cp16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_arrive(uint32_t b) { asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(b) : "memory"); }

// items of `rows` pixel rows x 128 B.  mode 0: consecutive pixels (pixel stride `pix_bytes`), mode 1: each pixel read
// twice in a row (nearest-upsampled source), mode 2: pseudo-random rows.
template <int PW>   // producer warps
__global__ void gather_kernel(const uint8_t *__restrict__ src, size_t src_pixels, int pix_bytes, int rows, int stages, int items_per_cta, int mode,
                              unsigned long long *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sA = smem_u32(smem);
    const uint32_t stage_bytes = rows * 128;
    const uint32_t bars = sA + stages * stage_bytes;      // full[stages], empty[stages]
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(bars + 8 * s, PW * 32); mbar_init(bars + 8 * (stages + s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t base_pixel = (size_t)blockIdx.x * 7919 * rows;
    if (warp < PW) {
        const int chunk = tid & 7, r0 = tid >> 3;          // PW*4 rows per pass
        for (int it = 0; it < items_per_cta; ++it) {
            const int s = it % stages;
            mbar_wait(bars + 8 * (stages + s), ((it / stages) & 1) ^ 1);
            for (int r = r0; r < rows; r += PW * 4) {
                uint32_t pix;                                    // src_pixels is a power of two: mask, no division
                if (mode == 0) pix = (uint32_t)base_pixel + (uint32_t)(it * rows + r);
                else if (mode == 1) pix = (uint32_t)base_pixel + (uint32_t)((it * rows + r) >> 1);
                else pix = ((uint32_t)base_pixel + (uint32_t)(it * rows + r)) * 2654435761u;
                pix &= (uint32_t)(src_pixels - 1);
                cp16(sA + s * stage_bytes + r * 128 + ((chunk ^ (r & 7)) << 4), src + (size_t)pix * pix_bytes + chunk * 16);
            }
            cp_arrive(bars + 8 * s);
            mbar_arrive(bars + 8 * s);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp == PW && (tid & 31) == 0) {
        unsigned long long acc = 0;
        for (int it = 0; it < items_per_cta; ++it) {
            const int s = it % stages;
            mbar_wait(bars + 8 * s, (it / stages) & 1);
            acc += smem[s * stage_bytes + (it & 127)];
            mbar_arrive(bars + 8 * (stages + s));
        }
        if (acc == 0xdeadbeefull) *sink = acc;
    }
}

template <int PW>
void run(const uint8_t *src, size_t pixels, int pix_bytes, int rows, int stages, int ctas_per_sm, int mode, unsigned long long *sink) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int items = 2000;
    const size_t smem = (size_t)stages * rows * 128 + 16 * stages + 64;
    cudaFuncSetAttribute(gather_kernel<PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<PW><<<grid, (PW + 1) * 32, smem>>>(src, pixels, pix_bytes, rows, stages, 200, mode, sink);
    cudaEventRecord(e0);
    gather_kernel<PW><<<grid, (PW + 1) * 32, smem>>>(src, pixels, pix_bytes, rows, stages, items, mode, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    const double bytes = (double)grid * items * rows * 128;
    const double cyc_per_item = ms * 1e-3 * 1.9e9 / items;
    printf("pw=%d rows=%3d stages=%d cta/sm=%d mode=%d pix=%4dB : %7.3f ms  %6.2f TB/s  %5.1f B/clk/SM  %6.0f cyc/item/CTA  in-flight %3zu KB/SM  %s\n", PW, rows, stages,
           ctas_per_sm, mode, pix_bytes, ms, bytes / (ms * 1e-3) / 1e12, bytes / (ms * 1e-3) / 1.9e9 / sms, cyc_per_item,
           (size_t)stages * rows * 128 * ctas_per_sm / 1024, err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    const size_t pixels = 8ull * 256 * 256;         // [8,256,256,C]
    const int pix_bytes_max = 512;
    uint8_t *src; unsigned long long *sink;
    cudaMalloc(&src, pixels * pix_bytes_max); cudaMemset(src, 1, pixels * pix_bytes_max);
    cudaMalloc(&sink, 8);
    printf("# footprint %zu MB at 128 B/pixel (L2-resident), %zu MB at 512 B/pixel\n", pixels * 128 >> 20, pixels * 512 >> 20);
    for (int mode = 0; mode < 3; ++mode)
        for (int pixb : {128, 512}) {
            run<4>(src, pixels, pixb, 128, 3, 1, mode, sink);
            run<4>(src, pixels, pixb, 128, 3, 2, mode, sink);
            run<4>(src, pixels, pixb, 128, 6, 1, mode, sink);
            run<4>(src, pixels, pixb, 128, 6, 2, mode, sink);
            run<8>(src, pixels, pixb, 128, 6, 1, mode, sink);
            run<8>(src, pixels, pixb, 128, 12, 1, mode, sink);
            run<16>(src, pixels, pixb, 128, 12, 1, mode, sink);
        }
    run<4>(src, pixels, 128, 160, 3, 1, 0, sink);     // halo-sized items
    run<4>(src, pixels, 128, 160, 3, 2, 0, sink);
    run<8>(src, pixels, 128, 160, 6, 1, 0, sink);
    printf("done\n");
    return 0;
}
