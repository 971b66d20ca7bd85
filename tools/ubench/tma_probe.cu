// tma_probe.cu -- probes for the TMA-fed implicit-GEMM design (run on one B200):
//   T1  4-D tiled TMA box {64 ch, bw, bh, bn} with negative / out-of-range coordinates (zero fill) -> row order + swizzle
//   T2  the same with elementStrides {1,2,2,1} (stride-2 convolutions)
//   T3  tcgen05.mma with a SWIZZLE_128B K-major A descriptor whose start address is shifted by whole 128-byte rows
//       (with and without the descriptor's base-offset field)
//   T4  throughput of the feeding path: A boxes (distinct per CTA) and B boxes (the same for every CTA), 1 CTA/SM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda -o tma_probe tma_probe.cu
#include "../../text_segmentation_image_inpainting_b200/csrc/pcb_ptx.cuh"
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc_fn() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return reinterpret_cast<EncodeTiledFn>(p);
}
static bool make_map(CUtensorMap *tm, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *es) {
    CUresult r = enc_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("  encode failed: %d\n", (int)r);
    return r == CUDA_SUCCESS;
}
__device__ int g_timeout = 0;
__device__ __forceinline__ void spin(uint32_t bar, uint32_t ph) {       // bounded: a protocol bug must not hang the GPU
    const long long t0 = clock64();
    while (!ptx::mbar_try_wait(bar, ph)) {
        if (clock64() - t0 > 400000000ll || *reinterpret_cast<volatile int *>(&g_timeout)) { g_timeout = 1; return; }
    }
}

// ---------------------------------------------------------------- T1 / T2
__global__ void box_kernel(const __grid_constant__ CUtensorMap tm, int c1, int c2, int c3, int rows, uint16_t *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    uint8_t *gen = smem + (base - ptx::smem_u32(smem));
    const uint32_t bar = base + 32768;
    for (int i = threadIdx.x; i < 32768 / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(gen)[i] = 0xFFFF;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive_expect_tx(bar, rows * 128);
        ptx::tma_load_4d(base, &tm, 0, c1, c2, c3, bar);
    }
    spin(bar, 0);
    for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
        const int r = i / 64, c = i % 64;
        const int chunk = c / 8;
        out[i] = *reinterpret_cast<uint16_t *>(gen + r * 128 + ((chunk ^ (r & 7)) << 4) + (c % 8) * 2);
    }
}

static void test_boxes() {
    const int N = 2, H = 8, W = 16, C = 64;
    std::vector<uint16_t> h(N * H * W * C);
    for (int p = 0; p < N * H * W; ++p) for (int c = 0; c < C; ++c) h[p * C + c] = (uint16_t)(p * 64 + c + 1);
    uint16_t *d, *o;
    cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 128 * 64 * 2);
    cudaFuncSetAttribute(box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int es = 1; es <= 2; ++es) {
        const int bw = es == 1 ? 16 : 8, bh = 4, bn = 2;       // loaded pixels per dim
        cuuint64_t dims[4] = {C, W, H, N};
        cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), (cuuint32_t)bn};
        cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
        CUtensorMap tm;
        printf("T%d: box {64,%d,%d,%d} elementStrides {1,%d,%d,1}\n", es, box[1], box[2], box[3], es, es);
        if (!make_map(&tm, d, 4, dims, strides, box, estr)) continue;
        const int rows = bw * bh * bn;
        const int c1 = -1, c2 = -1, c3 = 0;
        box_kernel<<<1, 128, 40000>>>(tm, c1, c2, c3, rows, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); return; }
        std::vector<uint16_t> r(rows * 64);
        cudaMemcpy(r.data(), o, rows * 128, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int row = 0; row < rows; ++row) {
            const int j = row % bw, i = (row / bw) % bh, n = row / (bw * bh);
            const int w = c1 + j * es, hh = c2 + i * es;
            for (int c = 0; c < 64; ++c) {
                const uint16_t want = (w < 0 || w >= W || hh < 0 || hh >= H) ? 0 : h[((n * H + hh) * W + w) * C + c];
                if (r[row * 64 + c] != want && bad++ < 4) printf("  row %d (n%d h%d w%d) c%d: got %u want %u\n", row, n, hh, w, c, r[row * 64 + c], want);
            }
        }
        printf("  %s (%d mismatches of %d)\n", bad ? "MISMATCH" : "OK: rows ordered (n, h, w), OOB rows zero", bad, rows * 64);
    }
    cudaFree(d); cudaFree(o);
}

// ---------------------------------------------------------------- T3
__global__ void shift_mma_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int shift, int use_base_offset) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    uint8_t *gen = smem + (base - ptx::smem_u32(smem));
    const uint32_t sA = base, sB = base + 144 * 128, bar = sB + 64 * 128, tptr = bar + 8;
    for (int i = threadIdx.x; i < 144 * 8; i += blockDim.x) {            // A: 144 rows x 8 chunks
        const int r = i / 8, ch = i % 8;
        *reinterpret_cast<uint4 *>(gen + r * 128 + ((ch ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(A + r * 64 + ch * 8);
    }
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        const int r = i / 8, ch = i % 8;
        *reinterpret_cast<uint4 *>(gen + 144 * 128 + r * 128 + ((ch ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(B + r * 64 + ch * 8);
    }
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc<64>(tptr); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t *>(gen + (tptr - base));
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 0);
        uint64_t da = ptx::make_smem_desc(sA + shift * 128, 16, 1024);
        if (use_base_offset) da |= static_cast<uint64_t>(shift & 7) << 49;
        const uint64_t db = ptx::make_smem_desc(sB, 16, 1024);
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
        ptx::umma_commit(bar);
    }
    spin(bar, 0);
    ptx::tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem + ((warp * 32u) << 16) + c0, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<64>(tmem);
}

static void test_shift() {
    std::vector<__nv_bfloat16> a(144 * 64), b(64 * 64);
    std::vector<float> af(144 * 64), bf(64 * 64);
    srand(1);
    for (size_t i = 0; i < a.size(); ++i) { af[i] = (float)(rand() % 7 - 3); a[i] = __float2bfloat16(af[i]); }
    for (size_t i = 0; i < b.size(); ++i) { bf[i] = (float)(rand() % 5 - 2); b[i] = __float2bfloat16(bf[i]); }
    __nv_bfloat16 *da, *db; float *dd;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, 128 * 64 * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(shift_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    printf("T3: SWIZZLE_128B K-major A descriptor, start address shifted by whole rows\n");
    for (int ubo = 0; ubo <= 1; ++ubo)
        for (int shift : {0, 1, 2, 3, 5, 8, 9, 12}) {
            shift_mma_kernel<<<1, 128, 40000>>>(da, db, dd, shift, ubo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); return; }
            std::vector<float> d(128 * 64);
            cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
                float ref = 0;
                for (int k = 0; k < 64; ++k) ref += af[(m + shift) * 64 + k] * bf[n * 64 + k];
                if (ref != d[m * 64 + n]) ++bad;
            }
            printf("  base_offset %s shift %2d : %s (%d wrong of 8192)\n", ubo ? "set  " : "zero ", shift, bad ? "WRONG" : "exact", bad);
        }
    cudaFree(da); cudaFree(db); cudaFree(dd);
}

// ---------------------------------------------------------------- T4
// persistent CTA: thread 0 = TMA producer, thread 32 = consumer (release only).  Items walk 3x3 taps over row tiles of a [8,256,256,C] tensor.
__global__ void feed_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB64, int stages, int items, int a_on, int b_bytes, int kcols) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t stage_bytes = 16384 + 32768;
    const uint32_t bars = base + stages * stage_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { ptx::mbar_init(bars + 8 * s, 1); ptx::mbar_init(bars + 8 * (stages + s), 1); }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0; uint32_t ph = 1;
        int tile = blockIdx.x, tap = 0, cb = 0;
        for (int it = 0; it < items; ++it) {
            spin(bars + 8 * (stages + s), ph);
            ptx::mbar_arrive_expect_tx(bars + 8 * s, (a_on ? 16384 : 0) + b_bytes);
            const int w0 = (tile & 1) * 128, h0 = (tile >> 1) & 255, n0 = (tile >> 9) & 7;
            if (a_on) ptx::tma_load_4d(base + s * stage_bytes, &tmA, cb * 64, w0 - 1 + tap % 3, h0 - 1 + tap / 3, n0, bars + 8 * s);
            const int kx = ((tap * 3 + cb) * 64) % kcols;
            if (b_bytes == 8192) ptx::tma_load_2d(base + s * stage_bytes + 16384, &tmB64, kx, 0, bars + 8 * s);
            else for (int b = 0; b < b_bytes; b += 16384) ptx::tma_load_2d(base + s * stage_bytes + 16384 + b, &tmB, kx, b / 128, bars + 8 * s);
            if (++cb == 3) { cb = 0; if (++tap == 9) { tap = 0; tile += gridDim.x; } }
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < items; ++it) {
            spin(bars + 8 * s, ph);
            ptx::mbar_arrive(bars + 8 * (stages + s));
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    }
}

static void test_feed() {
    const int N = 8, H = 256, W = 256, C = 192, KC = 27 * 64;
    __nv_bfloat16 *x, *w;
    cudaMalloc(&x, (size_t)N * H * W * C * 2); cudaMemset(x, 0, (size_t)N * H * W * C * 2);
    cudaMalloc(&w, (size_t)256 * KC * 2); cudaMemset(w, 0, (size_t)256 * KC * 2);
    cuuint64_t dims[4] = {C, W, H, N};
    cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, 128, 1, 1}, es[4] = {1, 1, 1, 1};
    CUtensorMap tmA, tmB;
    make_map(&tmA, x, 4, dims, strides, box, es);
    cuuint64_t d2[2] = {KC, 256}, s2[1] = {KC * 2};
    cuuint32_t b2[2] = {64, 128}, e2[2] = {1, 1};
    make_map(&tmB, w, 2, d2, s2, b2, e2);
    CUtensorMap tmB64;
    cuuint32_t b3[2] = {64, 64};
    make_map(&tmB64, w, 2, d2, s2, b3, e2);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("T4: TMA feed throughput, 1 CTA/SM, %d SMs (B/clk assumes %.2f GHz)\n", sms, clk * 1e-6);
    const int stages = 4, items = 4000;
    const size_t smem = stages * (16384 + 32768) + 1024 + 16 * stages;
    cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct { int a, b; const char *name; } cfg[] = {{1, 0, "A only (16 KB distinct per CTA)"}, {0, 16384, "B only (16 KB, same for all CTAs)"}, {0, 32768, "B only (32 KB, same for all CTAs)"},
                                                    {1, 8192, "A + B 8 KB  (N=64)"}, {1, 16384, "A + B 16 KB (N=128)"}, {1, 32768, "A + B 32 KB (N=256)"}};
    for (auto &c : cfg) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        feed_kernel<<<sms, 64, smem>>>(tmA, tmB, tmB64, stages, 400, c.a, c.b, KC);
        cudaEventRecord(e0);
        feed_kernel<<<sms, 64, smem>>>(tmA, tmB, tmB64, stages, items, c.a, c.b, KC);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t e = cudaGetLastError();
        int to = 0; cudaMemcpyFromSymbol(&to, g_timeout, sizeof(int));
        if (to) { printf("  %s: TIMEOUT in a barrier wait\n", c.name); return; }
        const double bytes = (double)items * ((c.a ? 16384 : 0) + c.b);
        printf("  %-36s %7.3f ms  %6.1f B/clk/SM  %6.0f cyc/item  %5.2f TB/s total %s\n", c.name, ms, bytes / (ms * 1e-3 * clk * 1e3), ms * 1e-3 * clk * 1e3 / items,
               bytes * sms / (ms * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaFree(x); cudaFree(w);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    test_boxes();
    test_shift();
    test_feed();
    printf("done\n");
    return 0;
}
