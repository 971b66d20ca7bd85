// dwconv.cu -- depthwise k x k convolution (groups == channels), the HBM-bound half of the depthwise-separable
// blocks: DSConvBlock (models/BaseModels.py:105-127), InvertedResidual / RFB branches (models/MobileNetV2.py:136-138,
// models/common.py:136-143: dilation up to 29), and the depthwise PARTIAL convolutions of PartialInvertedResidual
// (models/MobileNetV2.py:174-176).  NHWC, 8 channels (16 B of bf16 / 32 B of fp32) per thread, fp32 accumulation;
// one read of x and one write of y per element is the roofline (weights are k*k*C, negligible).
// It is an internal fast path of pcb_pconv_forward / backward_* (same semantics: optional hole mask with zero-fill,
// mask-sum renormalisation, `plain` mode for ordinary convolutions).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "pcb_common.cuh"

namespace {

struct DwParams {
    int n, h, w, c, kh, kw, stride, pad_h, pad_w, dil, ho, wo;
    int x_cstride, y_cstride;
    const uint8_t *mask;     // input hole plane [n, h>>mup, w>>mup] or null
    int mup;
    const float *msum;       // [mg][n*ho*wo] or null (plain): renormaliser
    int mg;                  // 1, or c (per-channel sums: groups > 1 && !same_holes)
    int no_guard;
};

__device__ __forceinline__ bool dw_mask(const DwParams &P, int nn, int hi, int wi) {
    if (!P.mask) return true;
    return P.mask[(static_cast<long long>(nn) * (P.h >> P.mup) + (hi >> P.mup)) * (P.w >> P.mup) + (wi >> P.mup)] != 0;
}

// w_t: [taps][c] (T), bias fp32 [c] or null
template <typename T>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const DwParams P, const T *__restrict__ x, const T *__restrict__ w_t,
                                                     const float *__restrict__ bias, T *__restrict__ y) {
    const int cv = P.c >> 3;
    const long long total = static_cast<long long>(P.n) * P.ho * P.wo;
    const long long nvec = total * cv;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = i / cv;
        const int ch = static_cast<int>(i - m * cv) * 8;
        const int ow = static_cast<int>(m % P.wo);
        const long long t = m / P.wo;
        const int oh = static_cast<int>(t % P.ho), nn = static_cast<int>(t / P.ho);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int tr = 0; tr < P.kh; ++tr) {
            const int hi = oh * P.stride - P.pad_h + tr * P.dil;
            if (hi < 0 || hi >= P.h) continue;
            for (int tc = 0; tc < P.kw; ++tc) {
                const int wi = ow * P.stride - P.pad_w + tc * P.dil;
                if (wi < 0 || wi >= P.w || !dw_mask(P, nn, hi, wi)) continue;
                float xv[8], wv[8];
                Vec8<T>::load(x + (static_cast<long long>(nn * P.h + hi) * P.w + wi) * P.x_cstride + ch, xv);
                Vec8<T>::load(w_t + static_cast<long long>(tr * P.kw + tc) * P.c + ch, wv);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += xv[j] * wv[j];
            }
        }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float b = bias ? bias[ch + j] : 0.f;
            if (!P.msum) { o[j] = acc[j] + b; continue; }
            const float s = P.msum[(P.mg == 1 ? 0 : static_cast<long long>(ch + j) * total) + m];
            if (P.no_guard) o[j] = acc[j] / s + b;
            else o[j] = (s == 0.f) ? 0.f : acc[j] / s + b;
        }
        Vec8<T>::store(y + m * P.y_cstride + ch, o);
    }
}

// dx[p][c] = mask(p) * sum_taps dc[(p + pad - tap*dil)/stride][c] * w[tap][c]
template <typename T>
__global__ void __launch_bounds__(256) dw_dgrad_kernel(const DwParams P, const T *__restrict__ dc, int dc_cstride, const T *__restrict__ w_t,
                                                       T *__restrict__ dx, int dx_cstride) {
    const int cv = P.c >> 3;
    const long long total = static_cast<long long>(P.n) * P.h * P.w;
    const long long nvec = total * cv;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = i / cv;
        const int ch = static_cast<int>(i - m * cv) * 8;
        const int iw = static_cast<int>(m % P.w);
        const long long t = m / P.w;
        const int ih = static_cast<int>(t % P.h), nn = static_cast<int>(t / P.h);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        if (dw_mask(P, nn, ih, iw)) {
            for (int tr = 0; tr < P.kh; ++tr) {
                const int th = ih + P.pad_h - tr * P.dil;
                if (th < 0) continue;
                const int oh = th / P.stride;
                if (oh * P.stride != th || oh >= P.ho) continue;
                for (int tc = 0; tc < P.kw; ++tc) {
                    const int tw = iw + P.pad_w - tc * P.dil;
                    if (tw < 0) continue;
                    const int ow = tw / P.stride;
                    if (ow * P.stride != tw || ow >= P.wo) continue;
                    float dv[8], wv[8];
                    Vec8<T>::load(dc + (static_cast<long long>(nn * P.ho + oh) * P.wo + ow) * dc_cstride + ch, dv);
                    Vec8<T>::load(w_t + static_cast<long long>(tr * P.kw + tc) * P.c + ch, wv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += dv[j] * wv[j];
                }
            }
        }
        Vec8<T>::store(dx + m * dx_cstride + ch, acc);
    }
}

// dw[c][tap] += sum_pixels dc[p][c] * (x*m)[p@tap][c];  TP taps per pass kept in registers
template <typename T, int TP>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const DwParams P, const T *__restrict__ dc, int dc_cstride, const T *__restrict__ x,
                                                       float *__restrict__ dw, int tap0) {
    __shared__ float s_red[256][8];
    const int cv = P.c >> 3, rpb = 256 / cv;
    const int r = threadIdx.x / cv, v = threadIdx.x - r * cv;
    const int taps = P.kh * P.kw;
    const long long total = static_cast<long long>(P.n) * P.ho * P.wo;
    float acc[TP][8];
#pragma unroll
    for (int tp = 0; tp < TP; ++tp)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[tp][j] = 0.f;
    if (r < rpb) {
        for (long long m = static_cast<long long>(blockIdx.x) * rpb + r; m < total; m += static_cast<long long>(gridDim.x) * rpb) {
            const int ow = static_cast<int>(m % P.wo);
            const long long t = m / P.wo;
            const int oh = static_cast<int>(t % P.ho), nn = static_cast<int>(t / P.ho);
            float dv[8];
            Vec8<T>::load(dc + m * dc_cstride + v * 8, dv);
#pragma unroll
            for (int tp = 0; tp < TP; ++tp) {
                const int tap = tap0 + tp;
                if (tap >= taps) continue;
                const int tr = tap / P.kw, tc = tap - tr * P.kw;
                const int hi = oh * P.stride - P.pad_h + tr * P.dil, wi = ow * P.stride - P.pad_w + tc * P.dil;
                if (hi < 0 || hi >= P.h || wi < 0 || wi >= P.w || !dw_mask(P, nn, hi, wi)) continue;
                float xv[8];
                Vec8<T>::load(x + (static_cast<long long>(nn * P.h + hi) * P.w + wi) * P.x_cstride + v * 8, xv);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[tp][j] += dv[j] * xv[j];
            }
        }
    }
#pragma unroll
    for (int tp = 0; tp < TP; ++tp) {
        const int tap = tap0 + tp;
        if (tap >= taps) continue;          // uniform across the block
        __syncthreads();
        if (r < rpb) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_red[threadIdx.x][j] = acc[tp][j];
        }
        __syncthreads();
        if (r == 0 && v < cv) {
            float tot[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tot[j] = 0.f;
            for (int rr = 0; rr < rpb; ++rr)
#pragma unroll
                for (int j = 0; j < 8; ++j) tot[j] += s_red[rr * cv + v][j];
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(dw + static_cast<long long>(v * 8 + j) * taps + tap, tot[j]);
        }
    }
}

template <typename T>
__global__ void dw_weight_transpose_kernel(const float *__restrict__ wm, int c, int taps, T *__restrict__ w_t) {
    const long long total = static_cast<long long>(c) * taps;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i % taps), ch = static_cast<int>(i / taps);
        w_t[static_cast<long long>(tap) * c + ch] = from_f32<T>(wm[i]);
    }
}

void fill(DwParams &P, const pcb_conv *c) {
    memset(&P, 0, sizeof(P));
    P.n = c->n; P.h = c->h; P.w = c->w; P.c = c->cin; P.kh = c->kh; P.kw = c->kw; P.stride = c->stride; P.pad_h = c->pad_h;
    P.pad_w = c->pad_w; P.dil = c->dil; P.ho = c->ho; P.wo = c->wo; P.x_cstride = c->parts[0].x_cstride;
    P.mask = c->parts[0].mask; P.mup = c->parts[0].mask_up; P.no_guard = c->no_guard;
    P.mg = (c->groups > 1 && !c->same_holes) ? c->groups : 1;
}

inline int dw_grid(long long items) {
    long long b = (items + 255) / 256;
    const long long cap = 32ll * pcb_num_sms();
    return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

bool pcb_dw_eligible(const pcb_conv *c) {
    if (c->force_generic || getenv("PCB_DISABLE_DW")) return false;
    if (!(c->groups == c->cin && c->cin == c->cout && c->groups > 1 && c->nparts == 1)) return false;
    const pcb_part &pt = c->parts[0];
    if (c->cin % 8 != 0 || c->cin > 2048 || pt.x_cstride % 8 != 0 || pt.x_up != 0) return false;
    const uintptr_t align = (c->dtype == PCB_BF16) ? 15 : 31;
    if (pt.x && (reinterpret_cast<uintptr_t>(pt.x) & align)) return false;
    return true;
}

int pcb_dw_weight_prepare(const pcb_conv *c, const float *w_master, void *w_t, cudaStream_t st) {
    const int taps = c->kh * c->kw;
    const int grid = dw_grid(static_cast<long long>(c->cin) * taps);
    if (c->dtype == PCB_BF16) dw_weight_transpose_kernel<bf16><<<grid, 256, 0, st>>>(w_master, c->cin, taps, static_cast<bf16 *>(w_t));
    else dw_weight_transpose_kernel<float><<<grid, 256, 0, st>>>(w_master, c->cin, taps, static_cast<float *>(w_t));
    PCB_LAUNCH_CHECK();
    return 0;
}

int pcb_dw_forward(const pcb_conv *c, const void *w_t, const float *bias, void *y, int y_cstride, const float *msum, cudaStream_t st) {
    DwParams P;
    fill(P, c);
    P.y_cstride = y_cstride;
    P.msum = msum;       // plain mode: mask_sums wrote 1.0 everywhere, so the same epilogue applies
    const long long nvec = static_cast<long long>(c->n) * c->ho * c->wo * (c->cin / 8);
    if (c->dtype == PCB_BF16) dw_fwd_kernel<bf16><<<dw_grid(nvec), 256, 0, st>>>(P, static_cast<const bf16 *>(c->parts[0].x), static_cast<const bf16 *>(w_t), bias, static_cast<bf16 *>(y));
    else dw_fwd_kernel<float><<<dw_grid(nvec), 256, 0, st>>>(P, static_cast<const float *>(c->parts[0].x), static_cast<const float *>(w_t), bias, static_cast<float *>(y));
    PCB_LAUNCH_CHECK();
    return 0;
}

int pcb_dw_dgrad(const pcb_conv *c, const void *dc, int dc_cstride, const void *w_t, void *dx, int dx_cstride, cudaStream_t st) {
    DwParams P;
    fill(P, c);
    const long long nvec = static_cast<long long>(c->n) * c->h * c->w * (c->cin / 8);
    if (c->dtype == PCB_BF16) dw_dgrad_kernel<bf16><<<dw_grid(nvec), 256, 0, st>>>(P, static_cast<const bf16 *>(dc), dc_cstride, static_cast<const bf16 *>(w_t), static_cast<bf16 *>(dx), dx_cstride);
    else dw_dgrad_kernel<float><<<dw_grid(nvec), 256, 0, st>>>(P, static_cast<const float *>(dc), dc_cstride, static_cast<const float *>(w_t), static_cast<float *>(dx), dx_cstride);
    PCB_LAUNCH_CHECK();
    return 0;
}

int pcb_dw_wgrad(const pcb_conv *c, const void *dc, int dc_cstride, float *dw, bool zero_dw, cudaStream_t st) {
    DwParams P;
    fill(P, c);
    const int taps = c->kh * c->kw;
    if (zero_dw) PCB_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * c->cin * taps, st));
    const long long total = static_cast<long long>(c->n) * c->ho * c->wo;
    const int rpb = 256 / (c->cin / 8);
    long long blocks = (total + rpb * 16 - 1) / (rpb * 16);
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(blocks, 8ll * pcb_num_sms())));
    constexpr int TP = 9;
    for (int tap0 = 0; tap0 < taps; tap0 += TP) {
        if (c->dtype == PCB_BF16) dw_wgrad_kernel<bf16, TP><<<grid, 256, 0, st>>>(P, static_cast<const bf16 *>(dc), dc_cstride, static_cast<const bf16 *>(c->parts[0].x), dw, tap0);
        else dw_wgrad_kernel<float, TP><<<grid, 256, 0, st>>>(P, static_cast<const float *>(dc), dc_cstride, static_cast<const float *>(c->parts[0].x), dw, tap0);
        PCB_LAUNCH_CHECK();
    }
    return 0;
}
