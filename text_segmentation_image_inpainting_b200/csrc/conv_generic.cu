// conv_generic.cu -- shape-general partial convolution kernels (any kernel size / stride / dilation /
// groups / channel count, fp32 or bf16 storage, fp32 accumulation).  They carry the layers the tensor-core
// path does not take (3-channel stem, 3-channel tail, depthwise, per-channel masks) and the exact-fp32 mode.
// Semantics: models/partial_convolution.py:49-80 (PartialConv), :121-137 (PartialConvNoHoles).
#include <algorithm>
#include <string.h>

#include "pcb_common.cuh"

namespace {

struct GPart {
    const void *x;
    const uint8_t *mask;
    int c, choff, cstride, xup, mup;
};

struct GParams {
    int n, h, w, cin, cout, kh, kw, stride, pad_h, pad_w, dil, groups, ho, wo;
    int same_holes, no_guard, plain, nparts;
    GPart parts[PCB_MAX_PARTS];
};

void fill(GParams &G, const pcb_conv *c) {
    G.n = c->n; G.h = c->h; G.w = c->w; G.cin = c->cin; G.cout = c->cout; G.kh = c->kh; G.kw = c->kw;
    G.stride = c->stride; G.pad_h = c->pad_h; G.pad_w = c->pad_w; G.dil = c->dil; G.groups = c->groups;
    G.ho = c->ho; G.wo = c->wo; G.same_holes = c->same_holes; G.no_guard = c->no_guard; G.plain = c->plain; G.nparts = c->nparts;
    int off = 0;
    for (int p = 0; p < c->nparts; ++p) {
        G.parts[p].x = c->parts[p].x; G.parts[p].mask = c->parts[p].mask; G.parts[p].c = c->parts[p].c;
        G.parts[p].choff = off; G.parts[p].cstride = c->parts[p].x_cstride; G.parts[p].xup = c->parts[p].x_up;
        G.parts[p].mup = c->parts[p].mask_up;
        off += c->parts[p].c;
    }
}

__device__ __forceinline__ bool mask_at(const GParams &G, int p, int nn, int hi, int wi) {
    const GPart &pt = G.parts[p];
    if (!pt.mask) return true;
    return pt.mask[(static_cast<long long>(nn) * (G.h >> pt.mup) + (hi >> pt.mup)) * (G.w >> pt.mup) + (wi >> pt.mup)] != 0;
}

// ------------------------------------------------------------------------------------------------
// mask box sums: s = sum over the receptive field (and the group's channels) of the mask; zero padding
// counts as hole.  The reference computes this as a dense all-ones convolution (partial_convolution.py:
// 41-47,59,63): here it is an integer box sum over uint8 planes.
// ------------------------------------------------------------------------------------------------
__global__ void mask_sums_kernel(const GParams G, float *msum, uint8_t *newmask) {
    const long long plane = static_cast<long long>(G.ho) * G.wo;
    const long long total = plane * G.n;
    const long long m = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (m >= total) return;
    const int mg = (G.groups > 1 && !G.same_holes) ? G.groups : 1;
    const int nn = static_cast<int>(m / plane);
    const int rem = static_cast<int>(m - nn * plane), oh = rem / G.wo, ow = rem - oh * G.wo;
    const int cig = G.cin / G.groups;
    for (int g = 0; g < mg; ++g) {
        if (G.plain) { msum[g * total + m] = 1.f; newmask[g * total + m] = 1; continue; }
        int s = 0;
        const int np = G.same_holes ? 1 : G.nparts;
        for (int p = 0; p < np; ++p) {
            int weight;
            if (G.same_holes) weight = G.cin;                                  // :61  count * in_channels
            else if (mg == 1) weight = G.parts[p].c;
            else {                                                             // channels of part p inside group g
                const int lo = max(G.parts[p].choff, g * cig), hi = min(G.parts[p].choff + G.parts[p].c, (g + 1) * cig);
                weight = max(0, hi - lo);
            }
            if (weight == 0) continue;
            int cnt = 0;
            for (int tr = 0; tr < G.kh; ++tr)
                for (int tc = 0; tc < G.kw; ++tc) {
                    const int hi = oh * G.stride - G.pad_h + tr * G.dil, wi = ow * G.stride - G.pad_w + tc * G.dil;
                    if (hi < 0 || hi >= G.h || wi < 0 || wi >= G.w) continue;
                    cnt += mask_at(G, p, nn, hi, wi) ? 1 : 0;
                }
            s += cnt * weight;
        }
        msum[g * total + m] = static_cast<float>(s);
        newmask[g * total + m] = (G.no_guard || s != 0) ? 1 : 0;               // :74-75 / :135
    }
}

// ------------------------------------------------------------------------------------------------
// forward: one thread = one output pixel x CO consecutive output channels
// ------------------------------------------------------------------------------------------------
template <typename T, int CO>
__global__ void generic_fwd_kernel(const GParams G, const T *__restrict__ w, const float *__restrict__ bias,
                                   const float *__restrict__ msum, T *__restrict__ y, int y_cstride) {
    const long long plane = static_cast<long long>(G.ho) * G.wo;
    const long long total = plane * G.n;
    const long long m = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (m >= total) return;
    const int co0 = blockIdx.y * CO;
    const int nn = static_cast<int>(m / plane);
    const int rem = static_cast<int>(m - nn * plane), oh = rem / G.wo, ow = rem - oh * G.wo;
    const int cig = G.cin / G.groups, cog = G.cout / G.groups;
    const int taps = G.kh * G.kw;
    const int mg = (G.groups > 1 && !G.same_holes) ? G.groups : 1;
    float acc[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[j] = 0.f;

    for (int tr = 0; tr < G.kh; ++tr)
        for (int tc = 0; tc < G.kw; ++tc) {
            const int hi = oh * G.stride - G.pad_h + tr * G.dil, wi = ow * G.stride - G.pad_w + tc * G.dil;
            if (hi < 0 || hi >= G.h || wi < 0 || wi >= G.w) continue;
            const int tap = tr * G.kw + tc;
            for (int p = 0; p < G.nparts; ++p) {
                if (!mask_at(G, p, nn, hi, wi)) continue;                      // x * mask, :51
                const GPart &pt = G.parts[p];
                const T *xp = static_cast<const T *>(pt.x) +
                              (static_cast<long long>(nn * (G.h >> pt.xup) + (hi >> pt.xup)) * (G.w >> pt.xup) + (wi >> pt.xup)) * pt.cstride;
                if (G.groups == 1) {
                    for (int cl = 0; cl < pt.c; ++cl) {
                        const float xv = to_f32(xp[cl]);
                        const int ci = pt.choff + cl;
#pragma unroll
                        for (int j = 0; j < CO; ++j)
                            if (co0 + j < G.cout) acc[j] += xv * to_f32(w[(static_cast<long long>(co0 + j) * taps + tap) * cig + ci]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CO; ++j) {
                        const int co = co0 + j;
                        if (co >= G.cout) continue;
                        const int g = co / cog;
                        const int lo = max(pt.choff, g * cig), hi2 = min(pt.choff + pt.c, (g + 1) * cig);
                        for (int ci = lo; ci < hi2; ++ci)
                            acc[j] += to_f32(xp[ci - pt.choff]) * to_f32(w[(static_cast<long long>(co) * taps + tap) * cig + (ci - g * cig)]);
                    }
                }
            }
        }
#pragma unroll
    for (int j = 0; j < CO; ++j) {
        const int co = co0 + j;
        if (co >= G.cout) {
            if (co < y_cstride) y[m * y_cstride + co] = from_f32<T>(0.f);      // channel padding stays finite (zero)
            continue;
        }
        const int g = (mg == 1) ? 0 : co / cog;
        const float s = msum ? msum[g * total + m] : 1.f;               // null: plain convolution (renormaliser 1)
        const float b = bias ? bias[co] : 0.f;
        float v;
        if (G.no_guard) v = acc[j] / s + b;                                    // :134 (NaN/inf on s == 0, as the reference)
        else v = (s == 0.f) ? 0.f : acc[j] / s + b;                            // :71-72
        y[m * y_cstride + co] = from_f32<T>(v);
    }
}

// ------------------------------------------------------------------------------------------------
// dgrad: one thread = one input pixel x CI consecutive input channels;  dx = convT(dc, W) * mask
// w is the forward KRSC weight [cout][kh][kw][cig]
// ------------------------------------------------------------------------------------------------
struct DxOut { void *ptr[PCB_MAX_PARTS]; int cstride[PCB_MAX_PARTS]; };

template <typename T, int CI>
__global__ void generic_dgrad_kernel(const GParams G, const T *__restrict__ dc, int dc_cstride, const T *__restrict__ w, const DxOut O) {
    const long long plane = static_cast<long long>(G.h) * G.w;
    const long long total = plane * G.n;
    const long long m = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (m >= total) return;
    const int ci0 = blockIdx.y * CI;
    const int nn = static_cast<int>(m / plane);
    const int rem = static_cast<int>(m - nn * plane), ih = rem / G.w, iw = rem - ih * G.w;
    const int cig = G.cin / G.groups, cog = G.cout / G.groups;
    const int taps = G.kh * G.kw;
    float acc[CI];
#pragma unroll
    for (int j = 0; j < CI; ++j) acc[j] = 0.f;
    for (int tr = 0; tr < G.kh; ++tr)
        for (int tc = 0; tc < G.kw; ++tc) {
            const int th = ih + G.pad_h - tr * G.dil, tw = iw + G.pad_w - tc * G.dil;
            if (th < 0 || tw < 0) continue;
            const int oh = th / G.stride, ow = tw / G.stride;
            if (oh * G.stride != th || ow * G.stride != tw || oh >= G.ho || ow >= G.wo) continue;
            const int tap = tr * G.kw + tc;
            const T *dp = dc + (static_cast<long long>(nn * G.ho + oh) * G.wo + ow) * dc_cstride;
            if (G.groups == 1) {
                for (int co = 0; co < G.cout; ++co) {
                    const float dv = to_f32(dp[co]);
                    const T *wp = w + (static_cast<long long>(co) * taps + tap) * cig + ci0;
#pragma unroll
                    for (int j = 0; j < CI; ++j)
                        if (ci0 + j < G.cin) acc[j] += dv * to_f32(wp[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < CI; ++j) {
                    const int ci = ci0 + j;
                    if (ci >= G.cin) continue;
                    const int g = ci / cig;
                    for (int co = g * cog; co < (g + 1) * cog; ++co)
                        acc[j] += to_f32(dp[co]) * to_f32(w[(static_cast<long long>(co) * taps + tap) * cig + (ci - g * cig)]);
                }
            }
        }
#pragma unroll
    for (int j = 0; j < CI; ++j) {
        const int ci = ci0 + j;
        if (ci >= G.cin) continue;
        int p = 0;
        while (p + 1 < G.nparts && ci >= G.parts[p].choff + G.parts[p].c) ++p;
        if (O.ptr[p] == nullptr) continue;
        const float mv = mask_at(G, p, nn, ih, iw) ? 1.f : 0.f;
        static_cast<T *>(O.ptr[p])[m * O.cstride[p] + (ci - G.parts[p].choff)] = from_f32<T>(acc[j] * mv);
    }
}

// ------------------------------------------------------------------------------------------------
// wgrad: tiled outer-product GEMM  dw[co][k] += sum_p dc[p][co] * patch[p][k],  k = (tap, ci_local),
// patch gathered on the fly (im2col with hole / padding zeros).  Block tile CO_T x K_T, thread tile 4x4,
// pixel chunks of PC staged through shared memory; partial sums over pixel slabs via fp32 atomics.
// ------------------------------------------------------------------------------------------------
template <typename T, int CO_T, int K_T, int PC>
__global__ void __launch_bounds__(256) generic_wgrad_kernel(const GParams G, const T *__restrict__ dc, int dc_cstride, float *__restrict__ dw,
                                                            int g, int pix_per_block) {
    static_assert((CO_T / 4) * (K_T / 4) == 256, "thread tiling");
    __shared__ float s_dc[PC][CO_T];
    __shared__ float s_px[PC][K_T];
    const int cig = G.cin / G.groups, cog = G.cout / G.groups;
    const int taps = G.kh * G.kw;
    const int kg = taps * cig;                       // reduction-free extent of k within this group
    const int k0 = blockIdx.x * K_T, cob = blockIdx.y * CO_T;   // co within the group
    const long long plane = static_cast<long long>(G.ho) * G.wo;
    const long long total = plane * G.n;
    const long long p_begin = static_cast<long long>(blockIdx.z) * pix_per_block;
    const long long p_end = min(total, p_begin + pix_per_block);
    const int tx = threadIdx.x % (K_T / 4), ty = threadIdx.x / (K_T / 4);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (long long pc = p_begin; pc < p_end; pc += PC) {
        // stage dc[PC][CO_T]
        for (int i = threadIdx.x; i < PC * CO_T; i += 256) {
            const int pp = i / CO_T, cc = i - pp * CO_T;
            const long long m = pc + pp;
            float v = 0.f;
            if (m < p_end && cob + cc < cog) v = to_f32(dc[m * dc_cstride + g * cog + cob + cc]);
            s_dc[pp][cc] = v;
        }
        // stage patch[PC][K_T]
        for (int i = threadIdx.x; i < PC * K_T; i += 256) {
            const int pp = i / K_T, kk = i - pp * K_T;
            const long long m = pc + pp;
            const int k = k0 + kk;
            float v = 0.f;
            if (m < p_end && k < kg) {
                const int tap = k / cig, cl = k - tap * cig, ci = g * cig + cl;
                const int tr = tap / G.kw, tc = tap - tr * G.kw;
                const int nn = static_cast<int>(m / plane);
                const int rem = static_cast<int>(m - nn * plane), oh = rem / G.wo, ow = rem - oh * G.wo;
                const int hi = oh * G.stride - G.pad_h + tr * G.dil, wi = ow * G.stride - G.pad_w + tc * G.dil;
                if (hi >= 0 && hi < G.h && wi >= 0 && wi < G.w) {
                    int p = 0;
                    while (p + 1 < G.nparts && ci >= G.parts[p].choff + G.parts[p].c) ++p;
                    if (mask_at(G, p, nn, hi, wi)) {
                        const GPart &pt = G.parts[p];
                        v = to_f32(static_cast<const T *>(pt.x)[(static_cast<long long>(nn * (G.h >> pt.xup) + (hi >> pt.xup)) * (G.w >> pt.xup) + (wi >> pt.xup)) * pt.cstride + (ci - pt.choff)]);
                    }
                }
            }
            s_px[pp][kk] = v;
        }
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < PC; ++pp) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = s_dc[pp][ty * 4 + i]; b[i] = s_px[pp][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = cob + ty * 4 + i;
        if (co >= cog) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < kg) atomicAdd(dw + (static_cast<long long>(g * cog + co)) * kg + k, acc[i][j]);
        }
    }
}

template <typename T>
int launch_fwd(const GParams &G, const void *w, const float *bias, const float *msum, void *y, int ycs, cudaStream_t st) {
    const long long total = static_cast<long long>(G.n) * G.ho * G.wo;
    const unsigned gx = static_cast<unsigned>((total + 127) / 128);
    if (ycs <= 4) {
        generic_fwd_kernel<T, 4><<<dim3(gx, (ycs + 3) / 4), 128, 0, st>>>(G, static_cast<const T *>(w), bias, msum, static_cast<T *>(y), ycs);
    } else {
        generic_fwd_kernel<T, 8><<<dim3(gx, (ycs + 7) / 8), 128, 0, st>>>(G, static_cast<const T *>(w), bias, msum, static_cast<T *>(y), ycs);
    }
    PCB_LAUNCH_CHECK();
    return 0;
}

template <typename T>
int launch_dgrad(const GParams &G, const void *dc, int dcs, const void *w, const DxOut &O, cudaStream_t st) {
    const long long total = static_cast<long long>(G.n) * G.h * G.w;
    const unsigned gx = static_cast<unsigned>((total + 127) / 128);
    generic_dgrad_kernel<T, 8><<<dim3(gx, (G.cin + 7) / 8), 128, 0, st>>>(G, static_cast<const T *>(dc), dcs, static_cast<const T *>(w), O);
    PCB_LAUNCH_CHECK();
    return 0;
}

template <typename T>
int launch_wgrad(const GParams &G, const void *dc, int dcs, float *dw, cudaStream_t st) {
    const int cig = G.cin / G.groups, cog = G.cout / G.groups;
    const int kg = G.kh * G.kw * cig;
    const long long total = static_cast<long long>(G.n) * G.ho * G.wo;
    for (int g = 0; g < G.groups; ++g) {
        if (cog <= 8) {
            constexpr int CO_T = 8, K_T = 512, PC = 8;
            const int kt = (kg + K_T - 1) / K_T, ct = (cog + CO_T - 1) / CO_T;
            long long slabs = std::max<long long>(1, std::min<long long>((8ll * pcb_num_sms()) / std::max(1, kt * ct), (total + 2047) / 2048));
            const int ppb = static_cast<int>(((total + slabs - 1) / slabs + PC - 1) / PC * PC);
            slabs = (total + ppb - 1) / ppb;
            generic_wgrad_kernel<T, CO_T, K_T, PC><<<dim3(kt, ct, (unsigned)slabs), 256, 0, st>>>(G, static_cast<const T *>(dc), dcs, dw, g, ppb);
        } else {
            constexpr int CO_T = 64, K_T = 64, PC = 16;
            const int kt = (kg + K_T - 1) / K_T, ct = (cog + CO_T - 1) / CO_T;
            long long slabs = std::max<long long>(1, std::min<long long>((8ll * pcb_num_sms()) / std::max(1, kt * ct), (total + 1023) / 1024));
            const int ppb = static_cast<int>(((total + slabs - 1) / slabs + PC - 1) / PC * PC);
            slabs = (total + ppb - 1) / ppb;
            generic_wgrad_kernel<T, CO_T, K_T, PC><<<dim3(kt, ct, (unsigned)slabs), 256, 0, st>>>(G, static_cast<const T *>(dc), dcs, dw, g, ppb);
        }
        PCB_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace

int pcb_mask_sums(const pcb_conv *c, float *msum, uint8_t *newmask, cudaStream_t st) {
    GParams G;
    fill(G, c);
    const long long total = static_cast<long long>(c->n) * c->ho * c->wo;
    mask_sums_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(G, msum, newmask);
    PCB_LAUNCH_CHECK();
    return 0;
}

int pcb_generic_forward(const pcb_conv *c, const void *w, const float *bias, void *y, int y_cstride, const float *msum, cudaStream_t st) {
    GParams G;
    fill(G, c);
    if (c->dtype == PCB_BF16) return launch_fwd<bf16>(G, w, bias, msum, y, y_cstride, st);
    return launch_fwd<float>(G, w, bias, msum, y, y_cstride, st);
}

int pcb_generic_dgrad(const pcb_conv *c, const void *dc, int dc_cstride, const void *w_krsc, void *const *dx, const int *dx_cstride,
                      cudaStream_t st) {
    GParams G;
    fill(G, c);
    DxOut O;
    memset(&O, 0, sizeof(O));
    for (int p = 0; p < c->nparts; ++p) { O.ptr[p] = dx[p]; O.cstride[p] = dx_cstride[p]; }
    if (c->dtype == PCB_BF16) return launch_dgrad<bf16>(G, dc, dc_cstride, w_krsc, O, st);
    return launch_dgrad<float>(G, dc, dc_cstride, w_krsc, O, st);
}

int pcb_generic_wgrad(const pcb_conv *c, const void *dc, int dc_cstride, float *dw, bool zero_dw, cudaStream_t st) {
    GParams G;
    fill(G, c);
    const size_t bytes = sizeof(float) * c->cout * c->kh * c->kw * (c->cin / c->groups);
    if (zero_dw) PCB_CUDA(cudaMemsetAsync(dw, 0, bytes, st));
    if (c->dtype == PCB_BF16) return launch_wgrad<bf16>(G, dc, dc_cstride, dw, st);
    return launch_wgrad<float>(G, dc, dc_cstride, dw, st);
}
