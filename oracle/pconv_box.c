/* oracle/pconv_box.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C, ATen-independent restatement of the reference's hard-gated partial
 * convolution forward (models/partial_convolution.py:49-80) in its *box-sum*
 * form: the frozen all-ones `mask_conv` (partial_convolution.py:38-47,59,63) is
 * algebraically  s[n,g,p] = sum_{ci in group g} sum_{tap} m[n,ci,p*stride-pad+tap*dil]
 * (zero padding counts as hole), identical for every output channel of a group.
 *
 * Layout: NCHW fp32, exactly what the reference modules exchange.
 * The feature convolution accumulates in double so this file is a tighter
 * numerical yardstick than either ATen or the GPU kernels (compare with a
 * tolerance); the mask sum / hole flag / new mask are integer-exact and are
 * compared bit-for-bit.
 */
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Returns 0 on success, nonzero on bad arguments.
 * mask has `mask_channels` channels (>=1 if same_holes -- only channel 0 is read,
 * partial_convolution.py:59 -- else Cin).
 * y        [N,Cout,Ho,Wo]
 * msum     [N,Cout,Ho,Wo]   the renormaliser actually divided by (1 where hole), :66
 * new_mask [N,Cout,Ho,Wo]   {0,1}, :74-77                                           */
int pconv_box_forward(const float *x, const float *mask, const float *w, const float *bias,
                      int N, int Cin, int H, int W, int Cout, int kh, int kw,
                      int stride, int pad, int dil, int groups, int same_holes, int mask_channels,
                      float *y, float *msum, float *new_mask)
{
    if (groups <= 0 || Cin % groups || Cout % groups) return 1;
    if (!same_holes && mask_channels != Cin) return 2;
    if (same_holes && mask_channels < 1) return 2;
    const int Ho = (H + 2 * pad - dil * (kh - 1) - 1) / stride + 1;
    const int Wo = (W + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return 3;
    const int cig = Cin / groups, cog = Cout / groups;

    for (int n = 0; n < N; ++n)
    for (int g = 0; g < groups; ++g)
    for (int ho = 0; ho < Ho; ++ho)
    for (int wo = 0; wo < Wo; ++wo) {
        /* ---- box sum of the mask over the receptive field ---- */
        long s = 0;
        if (same_holes) {
            for (int r = 0; r < kh; ++r) for (int c = 0; c < kw; ++c) {
                int hi = ho * stride - pad + r * dil, wi = wo * stride - pad + c * dil;
                if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                s += (mask[((size_t)n * mask_channels * H + hi) * W + wi] != 0.0f);
            }
        } else {
            for (int ci = g * cig; ci < (g + 1) * cig; ++ci)
            for (int r = 0; r < kh; ++r) for (int c = 0; c < kw; ++c) {
                int hi = ho * stride - pad + r * dil, wi = wo * stride - pad + c * dil;
                if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                s += (mask[(((size_t)n * Cin + ci) * H + hi) * W + wi] != 0.0f);
            }
        }
        const int hole = (s == 0);                               /* :60,64 */
        /* :61 -- the same_holes count is scaled by the conv's *total* in_channels,
         * even for depthwise convs (reference quirk, SURVEY 7 "hard parts") */
        const float denom = hole ? 1.0f : (same_holes ? (float)(s * Cin) : (float)s);

        for (int co = g * cog; co < (g + 1) * cog; ++co) {
            double acc = 0.0;
            for (int cl = 0; cl < cig; ++cl) {
                const int ci = g * cig + cl;
                const int mc = same_holes ? 0 : ci;
                for (int r = 0; r < kh; ++r) for (int c = 0; c < kw; ++c) {
                    int hi = ho * stride - pad + r * dil, wi = wo * stride - pad + c * dil;
                    if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                    /* the reference multiplies by the FULL mask even when same_holes (:51) */
                    float mv = mask[(((size_t)n * mask_channels + (same_holes ? (ci < mask_channels ? ci : 0) : mc)) * H + hi) * W + wi];
                    float xv = x[(((size_t)n * Cin + ci) * H + hi) * W + wi] * mv;
                    acc += (double)xv * (double)w[(((size_t)co * cig + cl) * kh + r) * kw + c];
                }
            }
            const size_t o = (((size_t)n * Cout + co) * Ho + ho) * Wo + wo;
            const float b = bias ? bias[co] : 0.0f;
            y[o] = hole ? 0.0f : (float)(acc / (double)denom) + b;  /* :71-72 */
            msum[o] = denom;
            new_mask[o] = hole ? 0.0f : 1.0f;                    /* :74-75 */
        }
    }
    return 0;
}

#ifdef __cplusplus
}
#endif
