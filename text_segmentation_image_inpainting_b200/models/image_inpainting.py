"""Inpainting U-Nets assembled from the partial-convolution blocks: the workloads of BASELINE.json configs
3 and 5.  Structural mirror of the reference's models/image_inpainting.py (ImageFill :9-86, ImageFillOrigin
:110-191, ImageFillOriginV2 :219-290) with identical module tree -> identical state_dict keys, so reference
checkpoints load.  (ImageFillOriginV3 is excluded: it raises a shape error in the reference itself)

Differences in *how* it runs: skip concatenation + nearest upsampling is never materialised (``ops.LazyCat``
for the features, ``HoleMask`` for the masks: the consumer's gather does the index math), and every layer is
one launch of libpconv_b200's implicit-GEMM kernel.
"""
import torch
from torch import nn

from .. import ops
from ..masks import as_hole_mask
from .BaseModels import BaseModule
from .MobileNetV2 import PartialInvertedResidual
from .partial_convolution import DoubleUpSample, partial_convolution_block as pcb


class _PartialUNet(BaseModule):
    """Shared encoder / (bottleneck) / decoder walk of the three networks (reference :67-86, :164-191, :272-290)."""

    def _bottleneck(self, x, mask):
        return x, mask

    def forward(self, args):
        x, mask = args                                  # mask: 1 = ground truth, 0 = hole
        x = ops.as_feature_padded(x)
        mask = as_hole_mask(mask)
        skips = [(x, mask)]
        for layer in self.encoder:
            x, mask = layer((x, mask))
            skips.append((x, mask))
        skips.pop()                                     # the deepest map is the decoder's input, not a skip
        x, mask = self._bottleneck(x, mask)
        for layer in self.decoder:
            sx, sm = skips.pop()
            # nearest x2 of (x, mask) + channel concat with the skip: never materialised -- the next partial
            # convolution gathers straight from both sources (features) / both planes (masks)
            xh = ops.LazyCat([x, sx], ups=(1, 0))
            mh = torch.cat([mask.upsampled(), sm], dim=1)
            x, mask = layer((xh, mh))
        return x


class ImageFillOrigin(_PartialUNet):
    """The paper-shaped dense partial-conv U-Net (reference :110-191)."""
    #           in,  out, k, s, p
    ENCODER = [(64, 128, 5, 2, 2), (128, 256, 5, 2, 2), (256, 512, 3, 2, 1), (512, 512, 3, 2, 1), (512, 512, 3, 2, 1),
               (512, 512, 3, 2, 1), (512, 512, 3, 2, 1)]
    DECODER = [(1024, 512, 3, 1, 1)] * 4 + [(768, 256, 3, 1, 1), (384, 128, 3, 1, 1), (192, 64, 3, 1, 1)]

    def __init__(self):
        super().__init__()
        self.double_upscale = DoubleUpSample(scale_factor=2, mode="nearest")
        self.encoder = nn.Sequential(
            pcb(3, 64, 7, 2, 3, 1, bias=True, BN=False, activation=nn.ReLU(), same_holes=True),
            *[nn.Sequential(pcb(i, o, k, s, p, 1, groups=1, BN=True, activation=nn.ReLU(), bias=False, same_holes=True))
              for i, o, k, s, p in self.ENCODER])
        self.decoder = nn.Sequential(
            *[nn.Sequential(pcb(i, o, k, s, p, 1, groups=1, BN=True, activation=nn.LeakyReLU(0.2), bias=False, same_holes=False))
              for i, o, k, s, p in self.DECODER],
            pcb(64 + 3, 3, 3, 1, 1, 1, bias=True, BN=False, activation=False, same_holes=False))


class DoublePartialResidual(BaseModule):
    """Two partial-conv blocks with padding == dilation == dilation_rate[i]; output conv2(conv1(x)) + conv1(x)
    (reference :194-216; the ctor's `padding`/`dilation` arguments are ignored there too)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, expansion=1, BN=True,
                 activation=True, bias=False, use_1_conv=False, no_holes_1_conv=False, same_holes=False,
                 dilation_rate=(1, 1), *args, **kwargs):
        super().__init__()
        kw = dict(BN=BN, activation=activation, bias=bias, use_1_conv=use_1_conv, no_holes_1_conv=no_holes_1_conv,
                  same_holes=same_holes)
        self.conv1 = pcb(in_channels, out_channels, kernel_size, stride, padding=dilation_rate[0], dilation=dilation_rate[0], **kw)
        self.conv2 = pcb(out_channels, out_channels, kernel_size, 1, padding=dilation_rate[1], dilation=dilation_rate[1], **kw)

    def forward(self, args):
        x1, m1 = self.conv1(args)
        if len(self.conv2) == 2 and hasattr(self.conv2[1], "bn_act"):       # residual folded into the BN+act pass
            y, m2 = self.conv2[0]((x1, m1))
            return self.conv2[1]((y, m2), residual=x1)
        x2, m2 = self.conv2((x1, m1))
        return x2 + x1, m2


class ImageFillOriginV2(_PartialUNet):
    """Double-conv residual variant (reference :219-290)."""
    ENCODER = [(64, 128), (128, 256), (256, 256), (256, 256), (256, 512), (512, 512), (512, 512)]
    DECODER = [(1024, 512), (1024, 512), (768, 256), (512, 256), (512, 256), (384, 128), (192, 64)]

    def __init__(self):
        super().__init__()
        self.double_upscale = DoubleUpSample(scale_factor=2, mode="nearest")
        lk = lambda: nn.LeakyReLU(0.2)  # noqa: E731
        self.encoder = nn.Sequential(
            pcb(3, 64, 5, 2, 2, 1, bias=False, BN=True, activation=lk(), same_holes=True),
            *self.make_layer_v2([(i, o, 3, 2, 1, 1, 1, 1) for i, o in self.ENCODER], lk(), True, (1, 2)))
        self.decoder = nn.Sequential(
            *self.make_layer_v2([(i, o, 3, 1, 1, 1, 1, 1) for i, o in self.DECODER], lk(), False, (2, 1)),
            pcb(64 + 3, 3, 3, 1, 1, 1, bias=True, BN=False, activation=nn.ReLU(), same_holes=False))

    @staticmethod
    def make_layer_v2(settings, act_fn, same_holes=False, dilation_rate=(1, 1)):
        return [nn.Sequential(DoublePartialResidual(i, o, k, s, p, d, BN=True, activation=act_fn, bias=False,
                                                    same_holes=same_holes, dilation_rate=dilation_rate))
                for i, o, k, s, p, d, t, n in settings]


class ImageFill(_PartialUNet):
    """MobileNet-style variant: depthwise partial convolutions in inverted-residual blocks (reference :9-86)."""
    #          in, out, k, s, p, d, t, n
    ENCODER = [(64, 128, 3, 2, 1, 1, 4, 2), (128, 256, 3, 2, 1, 1, 4, 2), (256, 256, 3, 2, 1, 1, 4, 2)]
    DILATED = [(256, 256, 3, 1, 2, 2, 4, 2), (256, 256, 3, 1, 4, 4, 4, 2), (256, 256, 3, 1, 8, 8, 4, 2)]
    DECODER = [(512, 256, 3, 1, 1, 1, 2, 1), (384, 128, 3, 1, 1, 1, 2, 1), (192, 32, 3, 1, 1, 1, 2, 1)]

    def __init__(self):
        super().__init__()
        self.act_fn = nn.LeakyReLU(0.3)
        self.double_upscale = DoubleUpSample(scale_factor=2, mode="nearest")
        self.encoder = nn.Sequential(pcb(3, 64, 7, 2, 3, 1, bias=True, BN=False, activation=self.act_fn),
                                     *self.make_layers(self.ENCODER, use_1_conv=True, same_holes=True))
        self.dilated_layers = nn.Sequential(*self.make_layers(self.DILATED, no_holes_1_conv=True, same_holes=True))
        self.decoder = nn.Sequential(*self.make_layers(self.DECODER, no_holes_1_conv=True, same_holes=True),
                                     pcb(32 + 3, 3, 3, 1, 1, 1, bias=True, BN=False, activation=False))

    def make_layers(self, settings, use_1_conv=False, no_holes_1_conv=False, same_holes=False):
        stages = []
        for in_c, out_c, k, s, p, d, t, n in settings:
            blocks = []
            for i in range(n):
                blocks.append(PartialInvertedResidual(in_c, out_c, k, s if i == 0 else 1, p, d, t, bias=False, BN=True,
                                                      activation=self.act_fn, use_1_conv=use_1_conv,
                                                      no_holes_1_conv=no_holes_1_conv, same_holes=same_holes))
                in_c = out_c
            stages.append(nn.Sequential(*blocks))
        return stages

    def _bottleneck(self, x, mask):
        return self.dilated_layers((x, mask))
