"""Inverted-residual blocks built from partial convolutions (mirror of the hot-path part of the
reference's models/MobileNetV2.py: PartialInvertedResidual :152-190).

The dense (non-partial) MobileNetV2 / InvertedResidual family of the segmentation encoder is listed as a
"next" row (SURVEY 8a rows Conv_block / DSConvBlock / InvertedResidual) and is not mirrored yet.
"""
from torch import nn

from .. import ops
from .BaseModels import BaseModule
from .partial_convolution import PartialActivatedBN, partial_convolution_block


class PartialInvertedResidual(BaseModule):
    """1x1 expand -> depthwise kxk (groups = mid, same_holes) -> 1x1 project, each a partial-conv block;
    identity shortcut when stride == 1 and in == out (reference :158,183-190)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, expansion=1, BN=True,
                 activation=True, bias=False, use_1_conv=False, no_holes_1_conv=False, same_holes=False, *args, **kwargs):
        super().__init__()
        self.res_connect = stride == 1 and in_channels == out_channels
        mid = int(in_channels * expansion)
        pw = dict(BN=BN, bias=bias, use_1_conv=use_1_conv, no_holes_1_conv=no_holes_1_conv)
        self.conv = nn.Sequential(
            partial_convolution_block(in_channels, mid, 1, 1, 0, 1, activation=activation, **pw),
            partial_convolution_block(mid, mid, kernel_size, stride, padding, dilation, groups=mid, BN=BN,
                                      activation=activation, bias=bias, same_holes=same_holes),
            partial_convolution_block(mid, out_channels, 1, 1, 0, 1, activation=None, **pw),
        )

    def forward(self, args):
        x, mask = args
        if not self.res_connect:
            return self.conv(args)
        # shortcut fused into the last block's BN(+identity act) pass when that block ends in a BN
        head, last = self.conv[:-1], self.conv[-1]
        y, m = head((x, mask))
        if len(last) == 2 and isinstance(last[1], PartialActivatedBN):
            y, m = last[0]((y, m))
            return last[1]((y, m), residual=ops.as_feature(x))
        y, m = last((y, m))
        return x + y, m
