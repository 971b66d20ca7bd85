"""scSE, ASP and RFB feature-pooling blocks -- mirror of the hot-path part of the reference's models/common.py
(SpatialChannelSqueezeExcitation :13-43, ASP :53-93, RFB :96-156).  The LSTM classifier head is out of scope."""
import torch
from torch import nn

from .. import ops
from .BaseModels import B200Conv2d, BaseModule, Conv_block


class B200AvgPool2d(nn.AvgPool2d):
    def forward(self, x):
        k = self.kernel_size if isinstance(self.kernel_size, int) else self.kernel_size[0]
        s = self.stride if isinstance(self.stride, int) else self.stride[0]
        p = self.padding if isinstance(self.padding, int) else self.padding[0]
        if self.ceil_mode or not self.count_include_pad or self.divisor_override is not None:
            raise NotImplementedError("AvgPool2d: only the reference's default options")
        return ops.avg_pool2d(x, k, s, p)


class SpatialChannelSqueezeExcitation(BaseModule):
    """y = x * cSE + x * sSE  (reference :32-43).  The squeeze (global average) and the gate application are
    single HBM passes; the two tiny Linear layers on [N, C] stay in torch."""

    def __init__(self, in_channel, reduction=16, activation=nn.ReLU()):
        super().__init__()
        linear_nodes = max(in_channel // reduction, 4)
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.channel_excite = nn.Sequential(nn.Linear(in_channel, linear_nodes), activation, nn.Linear(linear_nodes, in_channel),
                                            nn.Sigmoid())
        self.spatial_excite = nn.Sequential(nn.Conv2d(in_channel, 1, kernel_size=1, stride=1, padding=0, bias=False), nn.Sigmoid())

    def forward(self, x):
        x = ops.as_feature(x)
        squeeze = ops.global_avg_pool(x)                                   # [N, C] fp32
        cse = self.channel_excite(squeeze)                                 # [N, C]
        ws = self.spatial_excite[0].weight.view(-1)                        # [C]
        return ops.scse_gate(x, cse, ws)


def add_SCSE_block(model_block, in_channel=None):
    if in_channel is None:
        in_channel = model_block[0].out_channels
    model_block.add_module("SCSE", SpatialChannelSqueezeExcitation(in_channel))


class ASP(BaseModule):
    """Atrous spatial pyramid with average-pooled ("vortex") inputs: 3x3 conv || AvgPool(r) -> 3x3 dilated r conv for
    three rates, concat, 1x1 (reference :53-93)."""

    def __init__(self, in_channel=256, out_channel=256, act_fn=None, asp_rate=(3, 9, 27)):
        super().__init__()
        branches = [nn.Sequential(*Conv_block(in_channel, out_channel, kernel_size=3, stride=1, padding=1, bias=False, BN=True,
                                              activation=act_fn))]
        for r in asp_rate[:3]:
            branches.append(nn.Sequential(B200AvgPool2d(kernel_size=r, stride=1, padding=(r - 1) // 2),
                                          *Conv_block(in_channel, out_channel, kernel_size=3, stride=1, padding=r, dilation=r,
                                                      bias=False, BN=True, activation=act_fn)))
        self.asp = nn.Sequential(*branches)
        self.out_conv = nn.Sequential(*Conv_block(out_channel * 4, out_channel, kernel_size=1, bias=False, BN=True, activation=act_fn))

    def forward(self, x):
        pooled = ops.concat_features([branch(x) for branch in self.asp.children()])
        return self.out_conv(pooled)


class RFB(BaseModule):
    """Receptive-field block: four branches (1x1 | 1xk -> kx1, then depthwise 3x3 dilated 1/5/17/29), concat, linear 1x1
    (+scSE), plus a 1x1 skip; activation of the sum (reference :96-156)."""

    def __init__(self, in_channel, out_channel, activation, add_sece=False):
        super().__init__()
        asp_rate = [5, 17, 29]
        self.act_fn = activation
        self.input_down_channel = nn.Sequential(*Conv_block(in_channel, out_channel, kernel_size=1, bias=True, BN=True,
                                                            activation=activation))
        linear = [B200Conv2d(out_channel * 4, out_channel, kernel_size=1, bias=True)]
        if add_sece:
            linear.append(SpatialChannelSqueezeExcitation(in_channel=out_channel, activation=activation))
        self.rfb_linear_conv = nn.Sequential(*linear)
        self.rfb = nn.Sequential(
            self.make_pooling_branch(in_channel, out_channel, out_channel, 1, 1, activation, half_conv=False),
            *[self.make_pooling_branch(in_channel, out_channel // 2, out_channel, k, r, activation, half_conv=True)
              for k, r in zip((3, 5, 7), asp_rate)])

    @staticmethod
    def make_pooling_branch(in_channel, mid_channel, out_channel, conv_kernel, astro_rate, activation, half_conv=False):
        dw = Conv_block(out_channel, out_channel, kernel_size=3, dilation=astro_rate, padding=astro_rate, bias=False, BN=True,
                        activation=activation, groups=out_channel)
        if not half_conv:
            return nn.Sequential(*Conv_block(in_channel, out_channel, kernel_size=conv_kernel, padding=(conv_kernel - 1) // 2,
                                             bias=False, BN=True, activation=activation), *dw)
        p = (conv_kernel - 1) // 2
        return nn.Sequential(
            *Conv_block(in_channel, mid_channel, kernel_size=1, padding=0, bias=False, BN=True, activation=activation),
            *Conv_block(mid_channel, 3 * mid_channel // 2, kernel_size=(1, conv_kernel), padding=(0, p), bias=False, BN=True,
                        activation=None),
            *Conv_block(3 * mid_channel // 2, out_channel, kernel_size=(conv_kernel, 1), padding=(p, 0), bias=False, BN=True,
                        activation=None),
            *dw)

    def forward(self, x):
        pooled = ops.concat_features([branch(x) for branch in self.rfb.children()])
        pooled = self.rfb_linear_conv(pooled)
        return ops.activation_only(pooled + self.input_down_channel(x), self.act_fn)
