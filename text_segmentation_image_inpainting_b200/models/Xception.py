"""Slim Xception encoder -- mirror of the hot-path part of the reference's models/Xception.py (ResidualBlock :13-44,
Xception :47-114).  `XceptionClassifier` is out of scope and broken upstream."""
from torch import nn

from .BaseModels import BaseModule, Conv_block, DSConvBlock


class ResidualBlock(BaseModule):
    """Three depthwise-separable units (+ a strided 1x1 shortcut when shape changes), summed (reference :13-44)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=False, BN=True, activation=None,
                 expand_channel_first=True):
        super().__init__()
        mid = out_channels if expand_channel_first else in_channels
        self.conv = nn.Sequential(
            DSConvBlock(in_channels, mid, kernel_size, 1, padding, dilation, bias, BN, activation, activation),
            DSConvBlock(mid, out_channels, kernel_size, 1, padding, dilation, bias, BN, activation, activation),
            DSConvBlock(out_channels, out_channels, kernel_size, stride, padding, dilation, bias, BN, activation, None))
        if stride > 1 or in_channels != out_channels:
            self.residual_conv = nn.Sequential(*Conv_block(in_channels, out_channels, kernel_size=1, stride=stride, bias=False, BN=True,
                                                           activation=None))
        else:
            self.residual_conv = None

    def forward(self, x):
        shortcut = x if self.residual_conv is None else self.residual_conv(x)
        return self.conv(x) + shortcut


class Xception(BaseModule):
    def __init__(self, color_channel=3, act_fn=nn.LeakyReLU(0.3)):
        super().__init__()
        self.act_fn = act_fn
        rb = lambda i, o, s, r: ResidualBlock(i, o, 3, stride=s, padding=r, dilation=r, bias=False, BN=True, activation=act_fn)  # noqa: E731
        self.entry_flow_1 = nn.Sequential(                                                     # -> 1/4   (reference :56-65)
            *Conv_block(color_channel, 32, 3, stride=2, padding=1, bias=False, BN=True, activation=act_fn),
            *Conv_block(32, 64, 3, stride=1, padding=1, bias=False, BN=True, activation=act_fn),
            rb(64, 128, 2, 1))
        self.entry_flow_2 = nn.Sequential(rb(128, 256, 2, 1), rb(256, 512, 1, 2))               # -> 1/8   (:67-74)
        self.middle_flow = nn.Sequential(*[rb(512, 512, 1, 2) for _ in range(4)], *[rb(512, 512, 1, 4) for _ in range(4)])   # :76-92
        self.exit_flow = nn.Sequential(rb(512, 512, 1, 2), rb(512, 512, 1, 2), rb(512, 512, 1, 1), rb(512, 512, 1, 1))       # :94-105
        self.x4_feature_channels = 128
        self.last_feature_channels = 512

    def forward(self, x):
        x4 = self.entry_flow_1(x)
        x = self.exit_flow(self.middle_flow(self.entry_flow_2(x4)))
        return x, x4
