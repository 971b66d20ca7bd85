"""Text-segmentation encoder-decoders -- the workloads of BASELINE.json configs 2 and 4.  Mirror of the reference's
models/text_segmentation.py (TextSegament :18-84, XceptionTextSegment :87-114): identical module tree / state_dict."""
import torch
from torch import nn

from .. import ops
from .BaseModels import B200Conv2d, BaseModule, Conv_block
from .MobileNetV2 import DilatedMobileNetV2, InvertedResidual
from .Xception import Xception
from .common import ASP, RFB, B200AvgPool2d


class B200Upsample(nn.Upsample):
    def forward(self, x):
        if self.mode != "bilinear" or self.align_corners:
            raise NotImplementedError("only bilinear / align_corners=False (the reference's use)")
        return ops.bilinear_upsample(x, self.scale_factor)


class TextSegament(BaseModule):
    """Dilated MobileNetV2 (x2 width, scSE) + RFB pooling + DeepLabV3+-style decoder (reference :18-84)."""

    def __init__(self, encoder_checkpoint=None, free_last_blocks=-1, width_mult=2):
        super().__init__()
        self.act_fn = nn.LeakyReLU(0.3)
        self.encoder = DilatedMobileNetV2(width_mult=width_mult, activation=self.act_fn, bias=False, add_sece=True, add_partial=False)
        self.feature_avg_pool = B200AvgPool2d(kernel_size=3, stride=2, padding=1)
        feature_channels = sum(stage[0].out_channels for stage in self.encoder.features[3:])
        self.feature_pooling = RFB(feature_channels, 256, activation=self.act_fn, add_sece=True)
        concat_c = sum(stage[0].out_channels for stage in self.encoder.features[:3])
        self.feature_4x_conv = InvertedResidual(concat_c, 128, stride=1, expand_ratio=1, dilation=1, activation=self.act_fn, add_sece=True)
        self.smooth_feature_4x_conv = nn.Sequential(
            InvertedResidual(256 + 128, 128, stride=1, expand_ratio=1, dilation=2, activation=self.act_fn, add_sece=True),
            InvertedResidual(128, 128, stride=1, expand_ratio=1, dilation=1, activation=self.act_fn, add_sece=True))
        self.out_conv = nn.Sequential(B200Conv2d(128, 1, kernel_size=3, padding=1, bias=True, stride=1),
                                      B200Upsample(scale_factor=4, mode="bilinear", align_corners=False))
        self.initialize_weights()
        self.encoder.load_pre_train_checkpoint(encoder_checkpoint, free_last_blocks)

    def forward(self, x):
        shallow = []                                   # 1/2, 1/2, 1/4 feature maps
        for stage in self.encoder.features[:3]:
            x = stage(x)
            shallow.append(x)
        shallow = ops.concat_features([self.feature_avg_pool(shallow[0]), self.feature_avg_pool(shallow[1]), shallow[2]])
        deep = []                                      # 1/8 maps of the dilated stages
        for stage in self.encoder.features[3:]:
            x = stage(x)
            deep.append(x)
        x = self.feature_pooling(ops.concat_features(deep))
        x = ops.bilinear_upsample(x, 2)
        x = ops.concat_features([self.feature_4x_conv(shallow), x])
        return self.out_conv(self.smooth_feature_4x_conv(x))


class XceptionTextSegment(BaseModule):
    """Xception encoder + ASP pooling + 1/4-feature skip (reference :87-114)."""

    def __init__(self):
        super().__init__()
        self.act_fn = nn.LeakyReLU(0.3)
        self.encoder = Xception(color_channel=3, act_fn=self.act_fn)
        self.feature_pooling = ASP(self.encoder.last_feature_channels, 256, self.act_fn, asp_rate=(3, 5, 9))
        self.feature_4x_conv = nn.Sequential(*Conv_block(self.encoder.x4_feature_channels, 48, kernel_size=1, bias=False, BN=True,
                                                         activation=self.act_fn))
        self.out_conv = nn.Sequential(*Conv_block(48 + 256, 128, kernel_size=3, stride=1, padding=1, bias=False, BN=True,
                                                  activation=self.act_fn),
                                      B200Conv2d(128, 1, kernel_size=3, stride=1, padding=1))

    def forward(self, x):
        x, x4 = self.encoder(x)
        x4 = self.feature_4x_conv(x4)
        x = ops.bilinear_upsample(self.feature_pooling(x), 2)
        x = self.out_conv(ops.concat_features([x, x4]))
        return ops.bilinear_upsample(x, 4)
