"""One eager training step of the bench workload inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py
(launch list of exactly one step) and for `ncu --set full -k regex:...` captures.  Same network, batch and inputs as bench.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from text_segmentation_image_inpainting_b200 import _lib
from text_segmentation_image_inpainting_b200.engine import SegTrainStep, TrainStep
from text_segmentation_image_inpainting_b200.models.image_inpainting import ImageFillOrigin
from text_segmentation_image_inpainting_b200.models import text_segmentation as MT
from text_segmentation_image_inpainting_b200.synthetic import random_hole_masks

WORKLOAD = sys.argv[1] if len(sys.argv) > 1 else "unet"          # unet | textseg | xception

dev = torch.device("cuda:0")
_lib.load()
torch.manual_seed(0)
g = torch.Generator().manual_seed(1234)
if WORKLOAD == "unet":
    ts = TrainStep(ImageFillOrigin().to(dev), compute_dtype=torch.bfloat16, process_group=None, use_graph=False)
    x = torch.randn(8, 3, 512, 512, generator=g).to(dev)
    m = torch.from_numpy(random_hole_masks(8, 512, 512, seed=0)).to(dev)
else:
    cls, b = {"textseg": ("TextSegament", 8), "xception": ("XceptionTextSegment", 16)}[WORKLOAD]
    ts = SegTrainStep(getattr(MT, cls)().to(dev), compute_dtype=torch.bfloat16, process_group=None, use_graph=False)
    x = torch.randn(b, 3, 512, 512, generator=g).to(dev)
    m = None
for _ in range(2):
    ts.step(x, m)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ts.step(x, m)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
