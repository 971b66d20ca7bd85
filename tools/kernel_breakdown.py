"""Per-kernel time of one eager training step (torch.profiler / CUPTI, warm caches) -- a cheap substitute for an ncu launch
list when only the split between kernels is wanted.  Development tool; numbers taken under a profiler are not bench values."""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from text_segmentation_image_inpainting_b200 import _lib
from text_segmentation_image_inpainting_b200.engine import SegTrainStep, TrainStep
from text_segmentation_image_inpainting_b200.models.image_inpainting import ImageFillOrigin
from text_segmentation_image_inpainting_b200.models import text_segmentation as MT
from text_segmentation_image_inpainting_b200.synthetic import random_hole_masks

WORKLOAD = sys.argv[1] if len(sys.argv) > 1 else "unet"          # unet | textseg | xception

dev = torch.device("cuda:0")
_lib.load()
torch.manual_seed(0)
if WORKLOAD == "unet":
    ts = TrainStep(ImageFillOrigin().to(dev), compute_dtype=torch.bfloat16, process_group=None, use_graph=False)
    x = torch.randn(8, 3, 512, 512).to(dev)
    m = torch.from_numpy(random_hole_masks(8, 512, 512, seed=0)).to(dev)
else:
    cls, b = {"textseg": ("TextSegament", 8), "xception": ("XceptionTextSegment", 16)}[WORKLOAD]
    ts = SegTrainStep(getattr(MT, cls)().to(dev), compute_dtype=torch.bfloat16, process_group=None, use_graph=False)
    x = torch.randn(b, 3, 512, 512).to(dev)
    m = None
for _ in range(3):
    ts.step(x, m)
torch.cuda.synchronize()
STEPS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        ts.step(x, m)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", ev.name)
        name = re.sub(r"\(.*$", "", name)[:70]
        a = agg[name]
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
total = sum(a[1] for a in agg.values())
print(f"total kernel time per step {total / STEPS / 1e3:.3f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{a[1] / STEPS / 1e3:8.3f} ms {a[0] / STEPS:6.1f} launches  {k}")
