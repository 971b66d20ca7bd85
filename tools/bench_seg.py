"""Throughput of the segmentation workloads (BASELINE.json configs 2 and 4) on one B200: forward+backward,
bf16, train-mode BN, loss = out.abs().mean(); CUDA-graph replay of the captured step (no optimiser).
Prints one JSON line per workload; not the driver's headline bench (that is bench.py / config 3)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from text_segmentation_image_inpainting_b200 import _lib, ops  # noqa: E402
from text_segmentation_image_inpainting_b200.models import text_segmentation as MT  # noqa: E402

GF_PER_IMAGE = {"TextSegament": 90.67, "XceptionTextSegment": 148.70}       # SURVEY 8d (forward, feature convs)


def run(cls_name, batch, steps=10, warmup=3):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = getattr(MT, cls_name)().to(dev).train()
    x = torch.randn(batch, 3, 512, 512, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def step():
        for p in net.parameters():
            p.grad = None
        ops.bump_weight_epoch()
        loss = ops.l1_mean(net(x))
        loss.backward()
        return loss.detach()
    before = _lib.launch_count()
    step()
    launches = _lib.launch_count() - before
    step()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    for _ in range(warmup):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flop = 3 * GF_PER_IMAGE[cls_name] * 1e9 * batch
    print(json.dumps({"workload": f"{cls_name} 512x512 batch {batch} fwd+bwd bf16", "ms_per_step": ms, "images_per_sec": batch / (ms * 1e-3),
                      "algorithmic_tflops": flop / (ms * 1e-3) / 1e12, "launches_per_step": launches, "loss": float(loss)}), flush=True)


if __name__ == "__main__":
    run("TextSegament", 8)
    run("XceptionTextSegment", 16)
