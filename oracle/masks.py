"""Seeded hole-mask synthesis used by the golden generator -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The generator itself lives in the product package (it also feeds bench.py's synthetic inputs; the oracle may
import the product, never the reverse): thick lines + ellipses + square dilation after the reference's
Dataloader.py:142-162 / :119-121 / :128-129.  The reference rasterises with cv2/PIL, which cannot be
bit-matched without shipping cv2; this is a distribution-level restatement (blob/line holes that survive
several 3x3 layers)."""
from text_segmentation_image_inpainting_b200.synthetic import random_hole_masks, random_hole_plane  # noqa: F401
