"""Synthetic benchmark inputs: seeded free-form hole masks (thick lines + ellipses + square dilation) in the
style of the reference's training data (Dataloader.py:142-162 random_masks, :119-121 10x10 dilation,
:128-129 one plane repeated over RGB).  numpy only; mask convention 1 = valid, 0 = hole."""
import numpy as np


def _draw_line(canvas, x0, y0, x1, y1, width):
    h, w = canvas.shape
    yy, xx = np.mgrid[0:h, 0:w]
    dx, dy = x1 - x0, y1 - y0
    L2 = float(dx * dx + dy * dy) or 1.0
    t = np.clip(((xx - x0) * dx + (yy - y0) * dy) / L2, 0.0, 1.0)
    d2 = (xx - (x0 + t * dx)) ** 2 + (yy - (y0 + t * dy)) ** 2
    canvas[d2 <= (width / 2.0) ** 2] = 0


def _draw_ellipse(canvas, cx, cy, ax, ay):
    h, w = canvas.shape
    yy, xx = np.mgrid[0:h, 0:w]
    canvas[((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0] = 0


def _dilate_holes(canvas, k):
    hole = (canvas == 0)
    h, w = hole.shape
    pad = k // 2
    p = np.pad(hole, ((pad, k - 1 - pad), (pad, k - 1 - pad)))
    out = np.zeros_like(hole)
    for dy in range(k):
        for dx in range(k):
            out |= p[dy:dy + h, dx:dx + w]
    res = np.ones_like(canvas)
    res[out] = 0
    return res


def random_hole_plane(h, w, rng: np.random.Generator, dilate=10):
    canvas = np.ones((h, w), dtype=np.uint8)
    s = min(h, w) / 512.0
    for _ in range(int(rng.integers(1, 6))):
        x0, x1 = rng.integers(0, w, size=2)
        y0, y1 = rng.integers(0, h, size=2)
        _draw_line(canvas, int(x0), int(y0), int(x1), int(y1), max(2.0, float(rng.integers(15, 21)) * s))
    for _ in range(int(rng.integers(1, 6))):
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        ax, ay = rng.integers(20, 71, size=2)
        _draw_ellipse(canvas, cx, cy, max(2.0, float(ax) * s), max(2.0, float(ay) * s))
    if dilate:
        canvas = _dilate_holes(canvas, max(1, int(round(dilate * s))))
    return canvas


def random_hole_masks(n, h, w, seed=0, channels=3):
    """[n, channels, h, w] float32 {0,1} masks, one plane per image repeated over channels."""
    rng = np.random.Generator(np.random.PCG64(seed))
    planes = np.stack([random_hole_plane(h, w, rng) for _ in range(n)]).astype(np.float32)
    return np.repeat(planes[:, None], channels, axis=1)
