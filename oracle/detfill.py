"""Deterministic parameter / input fill shared by the golden generator and the tests.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference initialises weights from torch's global RNG at construction
(models/partial_convolution.py:35,99) so two fresh models never agree; goldens
therefore overwrite every tensor of a state_dict with values that depend only
on the tensor's *name and shape* (a splitmix64 counter hash seeded by crc32(name)), which is
reproducible on any machine without shipping 130 MB of weights.
"""
import zlib

import numpy as np
import torch


def _splitmix_uniform(seed: int, n: int) -> np.ndarray:
    """n uniforms in [0,1) from a splitmix64 counter hash -- pure integer numpy ops, so the
    stream is identical on every machine / numpy version (no dependence on Generator internals)."""
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)


def det_tensor(name: str, shape, kind: str = "normal", scale: float = 1.0) -> torch.Tensor:
    """kind 'normal': zero-mean, std == scale (uniform-shaped, which is all the tests need);
    kind 'uniform': U[0, scale)."""
    n = int(np.prod(shape)) if len(shape) else 1
    u = _splitmix_uniform(zlib.crc32(name.encode()), n)
    if kind == "normal":
        a = (u - np.float32(0.5)) * np.float32(3.4641016) * np.float32(scale)
    elif kind == "uniform":
        a = u * np.float32(scale)
    else:
        raise ValueError(kind)
    return torch.from_numpy(np.ascontiguousarray(a.reshape(tuple(shape)), dtype=np.float32))


def det_fill_state_dict(sd: dict) -> dict:
    """Return a new state_dict with every entry overwritten deterministically.

    * conv / linear weights : N(0, 2/fan_in)   (He-style, keeps activations O(1))
    * biases                : N(0, 0.1)
    * BN weight             : U(0.5, 1.5);  BN bias: N(0, 0.1)
    * running_mean          : N(0, 0.1);    running_var: U(0.5, 1.5)
    * num_batches_tracked   : 0
    * mask_conv.weight      : left at 1.0 (frozen all-ones, partial_convolution.py:44-47)
    """
    out = {}
    for k, v in sd.items():
        shp = tuple(v.shape)
        if k.endswith("mask_conv.weight"):
            t = torch.ones(shp, dtype=torch.float32)
        elif k.endswith("num_batches_tracked"):
            t = torch.zeros(shp, dtype=v.dtype)
        elif k.endswith("running_mean"):
            t = det_tensor(k, shp, "normal", 0.1)
        elif k.endswith("running_var"):
            t = det_tensor(k, shp, "uniform", 1.0) + 0.5
        elif k.endswith(".weight") and len(shp) == 1:      # BN gamma
            t = det_tensor(k, shp, "uniform", 1.0) + 0.5
        elif k.endswith(".bias"):
            t = det_tensor(k, shp, "normal", 0.1)
        elif k.endswith(".weight"):
            fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else shp[0]
            t = det_tensor(k, shp, "normal", float(np.sqrt(2.0 / max(fan_in, 1))))
        else:
            t = det_tensor(k, shp, "normal", 1.0)
        out[k] = t.to(v.dtype) if v.dtype.is_floating_point else t
    return out
