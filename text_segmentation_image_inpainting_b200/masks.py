"""HoleMask -- how hole masks travel between the partial-convolution modules.

The reference passes masks as dense fp32 ``[N, C, H, W]`` tensors and materialises ``torch.cat`` of them at
every decoder level (models/image_inpainting.py:185) although every mask in its networks is constant across
the channels of a concat group (it starts as one plane repeated 3x, Dataloader.py:128-129, and
``same_holes`` layers emit stride-0 expands, models/partial_convolution.py:76-77).  A dense mask at the last
decoder level of ``ImageFillOrigin`` (b=8) would be 402 MB of fp32 -- more traffic than the convolution.

``HoleMask`` is a ``torch.Tensor`` wrapper subclass with the same logical shape/dtype (so the reference's
model code, which only moves masks around, ``torch.cat``s them and hands them to the L1 modules, runs
unchanged) whose storage is a list of parts ``(uint8 plane [N, H>>up, W>>up], channels, up)``:
  * ``torch.cat(..., dim=1)`` concatenates part lists (no data movement),
  * nearest x2 upsampling bumps ``up`` (no data movement; the conv gathers ``plane[h>>1, w>>1]``),
  * any other torch function first materialises the dense fp32 tensor (semantically identical, slow).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib

Part = Tuple[torch.Tensor, int, int]      # (plane u8 [N, H>>up, W>>up], channels, up)

_CAT_FUNCS = {torch.cat, getattr(torch, "concat", torch.cat), getattr(torch, "concatenate", torch.cat)}
_META_NAMES = {"shape", "size", "dim", "ndim", "device", "dtype", "is_cuda", "requires_grad", "numel", "ndimension",
               "layout", "is_floating_point", "__len__", "grad_fn", "is_leaf", "names", "grad", "_version", "is_sparse",
               "is_quantized", "is_meta", "is_complex", "data_ptr"}


def _stream():
    return torch.cuda.current_stream().cuda_stream


class HoleMask(torch.Tensor):
    @staticmethod
    def __new__(cls, parts: Sequence[Part], n: int, h: int, w: int):
        parts = list(parts)
        c = sum(p[1] for p in parts)
        dev = parts[0][0].device
        r = torch.Tensor._make_wrapper_subclass(cls, (n, c, h, w), dtype=torch.float32, device=dev, requires_grad=False)
        r._parts = _merge(parts)
        return r

    # ------------------------------------------------------------------ constructors
    @staticmethod
    def from_dense(mask: torch.Tensor, channel_uniform: Optional[bool] = None) -> "HoleMask":
        """dense fp32/any [N,C,H,W] {0,1} mask -> uint8 planes.

        channel_uniform=True : trust the caller that all channels are equal (the reference's data path repeats one
                               plane over RGB, Dataloader.py:128-129) -> a single plane, no check, no sync;
        channel_uniform=None : check on the device (one sync; skipped while a CUDA graph is being captured) and
                               collapse to one plane when possible, else one plane per channel (C <= 8);
        channel_uniform=False: one plane per channel, no check."""
        if isinstance(mask, HoleMask):
            return mask
        if not mask.is_cuda:
            raise _lib.PcbError("HoleMask.from_dense: the CUDA path needs CUDA tensors (no CPU fallback)")
        n, c, h, w = mask.shape
        m = mask.detach().to(torch.float32).contiguous()
        if channel_uniform:
            m = m[:, :1].contiguous()
            plane = torch.empty((1, n, h, w), dtype=torch.uint8, device=mask.device)
            _lib.check(_lib.load().pcb_mask_planes_from_dense(m.data_ptr(), n, 1, h, w, plane.data_ptr(), _stream()))
            return HoleMask([(plane[0], c, 0)], n, h, w)
        planes = torch.empty((c, n, h, w), dtype=torch.uint8, device=mask.device)
        _lib.check(_lib.load().pcb_mask_planes_from_dense(m.data_ptr(), n, c, h, w, planes.data_ptr(), _stream()))
        if channel_uniform is None and c > 1 and not torch.cuda.is_current_stream_capturing():
            if bool((planes == planes[:1]).all()):              # one sync, eager mode only
                return HoleMask([(planes[0], c, 0)], n, h, w)
        # genuinely per-channel masks (partial_convolution.py:62-64 accepts any [N,C,H,W] mask): one plane per channel.  A
        # convolution whose (source, plane) partition exceeds the kernels' part table (PCB_MAX_PARTS) takes the general
        # dense-mask route in ops.partial_conv -- slower, still exact, still on the GPU.
        return HoleMask([(planes[i], 1, 0) for i in range(c)], n, h, w)

    @staticmethod
    def from_plane(plane: torch.Tensor, channels: int, up: int = 0) -> "HoleMask":
        n, hs, ws = plane.shape
        return HoleMask([(plane, channels, up)], n, hs << up, ws << up)

    # ------------------------------------------------------------------ accessors
    @property
    def parts(self) -> List[Part]:
        return self._parts

    def first_channel(self) -> "HoleMask":
        """mask[:, :1] (what same_holes layers read, partial_convolution.py:59)."""
        plane, _, up = self._parts[0]
        n, _, h, w = self.shape
        return HoleMask([(plane, 1, up)], n, h, w)

    def expand_channels(self, c: int) -> "HoleMask":
        """mask[:, :1].expand(N, c, H, W)  (partial_convolution.py:77,104)."""
        plane, _, up = self._parts[0]
        n, _, h, w = self.shape
        return HoleMask([(plane, c, up)], n, h, w)

    def upsampled(self) -> "HoleMask":
        """nearest x2 (DoubleUpSample on the mask, partial_convolution.py:231)."""
        n, _, h, w = self.shape
        new = []
        for plane, c, up in self._parts:
            if up == 0:
                new.append((plane, c, 1))
            else:   # already lazily upsampled once: materialise that level, stay lazy for the new one
                new.append((plane.repeat_interleave(2, 1).repeat_interleave(2, 2), c, 1))
        return HoleMask(new, n, 2 * h, 2 * w)

    def dense(self) -> torch.Tensor:
        n, c, h, w = self.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=self.device)
        c0 = 0
        lib = _lib.load()
        for plane, cc, up in self._parts:
            _lib.check(lib.pcb_mask_plane_to_dense(plane.data_ptr(), n, h, w, up, out.data_ptr(), c, c0, cc, _stream()))
            c0 += cc
        return out

    def __repr__(self):  # noqa: D105
        with torch._C.DisableTorchFunctionSubclass():
            shp = tuple(self.shape)
        return f"HoleMask(shape={shp}, parts={[(tuple(p.shape), c, up) for p, c, up in self._parts]})"

    # ------------------------------------------------------------------ torch function protocol
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name in _META_NAMES or (name == "__get__" and getattr(getattr(func, "__self__", None), "__name__", "") in _META_NAMES):
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if func in _CAT_FUNCS:
            tensors = args[0]
            dim = kwargs.get("dim", args[1] if len(args) > 1 else 0)
            if dim in (1, -3):
                hm = [t if isinstance(t, HoleMask) else HoleMask.from_dense(t) for t in tensors]
                n, _, h, w = hm[0].shape
                return HoleMask([p for t in hm for p in t.parts], n, h, w)
        if func in (F.interpolate,) and isinstance(args[0], HoleMask):
            if kwargs.get("mode", "nearest") == "nearest" and kwargs.get("scale_factor") in (2, 2.0) and kwargs.get("size") is None:
                return args[0].upsampled()
        if func is torch.Tensor.expand_as and isinstance(args[0], HoleMask):
            src, other = args
            if src.shape[1] == 1:
                return src.expand_channels(other.shape[1])
        if func is torch.Tensor.__getitem__ and isinstance(args[0], HoleMask):
            idx = args[1]
            if isinstance(idx, tuple) and len(idx) >= 2 and idx[0] == slice(None) and idx[1] == slice(None, 1) and \
                    all(i == slice(None) for i in idx[2:]):
                return args[0].first_channel()
        # anything else: semantically-identical dense fallback
        def conv(a):
            if isinstance(a, HoleMask):
                return a.dense()
            if isinstance(a, (list, tuple)):
                return type(a)(conv(b) for b in a)
            return a
        with torch._C.DisableTorchFunctionSubclass():
            return func(*conv(args), **{k: conv(v) for k, v in kwargs.items()})

    __torch_dispatch__ = None  # all handling happens at the torch-function level


def _merge(parts: List[Part]) -> List[Part]:
    out: List[Part] = []
    for plane, c, up in parts:
        if out and out[-1][0] is plane and out[-1][2] == up:
            out[-1] = (plane, out[-1][1] + c, up)
        else:
            out.append((plane, int(c), int(up)))
    return out


def as_hole_mask(mask) -> HoleMask:
    return mask if isinstance(mask, HoleMask) else HoleMask.from_dense(mask)
