"""autograd Functions over the C ABI (include/pconv_b200.h).  Tensors are NHWC in memory
(``torch.channels_last``) with the usual logical NCHW shape; dtype fp32 (exact mode) or bf16
(tensor-core mode).  There is no CPU / eager fallback: non-CUDA input raises."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_RELU6, PCB_BF16, PCB_F32, Conv, Part
from .masks import HoleMask, as_hole_mask

CL = torch.channels_last


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return PCB_F32
    if t.dtype == torch.bfloat16:
        return PCB_BF16
    raise _lib.PcbError(f"unsupported dtype {t.dtype}: the B200 path computes in float32 or bfloat16")


def as_feature(x: torch.Tensor) -> torch.Tensor:
    """CUDA + NHWC-contiguous view/copy of a 4-D activation (no dtype change)."""
    if not x.is_cuda:
        raise _lib.PcbError("text_segmentation_image_inpainting_b200 ops need CUDA tensors: there is no CPU fallback")
    if x.dim() != 4:
        raise _lib.PcbError(f"expected a 4-D NCHW activation, got shape {tuple(x.shape)}")
    _dtype_code(x)
    return x if x.is_contiguous(memory_format=CL) else x.contiguous(memory_format=CL)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def act_code(act) -> Tuple[int, float]:
    """nn.Module activation instance (or None/False) -> (code, negative slope)."""
    import torch.nn as nn
    if act is None or act is False:
        return ACT_NONE, 0.0
    if isinstance(act, nn.LeakyReLU):
        return ACT_LEAKY, float(act.negative_slope)
    if isinstance(act, nn.ReLU6):
        return ACT_RELU6, 0.0
    if isinstance(act, nn.ReLU):
        return ACT_RELU, 0.0
    raise NotImplementedError(f"activation {type(act).__name__} has no fused B200 kernel (supported: ReLU, ReLU6, LeakyReLU)")


# ------------------------------------------------------------------------------------------------
# partial convolution
# ------------------------------------------------------------------------------------------------
class ConvGeom:
    """Plain description of one partial-convolution call (everything but the pointers)."""

    def __init__(self, x_shape, cout, k, stride, padding, dilation, groups, same_holes, no_guard, dtype_code,
                 mask_parts: Sequence[Tuple[torch.Tensor, int, int]], plain=False):
        n, cin, h, w = x_shape
        kh, kw = (k, k) if isinstance(k, int) else k
        ph, pw = (padding, padding) if isinstance(padding, int) else padding
        s = stride if isinstance(stride, int) else stride[0]
        d = dilation if isinstance(dilation, int) else dilation[0]
        if not isinstance(stride, int) and stride[0] != stride[1] or not isinstance(dilation, int) and dilation[0] != dilation[1]:
            raise NotImplementedError("anisotropic stride / dilation")
        self.n, self.cin, self.h, self.w = n, cin, h, w
        self.cout, self.kh, self.kw, self.stride, self.ph, self.pw, self.dil, self.groups = cout, kh, kw, s, ph, pw, d, groups
        self.ho = (h + 2 * ph - d * (kh - 1) - 1) // s + 1
        self.wo = (w + 2 * pw - d * (kw - 1) - 1) // s + 1
        self.same_holes, self.no_guard, self.dtype, self.plain = int(same_holes), int(no_guard), dtype_code, int(plain)
        self.mask_parts = list(mask_parts)
        self.mg = groups if (groups > 1 and not same_holes) else 1
        if self.ho <= 0 or self.wo <= 0:
            raise _lib.PcbError(f"convolution output would be empty ({self.ho}x{self.wo})")

    def struct(self, x: Optional[torch.Tensor]) -> Conv:
        c = Conv()
        c.n, c.h, c.w, c.cin, c.cout, c.kh, c.kw = self.n, self.h, self.w, self.cin, self.cout, self.kh, self.kw
        c.stride, c.pad_h, c.pad_w, c.dil, c.groups, c.ho, c.wo = self.stride, self.ph, self.pw, self.dil, self.groups, self.ho, self.wo
        c.dtype, c.same_holes, c.no_guard, c.plain = self.dtype, self.same_holes, self.no_guard, self.plain
        esz = 2 if self.dtype == PCB_BF16 else 4
        parts = self.mask_parts
        if len(parts) > _lib.MAX_PARTS:
            raise NotImplementedError(f"more than {_lib.MAX_PARTS} mask parts")
        c.nparts = len(parts)
        off = 0
        for i, (plane, ch, up) in enumerate(parts):
            c.parts[i].x = (x.data_ptr() + off * esz) if x is not None else None
            c.parts[i].mask = plane.data_ptr() if plane is not None else None
            c.parts[i].c, c.parts[i].x_cstride, c.parts[i].x_up, c.parts[i].mask_up = ch, self.cin, 0, up
            off += ch
        if off != self.cin:
            raise _lib.PcbError(f"mask covers {off} channels but the input has {self.cin}")
        return c


_PROFILE = None     # optional list: (kind, geom, start_event, end_event) appended per conv kernel call


def set_profile(sink):
    """bench.py: pass a list to record CUDA-event-bracketed conv launches (eager mode only), or None to stop."""
    global _PROFILE
    _PROFILE = sink


class _Timed:
    def __init__(self, kind, geom):
        self.kind, self.geom = kind, geom

    def __enter__(self):
        if _PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.e.record()
            _PROFILE.append((self.kind, self.geom, self.s, self.e))
        return False


class PartialConvFn(torch.autograd.Function):
    """y, msum, newmask = pconv(x, W, b | mask)   (models/partial_convolution.py:49-80 / :121-137)."""

    @staticmethod
    def forward(ctx, x, weight, bias, geom: ConvGeom, wprep):
        lib = _lib.load()
        w_krsc, w_crsk = wprep
        c = geom.struct(x)
        y = torch.empty((geom.n, geom.cout, geom.ho, geom.wo), dtype=x.dtype, device=x.device, memory_format=CL)
        msum = torch.empty((geom.mg, geom.n, geom.ho, geom.wo), dtype=torch.float32, device=x.device)
        newmask = torch.empty((geom.mg, geom.n, geom.ho, geom.wo), dtype=torch.uint8, device=x.device)
        ws_bytes = lib.pcb_pconv_workspace(ctypes.byref(c))
        ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=x.device)
        b32 = bias.detach().float().contiguous() if bias is not None else None
        with _Timed("fwd", geom):
            _lib.check(lib.pcb_pconv_forward(ctypes.byref(c), w_krsc.data_ptr(), _ptr(b32), y.data_ptr(), msum.data_ptr(),
                                             newmask.data_ptr(), ws.data_ptr(), _stream()))
        ctx.geom, ctx.wprep, ctx.has_bias = geom, wprep, bias is not None
        ctx.save_for_backward(x, msum)
        ctx.mark_non_differentiable(msum, newmask)
        return y, msum, newmask

    @staticmethod
    def backward(ctx, gy, _gmsum, _gnewmask):
        lib = _lib.load()
        x, msum = ctx.saved_tensors
        geom: ConvGeom = ctx.geom
        w_krsc, w_crsk = ctx.wprep
        gy = gy.contiguous(memory_format=CL)
        if gy.dtype != x.dtype:
            gy = gy.to(x.dtype)
        c = geom.struct(x)
        dc = torch.empty_like(gy, memory_format=CL)
        dbias = torch.empty((geom.cout,), dtype=torch.float32, device=x.device) if ctx.has_bias else None
        _lib.check(lib.pcb_pconv_renorm_backward(ctypes.byref(c), gy.data_ptr(), msum.data_ptr(), dc.data_ptr(), _ptr(dbias), _stream()))
        dx = dw = None
        if ctx.needs_input_grad[1]:
            dw = torch.empty((geom.cout, geom.cin // geom.groups, geom.kh, geom.kw), dtype=torch.float32, device=x.device,
                             memory_format=CL)
            ws_bytes = lib.pcb_pconv_workspace(ctypes.byref(c))
            ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=x.device)
            with _Timed("wgrad", geom):
                _lib.check(lib.pcb_pconv_backward_weight(ctypes.byref(c), dc.data_ptr(), dw.data_ptr(), ws.data_ptr(), _stream()))
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x, memory_format=CL)
            with _Timed("dgrad", geom):
                _lib.check(lib.pcb_pconv_backward_data(ctypes.byref(c), dc.data_ptr(), w_krsc.data_ptr(), _ptr(w_crsk), dx.data_ptr(), _stream()))
        return dx, dw, dbias, None, None


_WEIGHT_EPOCH = 0


def bump_weight_epoch():
    """Invalidate every cached compute-dtype weight copy (call after updating parameters through raw pointers,
    e.g. `sgd_step`, which does not bump tensor version counters)."""
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def prepare_weight(weight: torch.Tensor, dtype: torch.dtype, groups: int, cache: dict):
    """fp32 master weight (OIHW logical) -> (KRSC, CRSK) copies in the compute dtype, cached per parameter version."""
    key = (weight.data_ptr(), weight._version, dtype, str(weight.device), _WEIGHT_EPOCH)
    hit = cache.get("key")
    if hit == key:
        return cache["val"]
    lib = _lib.load()
    wm = weight.detach()
    if wm.dtype != torch.float32:
        wm = wm.float()
    wm = wm.contiguous(memory_format=CL)          # physical [cout][kh][kw][cig]
    cout, cig, kh, kw = wm.shape
    code = PCB_BF16 if dtype == torch.bfloat16 else PCB_F32
    if code == PCB_F32:
        w_krsc = wm
        w_crsk = None
    else:
        w_krsc = torch.empty((cout, kh, kw, cig), dtype=dtype, device=wm.device)
        w_crsk = torch.empty((cig, kh, kw, cout), dtype=dtype, device=wm.device) if groups == 1 else None
        _lib.check(lib.pcb_weight_prepare(wm.data_ptr(), cout, kh, kw, cig, code, w_krsc.data_ptr(), _ptr(w_crsk), _stream()))
    cache["key"], cache["val"] = key, (w_krsc, w_crsk)
    return cache["val"]


def partial_conv(x, mask, weight, bias, stride, padding, dilation, groups, same_holes=False, no_guard=False, cache=None,
                 plain=False):
    """Returns (y, new_mask: HoleMask).  `mask` may be a HoleMask or a dense tensor; with ``plain=True`` the
    mask is ignored and an ordinary convolution is computed (same kernels, renormaliser 1)."""
    x = as_feature(x)
    n, cin, h, w = x.shape
    if plain:
        parts = [(None, cin, 0)]
    else:
        hm = as_hole_mask(mask)
        if tuple(hm.shape[2:]) != (h, w) or hm.shape[0] != n:
            raise _lib.PcbError(f"mask shape {tuple(hm.shape)} does not match input {tuple(x.shape)}")
        parts = hm.parts
        if hm.shape[1] != cin:
            if hm.shape[1] == 1:                       # broadcast of a 1-channel mask over x (x * mask, :51)
                parts = [(parts[0][0], cin, parts[0][2])]
            else:
                raise _lib.PcbError(f"mask has {hm.shape[1]} channels, input has {cin}")
    cout = weight.shape[0]
    geom = ConvGeom(x.shape, cout, tuple(weight.shape[2:]), stride, padding, dilation, groups, same_holes, no_guard,
                    _dtype_code(x), parts, plain=plain)
    wprep = prepare_weight(weight, x.dtype, groups, cache if cache is not None else {})
    y, msum, newmask = PartialConvFn.apply(x, weight, bias, geom, wprep)
    if geom.mg == 1:
        new = HoleMask.from_plane(newmask[0], cout, 0)
    else:
        cog = cout // groups
        new = HoleMask([(newmask[g], cog, 0) for g in range(groups)], n, geom.ho, geom.wo)
    return y, new


# ------------------------------------------------------------------------------------------------
# BatchNorm2d (+ activation, + residual)
# ------------------------------------------------------------------------------------------------
class BNActFn(torch.autograd.Function):
    """y = act(BN(x)) [+ residual]; BN optional (gamma None => plain activation)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, running_mean, running_var, nbt, training, momentum, eps, act, slope):
        lib = _lib.load()
        n, c, h, w = x.shape
        count = n * h * w
        code = _dtype_code(x)
        dev = x.device
        has_bn = gamma is not None
        scale = shift = mean = invstd = None
        if has_bn:
            scale = torch.empty((c,), dtype=torch.float32, device=dev)
            shift = torch.empty_like(scale)
            use_batch = training or running_mean is None
            if use_batch:
                if count <= 1 and training:
                    raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x.shape)}")
                sums = torch.empty((2, c), dtype=torch.float64, device=dev)
                _lib.check(lib.pcb_bn_stats(x.data_ptr(), code, count, c, sums[0].data_ptr(), sums[1].data_ptr(), _stream()))
                mean = torch.empty_like(scale)
                invstd = torch.empty_like(scale)
                _lib.check(lib.pcb_bn_finalize(sums[0].data_ptr(), sums[1].data_ptr(), count, c, gamma.data_ptr(), beta.data_ptr(),
                                               _ptr(running_mean) if training else None, _ptr(running_var) if training else None,
                                               _ptr(nbt) if training else None, float(momentum), float(eps), 1,
                                               scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), _stream()))
            else:
                _lib.check(lib.pcb_bn_finalize(None, None, count, c, gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                                               running_var.data_ptr(), None, float(momentum), float(eps), 0,
                                               scale.data_ptr(), shift.data_ptr(), None, None, _stream()))
        y = torch.empty_like(x, memory_format=CL)
        _lib.check(lib.pcb_bn_act_forward(x.data_ptr(), code, count, c, _ptr(scale), _ptr(shift), act, float(slope),
                                          _ptr(residual), y.data_ptr(), _stream()))
        ctx.cfg = (count, c, code, act, float(slope), has_bn, mean is not None, residual is not None)
        ctx.save_for_backward(x, scale, shift, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, scale, shift, mean, invstd = ctx.saved_tensors
        count, c, code, act, slope, has_bn, batch_stats, has_res = ctx.cfg
        gy = gy.contiguous(memory_format=CL)
        if gy.dtype != x.dtype:
            gy = gy.to(x.dtype)
        dx = torch.empty_like(x, memory_format=CL)
        dgamma = dbeta = None
        if has_bn and batch_stats:
            sums = torch.empty((2, c), dtype=torch.float64, device=x.device)
            _lib.check(lib.pcb_bn_act_backward_reduce(gy.data_ptr(), x.data_ptr(), code, count, c, scale.data_ptr(), shift.data_ptr(),
                                                      mean.data_ptr(), invstd.data_ptr(), act, slope, sums[0].data_ptr(),
                                                      sums[1].data_ptr(), _stream()))
            dgamma = torch.empty((c,), dtype=torch.float32, device=x.device)
            dbeta = torch.empty_like(dgamma)
            _lib.check(lib.pcb_bn_act_backward_apply(gy.data_ptr(), x.data_ptr(), code, count, c, scale.data_ptr(), shift.data_ptr(),
                                                     mean.data_ptr(), invstd.data_ptr(), act, slope, sums[0].data_ptr(),
                                                     sums[1].data_ptr(), 1, dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), _stream()))
        elif has_bn:   # eval-mode BN: a fixed affine map (parameter grads not produced in eval)
            _lib.check(lib.pcb_bn_act_backward_apply(gy.data_ptr(), x.data_ptr(), code, count, c, scale.data_ptr(), shift.data_ptr(),
                                                     None, None, act, slope, None, None, 0, dx.data_ptr(), None, None, _stream()))
        else:
            _lib.check(lib.pcb_bn_act_backward_apply(gy.data_ptr(), x.data_ptr(), code, count, c, None, None, None, None, act, slope,
                                                     None, None, 0, dx.data_ptr(), None, None, _stream()))
        return dx, dgamma, dbeta, (gy if has_res else None), None, None, None, None, None, None, None, None


def bn_act(x, bn, act, residual=None):
    """`bn`: nn.BatchNorm2d or None; `act`: nn activation module / None."""
    x = as_feature(x)
    code, slope = act_code(act)
    if x.shape[1] % 8 != 0:
        raise NotImplementedError("BatchNorm / activation kernels need channels % 8 == 0")
    if residual is not None:
        residual = as_feature(residual)
    if bn is None:
        return BNActFn.apply(x, None, None, residual, None, None, None, False, 0.0, 0.0, code, slope)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return BNActFn.apply(x, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                         bn.training, momentum, bn.eps, code, slope)


def activation_only(x, act, residual=None):
    x = as_feature(x)
    code, slope = act_code(act)
    if x.shape[1] % 8 != 0:      # tiny-channel tensors (e.g. the 3-channel tail): let torch do it
        y = act(x) if act else x
        return y + residual if residual is not None else y
    return BNActFn.apply(x, None, None, residual, None, None, None, False, 0.0, 0.0, code, slope)


# ------------------------------------------------------------------------------------------------
# nearest upsample / channel concat
# ------------------------------------------------------------------------------------------------
class ConcatFn(torch.autograd.Function):
    """cat([up2x?(x_i)], dim=1) in one pass (image_inpainting.py:183-184)."""

    @staticmethod
    def forward(ctx, ups: Tuple[int, ...], *xs):
        lib = _lib.load()
        x0 = xs[0]
        n = x0.shape[0]
        h, w = x0.shape[2] << ups[0], x0.shape[3] << ups[0]
        code = _dtype_code(x0)
        parts = (Part * len(xs))()
        ctot = 0
        for i, (x, up) in enumerate(zip(xs, ups)):
            if (x.shape[2] << up, x.shape[3] << up) != (h, w) or x.dtype != x0.dtype:
                raise _lib.PcbError("concat: mismatched spatial size or dtype")
            parts[i].x, parts[i].mask, parts[i].c, parts[i].x_cstride, parts[i].x_up, parts[i].mask_up = x.data_ptr(), None, x.shape[1], x.shape[1], up, 0
            ctot += x.shape[1]
        y = torch.empty((n, ctot, h, w), dtype=x0.dtype, device=x0.device, memory_format=CL)
        _lib.check(lib.pcb_concat_forward(parts, len(xs), code, n, h, w, y.data_ptr(), _stream()))
        ctx.meta = (ups, [x.shape[1] for x in xs], code, n, h, w)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        ups, cs, code, n, h, w = ctx.meta
        gy = gy.contiguous(memory_format=CL)
        outs: List[Optional[torch.Tensor]] = []
        ptrs = (ctypes.c_void_p * len(cs))()
        for i, (c, up) in enumerate(zip(cs, ups)):
            if ctx.needs_input_grad[i + 1]:
                g = torch.empty((n, c, h >> up, w >> up), dtype=gy.dtype, device=gy.device, memory_format=CL)
                outs.append(g)
                ptrs[i] = g.data_ptr()
            else:
                outs.append(None)
                ptrs[i] = None
        carr = (ctypes.c_int32 * len(cs))(*cs)
        uarr = (ctypes.c_int32 * len(cs))(*ups)
        _lib.check(lib.pcb_concat_backward(gy.data_ptr(), carr, uarr, len(cs), code, n, h, w, ptrs, _stream()))
        return (None, *outs)


def concat_features(xs: Sequence[torch.Tensor], ups: Optional[Sequence[int]] = None) -> torch.Tensor:
    xs = [as_feature(x) for x in xs]
    ups = tuple(int(u) for u in (ups if ups is not None else [0] * len(xs)))
    dt = xs[0].dtype
    xs = [x if x.dtype == dt else x.to(dt) for x in xs]
    return ConcatFn.apply(ups, *xs)


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    return ConcatFn.apply((1,), as_feature(x))


# ------------------------------------------------------------------------------------------------
# benchmark-step helpers: L1-mean loss and fused SGD
# ------------------------------------------------------------------------------------------------
class L1MeanFn(torch.autograd.Function):
    """loss = x.abs().mean()  (SURVEY 8d benchmark loss)."""

    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        x = x.contiguous(memory_format=CL) if x.dim() == 4 else x.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        scratch = torch.empty((1,), dtype=torch.float64, device=x.device)
        _lib.check(lib.pcb_l1_mean_forward(x.data_ptr(), _dtype_code(x), x.numel(), loss.data_ptr(), scratch.data_ptr(), _stream()))
        ctx.save_for_backward(x)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (x,) = ctx.saved_tensors
        gx = torch.empty_like(x)
        # the incoming gradient of a scalar loss is 1 in the benchmark; fold a general scale in on the host only
        # when it is a Python number -- otherwise multiply afterwards (keeps the kernel sync-free)
        _lib.check(lib.pcb_l1_mean_backward(x.data_ptr(), _dtype_code(x), x.numel(), 1.0 / x.numel(), gx.data_ptr(), _stream()))
        return gx * g.to(gx.dtype) if g is not None else gx


def l1_mean(x):
    return L1MeanFn.apply(as_feature(x) if x.dim() == 4 else x)


def sgd_step(param, grad, buf, lr, momentum=0.0, weight_decay=0.0, nesterov=False, first_step=False):
    lib = _lib.load()
    if not (param.is_contiguous() or param.is_contiguous(memory_format=CL)) or param.stride() != grad.stride():
        raise _lib.PcbError("sgd_step: param and grad must be dense with identical strides")
    _lib.check(lib.pcb_sgd_step(param.data_ptr(), grad.data_ptr(), _ptr(buf), param.numel(), float(lr), float(momentum),
                                float(weight_decay), int(nesterov), int(first_step), _stream()))
    bump_weight_epoch()
