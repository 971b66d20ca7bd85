"""Staged on-GPU self-check (run under gpurun).  Each stage is independent and failure-tolerant; everything
is written to gpurun_out/diag.log so one box lease yields the full picture.  Not part of the product."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "diag.log"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import pconv_torch as O  # noqa: E402
from oracle.detfill import det_fill_state_dict, det_tensor  # noqa: E402
from text_segmentation_image_inpainting_b200 import _lib, ops  # noqa: E402
from text_segmentation_image_inpainting_b200.masks import HoleMask  # noqa: E402
from text_segmentation_image_inpainting_b200.models import partial_convolution as PC  # noqa: E402
from text_segmentation_image_inpainting_b200.models import image_inpainting as PII  # noqa: E402

dev = torch.device("cuda:0")
RESULTS = {}


def pipeline_status():
    code = _lib.ctypes.c_int(0)
    torch.cuda.synchronize()
    _lib.check(_lib.load().pcb_debug_pipeline_status(_lib.ctypes.byref(code)))
    return code.value


def stage(name):
    def deco(fn):
        t0 = time.time()
        try:
            res = fn()
            st = pipeline_status()
            ok = bool(res) and st == 0
            RESULTS[name] = {"ok": ok, "pipeline": st}
            log(f"[{'PASS' if ok else 'FAIL'}] {name} (pipeline={st}, {time.time() - t0:.1f}s)")
        except Exception as e:  # noqa: BLE001
            RESULTS[name] = {"ok": False, "error": repr(e)}
            log(f"[ERROR] {name}: {e!r}")
            log(traceback.format_exc())
            try:
                log("pipeline status:", pipeline_status())
            except Exception as e2:  # noqa: BLE001
                log("pipeline status unavailable:", repr(e2))
        return fn
    return deco


def blob(n, c, h, w, seed, per_channel=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    m = np.ones((n, c, h, w), np.float32)
    for i in range(n):
        for ch in range(c if per_channel else 1):
            for _ in range(2):
                y0, x0 = rng.integers(0, max(1, h - 3)), rng.integers(0, max(1, w - 3))
                hh, ww = rng.integers(2, max(3, h // 2)), rng.integers(2, max(3, w // 2))
                if per_channel:
                    m[i, ch, y0:y0 + hh, x0:x0 + ww] = 0
                else:
                    m[i, :, y0:y0 + hh, x0:x0 + ww] = 0
    return torch.from_numpy(m)


def relerr(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    fin = torch.isfinite(b)
    if not torch.equal(torch.isfinite(a), fin):
        return float("inf")
    d = (a[fin] - b[fin]).abs().max().item() if fin.any() else 0.0
    return d / max(b[fin].abs().max().item() if fin.any() else 1.0, 1e-20)


def conv_case(tag, cin, cout, k, s, p, d, g, bias, same_holes, n, h, w, dtype, per_channel=False, cls="pc", tol=None,
              dump=False):
    """One PartialConv* module: fwd + bwd on the GPU vs the torch-CPU oracle (inputs rounded to `dtype`)."""
    if cls == "pc":
        mod = PC.PartialConv(cin, cout, k, s, p, d, g, bias, same_holes)
    elif cls == "1x1":
        mod = PC.PartialConv1x1(cin, cout, k, s, p, d, g, bias)
    else:
        mod = PC.PartialConvNoHoles(cin, cout, k, s, p, d, g, bias)
    sd = det_fill_state_dict(mod.state_dict())
    mod.load_state_dict(sd)
    x = det_tensor(tag + ".x", (n, cin, h, w))
    mask = blob(n, cin, h, w, seed=len(tag) * 7 + k, per_channel=per_channel)
    if cls == "nh":
        mask[:, 0] = 1.0
    wq = sd["feature_conv.weight"].to(dtype).float()
    xq = x.to(dtype).float()
    bq = sd["feature_conv.bias"] if bias else None
    # ---- oracle (CPU, fp32 arithmetic on the dtype-rounded operands)
    xo = xq.clone().requires_grad_(True); wo = wq.clone().requires_grad_(True)
    bo = bq.clone().requires_grad_(True) if bias else None
    if cls == "pc":
        yo, mo = O.partial_conv(xo, mask, wo, bo, s, p, d, g, same_holes)
    elif cls == "1x1":
        yo, mo = O.partial_conv_1x1(xo, mask, wo, bo, g)
    else:
        yo, mo = O.partial_conv_no_holes(xo, mask, wo, bo, s, p, d)
    gy = det_tensor(tag + ".gy", tuple(yo.shape)).to(dtype).float()
    (yo * gy).sum().backward()
    # ---- device
    with torch.no_grad():
        mod.feature_conv.weight.copy_(wq)
    mod = mod.to(dev)
    xd = xq.to(dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    yd, md = mod((xd, mask.to(dev)))
    yd.backward(gy.to(dev).to(dtype))
    torch.cuda.synchronize()
    tol = tol or (2e-2 if dtype == torch.bfloat16 else 1e-4)
    e_y = relerr(yd, yo.detach())
    e_gx = relerr(xd.grad, xo.grad)
    e_gw = relerr(mod.feature_conv.weight.grad, wo.grad)
    e_gb = relerr(mod.feature_conv.bias.grad, bo.grad) if bias else 0.0
    m_ok = torch.equal(md.dense().cpu(), mo.contiguous())
    ok = e_y <= tol and e_gx <= tol and e_gw <= tol and e_gb <= tol and m_ok
    c = ops.ConvGeom(xd.shape, cout, k, s, p, d, g, same_holes, cls == "nh", 1 if dtype == torch.bfloat16 else 0,
                     [(None, cin, 0)]).struct(xd)
    tc = _lib.load().pcb_conv_uses_tensor_cores(_lib.ctypes.byref(c))
    log(f"   {'ok ' if ok else 'BAD'} {tag:28s} tc={tc} y={e_y:.2e} gx={e_gx:.2e} gw={e_gw:.2e} gb={e_gb:.2e} mask={'eq' if m_ok else 'DIFF'}")
    if (not ok) and dump:
        np.save(os.path.join(OUT, f"dump_{tag}_y_dev.npy"), yd.detach().float().cpu().numpy())
        np.save(os.path.join(OUT, f"dump_{tag}_y_ref.npy"), yo.detach().numpy())
        np.save(os.path.join(OUT, f"dump_{tag}_gw_dev.npy"), mod.feature_conv.weight.grad.float().cpu().numpy())
        np.save(os.path.join(OUT, f"dump_{tag}_gw_ref.npy"), wo.grad.numpy())
        np.save(os.path.join(OUT, f"dump_{tag}_gx_dev.npy"), xd.grad.float().cpu().numpy())
        np.save(os.path.join(OUT, f"dump_{tag}_gx_ref.npy"), xo.grad.numpy())
    return ok


log("torch", torch.__version__, "device", torch.cuda.get_device_name(0), "lib", _lib.lib_path(), "version", _lib.load().pcb_version())


@stage("elementwise_bn_act")
def _():
    ok = True
    for dtype in (torch.float32, torch.bfloat16):
        for act, slope in ((torch.nn.ReLU(), 0), (torch.nn.LeakyReLU(0.2), 0.2), (None, 0), (torch.nn.ReLU6(), 0)):
            bn = torch.nn.BatchNorm2d(24)
            sd = det_fill_state_dict(bn.state_dict()); bn.load_state_dict(sd)
            ref = torch.nn.BatchNorm2d(24); ref.load_state_dict(sd)
            x = det_tensor("bn.x", (3, 24, 9, 11)) * 2 + 0.3
            xq = x.to(dtype).float()
            xr = xq.clone().requires_grad_(True)
            yr = ref(xr); yr = act(yr) if act else yr
            gy = det_tensor("bn.gy", tuple(yr.shape)).to(dtype).float()
            (yr * gy).sum().backward()
            bn = bn.to(dev)
            xd = xq.to(dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            yd = ops.bn_act(xd, bn, act)
            yd.backward(gy.to(dev).to(dtype))
            tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
            errs = [relerr(yd, yr.detach()), relerr(xd.grad, xr.grad), relerr(bn.weight.grad, ref.weight.grad),
                    relerr(bn.bias.grad, ref.bias.grad), relerr(bn.running_mean, ref.running_mean),
                    relerr(bn.running_var, ref.running_var)]
            good = all(e <= tol for e in errs) and int(bn.num_batches_tracked) == 1
            ok &= good
            log(f"   {'ok ' if good else 'BAD'} bn {dtype} act={type(act).__name__} errs={['%.1e' % e for e in errs]}")
    return ok


@stage("concat_upsample_masks")
def _():
    ok = True
    for dtype in (torch.float32, torch.bfloat16):
        for ca, cb in ((16, 8), (64, 3)):
            a = det_tensor("cat.a", (2, ca, 5, 6)).to(dtype); b = det_tensor("cat.b", (2, cb, 10, 12)).to(dtype)
            ar = a.float().requires_grad_(True); br = b.float().requires_grad_(True)
            yr = torch.cat([F.interpolate(ar, scale_factor=2, mode="nearest"), br], 1)
            gy = det_tensor("cat.gy", tuple(yr.shape)).to(dtype).float()
            (yr * gy).sum().backward()
            ad = a.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            bd = b.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            yd = ops.concat_features([ad, bd], ups=(1, 0))
            yd.backward(gy.to(dev).to(dtype))
            tol = 2e-2 if dtype == torch.bfloat16 else 1e-6
            e = [relerr(yd, yr.detach()), relerr(ad.grad, ar.grad), relerr(bd.grad, br.grad)]
            good = all(v <= tol for v in e); ok &= good
            log(f"   {'ok ' if good else 'BAD'} concat {dtype} c=({ca},{cb}) errs={['%.1e' % v for v in e]}")
    m = blob(2, 3, 8, 10, 3, per_channel=True)
    hm = HoleMask.from_dense(m.to(dev))
    good = torch.equal(hm.dense().cpu(), m) and torch.equal(hm.upsampled().dense().cpu(), F.interpolate(m, scale_factor=2, mode="nearest"))
    cat = torch.cat([hm.upsampled(), HoleMask.from_dense(F.interpolate(m, scale_factor=2).to(dev))], 1)
    good &= torch.equal(cat.dense().cpu(), torch.cat([F.interpolate(m, scale_factor=2)] * 2, 1))
    log(f"   {'ok ' if good else 'BAD'} HoleMask dense/upsample/cat round trip")
    return ok and good


@stage("generic_conv_fp32")
def _():
    ok = True
    f32 = torch.float32
    ok &= conv_case("g_k3", 4, 6, 3, 1, 1, 1, 1, True, False, 2, 20, 24, f32)
    ok &= conv_case("g_k3_sh", 4, 6, 3, 1, 1, 1, 1, True, True, 2, 20, 24, f32)
    ok &= conv_case("g_k3_pc", 4, 6, 3, 1, 1, 1, 1, False, False, 2, 20, 24, f32, per_channel=True)
    ok &= conv_case("g_k5_s2", 8, 8, 5, 2, 2, 1, 1, False, True, 2, 21, 26, f32)
    ok &= conv_case("g_k7_s2", 3, 8, 7, 2, 3, 1, 1, True, True, 2, 32, 32, f32)
    ok &= conv_case("g_d2", 4, 4, 3, 1, 2, 2, 1, False, False, 1, 24, 24, f32)
    ok &= conv_case("g_d4_sh", 4, 4, 3, 1, 4, 4, 1, True, True, 1, 24, 24, f32)
    ok &= conv_case("g_dw_sh", 8, 8, 3, 1, 1, 1, 8, False, True, 2, 16, 16, f32)
    ok &= conv_case("g_dw_s2", 8, 8, 3, 2, 1, 1, 8, False, True, 2, 17, 19, f32)
    ok &= conv_case("g_groups2", 4, 6, 3, 1, 1, 1, 2, True, False, 1, 12, 12, f32, per_channel=True)
    ok &= conv_case("g_1x1", 4, 8, 1, 1, 0, 1, 1, True, False, 2, 12, 12, f32, cls="1x1")
    ok &= conv_case("g_nh_1x1", 6, 4, 1, 1, 0, 1, 1, False, False, 2, 12, 12, f32, per_channel=True, cls="nh")
    ok &= conv_case("g_tail", 67, 3, 3, 1, 1, 1, 1, True, False, 1, 16, 16, f32)
    ok &= conv_case("g_bf16_k3", 8, 8, 3, 1, 1, 1, 1, True, False, 2, 12, 12, torch.bfloat16)
    return ok


@stage("tc_forward_backward_bf16")
def _():
    ok = True
    bf = torch.bfloat16
    # simplest first: pure GEMMs (1x1), one k-block, then several, then real convolutions
    ok &= conv_case("tc_1x1_k64_n64", 64, 64, 1, 1, 0, 1, 1, False, False, 1, 16, 16, bf, cls="1x1", dump=True)
    ok &= conv_case("tc_1x1_k128_n128", 128, 128, 1, 1, 0, 1, 1, True, False, 1, 16, 16, bf, cls="1x1", dump=True)
    ok &= conv_case("tc_1x1_k256_n64_m", 256, 64, 1, 1, 0, 1, 1, True, False, 2, 12, 13, bf, cls="1x1", dump=True)
    ok &= conv_case("tc_k3_64_64", 64, 64, 3, 1, 1, 1, 1, True, False, 1, 16, 16, bf, dump=True)
    ok &= conv_case("tc_k3_64_128_sh", 64, 128, 3, 1, 1, 1, 1, False, True, 2, 20, 24, bf, dump=True)
    ok &= conv_case("tc_k3_128_64", 128, 64, 3, 1, 1, 1, 1, True, False, 2, 20, 24, bf, dump=True)
    ok &= conv_case("tc_k5_s2_64_128", 64, 128, 5, 2, 2, 1, 1, False, True, 2, 21, 26, bf, dump=True)
    ok &= conv_case("tc_k3_s2_128_256", 128, 256, 3, 2, 1, 1, 1, False, True, 2, 18, 18, bf, dump=True)
    ok &= conv_case("tc_k3_d2_64_64", 64, 64, 3, 1, 2, 2, 1, False, False, 1, 24, 24, bf, dump=True)
    ok &= conv_case("tc_nh_1x1_128_64", 128, 64, 1, 1, 0, 1, 1, False, False, 2, 12, 12, bf, cls="nh", dump=True)
    ok &= conv_case("tc_k3_192_64", 192, 64, 3, 1, 1, 1, 1, False, False, 1, 32, 32, bf, dump=True)
    ok &= conv_case("tc_k3_512_512_tiny", 512, 512, 3, 2, 1, 1, 1, False, True, 2, 4, 4, bf, dump=True)
    return ok


def run_net(cls_name, n, hw, dtype, tol):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"net_{cls_name}.npz"))
    assert int(g["n"]) == n and int(g["hw"]) == hw
    step = int(g["step"])
    net = getattr(PII, cls_name)()
    sd = det_fill_state_dict(net.state_dict())
    net.load_state_dict(sd)
    net = net.to(dev).train()
    plane = np.unpackbits(g["mask_bits"])[: n * hw * hw].reshape(n, 1, hw, hw).astype(np.float32)
    mask = torch.from_numpy(np.repeat(plane, 3, 1))
    x = det_tensor(cls_name + ".x", (n, 3, hw, hw))
    xin = (x * mask).to(dev).to(dtype).contiguous(memory_format=torch.channels_last)
    t0 = time.time()
    out = net((xin, mask.to(dev)))
    loss = ops.l1_mean(out)
    loss.backward()
    torch.cuda.synchronize()
    dt = time.time() - t0
    e_out = relerr(out[..., ::step, ::step], torch.from_numpy(g["out_sub"]))
    e_loss = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
    errs = {"out": e_out, "loss": e_loss}
    params = dict(net.named_parameters())
    for k in [k for k in g.files if k.startswith("g.")]:
        errs[k] = relerr(params[k[2:]].grad, torch.from_numpy(g[k]))
    sdn = net.state_dict()
    for k in [k for k in g.files if k.startswith("bn.")][:4]:
        errs[k] = relerr(sdn[k[3:]], torch.from_numpy(g[k]))
    worst = max(errs.values())
    log(f"   {cls_name} {dtype} first-step wall {dt:.2f}s worst={worst:.2e} " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    return worst <= tol


@stage("network_fp32_exact_mode")
def _():
    ok = run_net("ImageFillOrigin", 2, 256, torch.float32, 2e-3)
    ok &= run_net("ImageFill", 2, 128, torch.float32, 2e-3)
    ok &= run_net("ImageFillOriginV2", 2, 256, torch.float32, 2e-3)
    return ok


@stage("network_bf16_tc_mode")
def _():
    ok = run_net("ImageFillOrigin", 2, 256, torch.bfloat16, 0.15)
    ok &= run_net("ImageFillOriginV2", 2, 256, torch.bfloat16, 0.15)
    ok &= run_net("ImageFill", 2, 128, torch.bfloat16, 0.15)
    return ok


@stage("bench_shape_step_timing")
def _():
    """cfg 3 shape (b8 @512^2 bf16): eager step timing + per-layer share, to direct optimisation."""
    net = PII.ImageFillOrigin().to(dev).train()
    x = torch.randn(8, 3, 512, 512, device=dev)
    plane = (torch.rand(8, 1, 512, 512, device=dev) > 0.1).float()
    mask = plane.expand(8, 3, 512, 512).contiguous()
    xin = (x * mask).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        for p in net.parameters():
            p.grad = None
        out = net((xin, mask)); loss = ops.l1_mean(out); loss.backward()
        torch.cuda.synchronize()
        log(f"   eager step {it}: {(time.time() - t0) * 1e3:.1f} ms  loss={float(loss):.4f} launches={_lib.launch_count()}")
    return True


json.dump(RESULTS, open(os.path.join(OUT, "diag.json"), "w"), indent=1)
log(json.dumps(RESULTS))
