"""Staged on-GPU self-check (run under gpurun): every conv / lazy-cat case + the networks, failure tolerant,
everything logged to gpurun_out/diag.log so one box lease yields the full picture.  Not part of the product."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "diag.log"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


import torch  # noqa: E402

import gpu_cases as G  # noqa: E402
from text_segmentation_image_inpainting_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
RESULTS = {}
ONLY = sys.argv[1:] or None


def pipeline_status():
    code = _lib.ctypes.c_int(0)
    torch.cuda.synchronize()
    _lib.check(_lib.load().pcb_debug_pipeline_status(_lib.ctypes.byref(code)))
    return code.value


def guarded(name, fn):
    if ONLY and not any(o in name for o in ONLY):
        return
    t0 = time.time()
    try:
        res = fn()
        st = pipeline_status()
        ok = bool(res.get("ok", False)) and st == 0
        RESULTS[name] = {"ok": ok, "pipeline": st}
        log(f"[{'ok ' if ok else 'BAD'}] {name:34s} pipeline={st} " +
            " ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}" for k, v in res.items() if k != "ok") + f" ({time.time() - t0:.1f}s)")
    except Exception as e:  # noqa: BLE001
        RESULTS[name] = {"ok": False, "error": repr(e)}
        log(f"[ERR] {name}: {e!r}")
        log(traceback.format_exc())
        try:
            log("pipeline status:", pipeline_status())
        except Exception as e2:  # noqa: BLE001
            log("pipeline status unavailable:", repr(e2))


log("torch", torch.__version__, "device", torch.cuda.get_device_name(0), "lib", _lib.lib_path())
for tag in sorted(G.CONV_CASES):
    guarded("conv:" + tag, lambda tag=tag: G.conv_case(tag, dev, dump_dir=OUT))
for tag in sorted(G.LAZYCAT_CASES):
    guarded("lazycat:" + tag, lambda tag=tag: G.lazycat_case(tag, dev))
for cls_name in ("ImageFillOrigin", "ImageFillOriginV2", "ImageFill"):
    for dt, tol in ((G.F32, 2e-3), (G.BF, 1e-1)):
        def run(cls_name=cls_name, dt=dt, tol=tol):
            errs = G.run_net(cls_name, dev, dt)
            errs["ok"] = max(errs.values()) <= tol
            return errs
        guarded(f"net:{cls_name}:{'f32' if dt == G.F32 else 'bf16'}", run)
for name in G.seg_block_builders():
    for dt in (G.F32, G.BF):
        def runb(name=name, dt=dt):
            errs = G.seg_block_case(name, dev, dt)
            errs["ok"] = max(errs.values()) <= (2e-4 if dt == G.F32 else 0.15)
            return errs
        guarded(f"seg:{name}:{'f32' if dt == G.F32 else 'bf16'}", runb)
for dt in (G.F32, G.BF):
    def runp(dt=dt):
        errs = G.pool_bilinear_case(dev, dt)
        errs["ok"] = max(errs.values()) <= (1e-5 if dt == G.F32 else 2e-2)
        return errs
    guarded(f"seg:pool_bilinear:{'f32' if dt == G.F32 else 'bf16'}", runp)
for cls_name in ("TextSegament", "XceptionTextSegment"):
    for dt, tol in ((G.F32, 2e-2), (G.BF, 1.0)):
        def runs(cls_name=cls_name, dt=dt, tol=tol):
            errs = G.run_segnet(cls_name, dev, dt)
            errs["ok"] = max(errs.values()) <= tol
            return errs
        guarded(f"segnet:{cls_name}:{'f32' if dt == G.F32 else 'bf16'}", runs)
json.dump(RESULTS, open(os.path.join(OUT, "diag.json"), "w"), indent=1)
bad = [k for k, v in RESULTS.items() if not v["ok"]]
log("FAILED:", bad if bad else "none")
