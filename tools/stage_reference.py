"""Stage the UNMODIFIED reference next to the repo so that it travels to the GPU box.

The reference (/root/reference, read-only, build container only) has no setup.py / pyproject, so the contract's
`pip install --target baseline/_ref /root/reference` cannot run; this does what that install would have done: a verbatim copy of
the importable Python sources (the `models` package, loss.py, Dataloader.py) into `baseline/_ref/`.  The directory is git-ignored
(never part of the history, never product source) but not gpurun-ignored, so

  * `bench.py --impl reference` can time the reference's OWN `models.image_inpainting.ImageFillOrigin` on the host cores
    (`cpu_baseline.kind == "reference"`), and
  * the `-m gpu` boundary tests can import the reference's own `models/image_inpainting.py` / `models/text_segmentation.py`
    on top of this repo's layer library on the GPU box, where /root/reference does not exist.

Nothing under baseline/_ref is imported by the product package."""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PCB_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
ITEMS = ["models", "loss.py", "Dataloader.py"]


def stage(verbose=False):
    """Returns the staged directory, or None when the reference is not available here and nothing was staged before."""
    if not os.path.isdir(os.path.join(SRC, "models")):
        return DST if os.path.isdir(os.path.join(DST, "models")) else None
    os.makedirs(DST, exist_ok=True)
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            for dirpath, dirnames, filenames in os.walk(s):
                dirnames[:] = [x for x in dirnames if x != "__pycache__"]
                rel = os.path.relpath(dirpath, s)
                os.makedirs(os.path.join(d, rel), exist_ok=True)
                for f in filenames:
                    if f.endswith(".py") or f.endswith(".md"):
                        a, b = os.path.join(dirpath, f), os.path.join(d, rel, f)
                        if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
                            shutil.copyfile(a, b)
        elif not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    if verbose:
        print("reference staged at", DST)
    return DST


def reference_dir():
    """Where the unmodified reference can be imported from on this machine (None if nowhere)."""
    if os.path.isdir(os.path.join(SRC, "models")):
        return SRC
    return DST if os.path.isdir(os.path.join(DST, "models")) else None


if __name__ == "__main__":
    print(stage(verbose=True))
    sys.exit(0)
