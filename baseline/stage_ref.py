"""Stages the UNMODIFIED reference files of the hot path into git-ignored baseline/_ref/ and imports them from there.

The reference (clolckliang/unet-pytorch) is a directory of Python scripts with no setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` cannot work; what the hot path needs is a handful of pure-Python
files.  `stage()` copies them byte for byte from /root/reference (present in the build container only) into
baseline/_ref/ -- git-ignored, NOT gpurun-ignored, so the copy travels to the GPU box with the snapshot.  Nothing here
is product code: only bench.py's reference / cpu_baseline legs, tests/ and scripts/ import it.

    files staged                          used for
    nets/*.py                             the reference models (nets.unet.Unet, TraditionalUnet, LightWeightUnet, ...)
    nets/unet_training.py                 CE_Loss / Focal_Loss / Dice_loss, weights_init, get_lr_scheduler
    utils/utils_fit.py                    fit_one_epoch / fit_one_epoch_no_val (the training loop body)
    utils/utils_metrics.py                f_score, fast_hist, per_class_iu, compute_mIoU
    utils/utils.py, utils/__init__.py     get_lr, cvtColor, resize_image, preprocess_input
    unet.py                               the predictor class (detect_image / get_FPS / get_miou_png)

`matplotlib` is not installed in this image and utils_metrics.py imports it at module level for a plotting helper the
hot path never reaches: `import_reference()` registers empty stub modules for it before importing.
"""
import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("B2U_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")

FILES = ["nets/__init__.py", "nets/unet.py", "nets/vgg.py", "nets/resnet.py", "nets/unet_training.py",
         "nets/TraditionalUnet.py", "nets/LightWeightUnet.py", "nets/UltraLightweightUnet.py",
         "nets/UltraLightweightUnet_large.py", "nets/UltraLightweightUnet_large_optimized.py",
         "nets/RepVGG_Unet.py", "nets/HybridEfficientSeg.py", "nets/SegNets.py",
         "utils/__init__.py", "utils/utils.py", "utils/utils_fit.py", "utils/utils_metrics.py", "unet.py"]


def stage(force=False):
    """Copies the files (unmodified) when /root/reference is present; returns the staged directory or None."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if available() else None
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        if not os.path.exists(src):
            continue
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return REF_DST


def available():
    return all(os.path.exists(os.path.join(REF_DST, rel)) for rel in ("nets/unet.py", "nets/unet_training.py",
                                                                       "utils/utils_fit.py", "utils/utils_metrics.py"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    try:
        importlib.import_module("matplotlib")
        return
    except Exception:
        pass
    m = types.ModuleType("matplotlib")
    m.use = lambda *a, **k: None
    p = types.ModuleType("matplotlib.pyplot")
    m.pyplot = p
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, p


class _RefNamespace:
    """Modules of the staged reference, imported under a private prefix so they never shadow (or get shadowed by) the
    product's own `nets` / `utils` packages.  `ref.nets.unet.Unet`, `ref.utils.utils_fit.fit_one_epoch_no_val`, ..."""

    def __init__(self, root):
        self.root = root

    def module(self, dotted):
        """Imports e.g. 'nets.unet' from the staged tree with sys.path/sys.modules restored afterwards for every name
        that is not part of the reference (the reference's files import each other as `nets.*` / `utils.*`)."""
        return _import_from(self.root, dotted)


_CACHE = {}


def _import_from(root, dotted):
    key = (root, dotted)
    if key in _CACHE:
        return _CACHE[key]
    _stub_matplotlib()
    saved = {k: v for k, v in sys.modules.items() if k == "nets" or k.startswith("nets.") or k == "utils" or k.startswith("utils.")
             or k == "unet"}
    for k in saved:
        del sys.modules[k]
    for (r, d), mod in _CACHE.items():          # earlier reference modules stay visible to their siblings
        if r == root:
            sys.modules[d] = mod
            if "." in d:
                sys.modules.setdefault(d.split(".")[0], _CACHE.get((root, d.split(".")[0]), mod))
    sys.path.insert(0, root)
    try:
        mod = importlib.import_module(dotted)
        for k, v in list(sys.modules.items()):
            if (k == "nets" or k.startswith("nets.") or k == "utils" or k.startswith("utils.") or k == "unet") and \
                    getattr(v, "__file__", "") and str(v.__file__).startswith(root):
                _CACHE[(root, k)] = v
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "nets" or k.startswith("nets.") or k == "utils" or k.startswith("utils.")
                  or k == "unet"]:
            del sys.modules[k]
        sys.modules.update(saved)
    _CACHE[key] = mod
    return mod


def import_reference(dotted):
    """import_reference('nets.unet') -> the staged reference's module (staging first if /root/reference is here)."""
    root = stage()
    if root is None or not available():
        raise ImportError("the reference is not staged: run `python -m baseline.stage_ref` where /root/reference exists "
                          "(baseline/_ref/ travels to the GPU box with the gpurun snapshot)")
    return _import_from(root, dotted)


def load_with_shims(rel_path, shims, name=None):
    """Executes ONE staged reference file (e.g. 'utils/utils_fit.py') as a fresh module while `shims` (dotted name ->
    module) temporarily replace entries of sys.modules -- the import swap INTEGRATION.md describes (`nets.unet_training`
    and `utils.utils_metrics` pointing at the drop-in), applied to the unmodified file.  Names of the reference that are
    not shimmed resolve to the staged reference itself."""
    import importlib.util
    root = stage()
    if root is None or not available():
        raise ImportError("the reference is not staged (baseline/_ref/)")
    _stub_matplotlib()
    # make the un-shimmed reference modules importable under their own names for the duration of the exec
    for dep in ("utils.utils", "nets.unet_training", "utils.utils_metrics"):
        if dep not in shims:
            _import_from(root, dep)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("nets", "utils")}
    for k in saved:
        del sys.modules[k]
    for (r, d), mod in _CACHE.items():
        if r == root:
            sys.modules[d] = mod
    for k, v in shims.items():
        sys.modules[k] = v
        top = k.split(".")[0]
        if top not in sys.modules:
            sys.modules[top] = types.ModuleType(top)
    try:
        spec = importlib.util.spec_from_file_location(name or ("b2u_shimmed_" + rel_path.replace("/", "_")[:-3]),
                                                      os.path.join(root, rel_path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("nets", "utils")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mod


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
