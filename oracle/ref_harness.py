"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (Yuki-11/CSBSR) in-process on CPU.

Used by `tests/golden/gen_golden.py` (run in the build container, where /root/reference exists) to
produce the committed golden fixtures, and by `bench.py --impl reference`'s optional local mode.
Nothing in the product package (`csbsr_b200/`) may import this file.  The recipe follows SURVEY.md
section 8(c) / Appendix D: shim packages for yacs / skimage / timm / matplotlib, NumPy-2 aliases,
pretrained-download suppression and "cuda" -> cpu redirection.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    """/root/reference in the build container; the verbatim copy baseline/install_ref.py made (baseline/_ref, git-ignored,
    travels with gpurun) on the GPU box."""
    cands = [os.environ.get("CSBSR_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "model")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_reference()
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_state = {}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def setup(device="cpu"):
    """Patch the process so that `import model....` resolves to the reference. Idempotent."""
    if _state:
        return _state
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import numpy as np
    import torch
    import torchvision

    sys.dont_write_bytecode = True
    sys.path[:0] = [_SHIMS, REFERENCE_ROOT]
    os.environ.setdefault("WANDB_MODE", "disabled")
    _state["cwd"] = os.getcwd()
    os.chdir(REFERENCE_ROOT)                      # build_model.py:235 uses a relative json path
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                           # surface_distance.py:255,261,322
    if not hasattr(np, "NaN"):
        np.NaN = np.nan

    if device == "cpu":
        def _fix(a):
            return ["cpu" if isinstance(x, str) and x.startswith("cuda") else x for x in a]
        _to = torch.Tensor.to
        torch.Tensor.to = lambda self, *a, **k: _to(self, *_fix(a), **k)
        torch.Tensor.cuda = lambda self, *a, **k: self
        _mto = torch.nn.Module.to
        torch.nn.Module.to = lambda self, *a, **k: _mto(self, *_fix(a), **k)
        torch.nn.Module.cuda = lambda self, *a, **k: self

    _vgg = torchvision.models.vgg16
    torchvision.models.vgg16 = lambda pretrained=False, **k: _vgg(weights=None)
    from torch.utils import model_zoo
    model_zoo.load_url = lambda *a, **k: {}
    import model.modeling.pspnet_pytorch.extractors as ex
    ex.load_weights_sequential = lambda target, src: None

    from model.config import cfg
    _state["cfg_proto"] = cfg
    return _state


def make_cfg(detector="PSPNet", wf_amp=0.0):
    st = setup()
    cfg = st["cfg_proto"].clone()
    cfg.merge_from_file(os.path.join(REFERENCE_ROOT, "config", "config_csbsr_pspnet.yaml"))
    cfg.MODEL.SR_SCRATCH = True
    cfg.MODEL.DETECTOR_TYPE = detector
    cfg.SOLVER.SEG_FAIL_ORIENTED_WEIGHT4SS_AMP = wf_amp
    return cfg


def patch_hrnet_configer():
    """H_48_D_4_composite.json names an ImageNet checkpoint that is not here: null network.pretrained."""
    setup()
    import model.modeling.build_model as bm
    if not _state.get("configer_patched"):
        _orig = bm.set_configer

        def _no_pretrained(path):
            c = _orig(path)
            c.update(["network", "pretrained"], None)
            return c
        bm.set_configer = _no_pretrained
        _state["configer_patched"] = True


def joint_model(cfg):
    setup()
    import contextlib
    import io
    from model.modeling.build_model import JointModel
    patch_hrnet_configer()
    with contextlib.redirect_stdout(io.StringIO()):
        m = JointModel(cfg)
    return m.eval()


def degrade_fns():
    setup()
    from model.data.blur.blur import GaussianBlur, conv_kernel2d
    from model.data.transforms.transforms import FactorResize
    return GaussianBlur, conv_kernel2d, FactorResize


def metric_fns():
    setup()
    from model.engine.inference import calc_distance_metrics
    from model.utils.estimate_metrics import IoU
    from model.utils.metrics.surface_distance.metrics import surface_distance
    return IoU, calc_distance_metrics, surface_distance
