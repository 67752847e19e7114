"""Import the reference's OWN hot-path files, unmodified, on CPU
(TEST INFRASTRUCTURE; works only where ``/root/reference`` exists, i.e. in the
build container — never on the GPU box, and never from the product path).

``import deblur_e_nerf`` itself fails here (its ``__init__`` chain pulls in
``pytorch_lightning``, ``pypose``, ``torchmetrics``, ``lpips``), so this module

1. registers empty package objects for ``deblur_e_nerf`` and its sub-packages
   whose ``__path__`` points into ``/root/reference`` (sub-modules then import
   without running the package ``__init__`` files), and
2. puts the oracle restatements of ``easydict``, ``roma``, ``nerfacc`` and
   ``tinycudann`` into ``sys.modules`` under their upstream names,

after which ``load("models.pixel_bandwidth")`` etc. return the reference's
modules.  Used by ``tests/test_oracle_vs_reference.py`` and by
``tests/golden/make_golden.py`` (the committed generator of the fixtures).
"""

import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("DEN_REFERENCE_ROOT", "/root/reference")
_PKG = "deblur_e_nerf"
_SUBPACKAGES = ("utils", "data", "models", "external", "loss_metric")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, _PKG))


def install(nerfacc_module=None, tinycudann_module=None):
    """Register the shim.  Other ``nerfacc``/``tinycudann`` implementations (e.g.
    the CUDA drop-ins of this repo) can be injected for B1 drop-in tests."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    from . import easydict_ref, nerfacc_ref, roma_ref, tcnn_ref

    root = os.path.join(REFERENCE_ROOT, _PKG)
    if _PKG not in sys.modules or not getattr(sys.modules[_PKG], "_den_shim", False):
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [root]
        pkg._den_shim = True
        sys.modules[_PKG] = pkg
        for sub in _SUBPACKAGES:
            mod = types.ModuleType(f"{_PKG}.{sub}")
            mod.__path__ = [os.path.join(root, sub)]
            mod._den_shim = True
            sys.modules[f"{_PKG}.{sub}"] = mod
            setattr(pkg, sub, mod)

    sys.modules["easydict"] = easydict_ref
    sys.modules["roma"] = roma_ref
    sys.modules["nerfacc"] = nerfacc_module or nerfacc_ref
    sys.modules["tinycudann"] = tinycudann_module or tcnn_ref


def load(name):
    """``load("models.nerf")`` -> the reference's ``deblur_e_nerf/models/nerf.py``."""
    install_needed = _PKG not in sys.modules or not getattr(
        sys.modules[_PKG], "_den_shim", False)
    if install_needed:
        install()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        warnings.simplefilter("ignore", FutureWarning)
        module = importlib.import_module(f"{_PKG}.{name}")
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[f"{_PKG}.{parent}"], leaf, module)
    return module


def uninstall():
    for key in [k for k in sys.modules if k == _PKG or k.startswith(_PKG + ".")]:
        if getattr(sys.modules[key], "_den_shim", False) or key.startswith(_PKG + "."):
            del sys.modules[key]
    for key in ("easydict", "roma", "nerfacc", "tinycudann"):
        sys.modules.pop(key, None)


# --------------------------------------------------------------------------- #
# the reference's LightningModule, without Lightning
# --------------------------------------------------------------------------- #
def load_lightning_module():
    """Return the reference's ``DeblurENeRF`` class (models/deblur_e_nerf.py) with
    ``pytorch_lightning`` / ``pypose`` stubbed: only ``training_step`` and the
    render helpers it calls (lines 396-586, 1129-1308) are exercised."""
    import torch

    install()
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            pass

        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl
    if "pypose" not in sys.modules:
        sys.modules["pypose"] = types.ModuleType("pypose")
    for name in ("utils.autograd", "utils.modules", "utils.tensor_ops", "data.datasets",
                 "external.mlp", "external.sh_encoder", "external.ngp",
                 "external.vol_rendering", "external.utils", "external.optimizer",
                 "loss_metric.loss", "models.event_generation_params", "models.nerf",
                 "models.pixel_bandwidth", "models.trajectories"):
        try:
            load(name)
        except Exception:
            if name != "external.optimizer":
                raise
    if "deblur_e_nerf.models.offset_gamma_correction" not in sys.modules:
        stub = types.ModuleType("deblur_e_nerf.models.offset_gamma_correction")
        sys.modules["deblur_e_nerf.models.offset_gamma_correction"] = stub
        sys.modules["deblur_e_nerf.models"].offset_gamma_correction = stub
    return load("models.deblur_e_nerf").DeblurENeRF


def make_reference_module(hparams, components, train_ray_sample_batch_size=131072):
    """Instantiate the reference's LightningModule WITHOUT running its constructor
    (which needs a dataset on disk and Lightning): attributes used by ``training_step``
    are attached directly.  ``components`` maps attribute name -> reference module."""
    import torch

    cls = load_lightning_module()
    easydict = sys.modules["easydict"]
    inst = cls.__new__(cls)
    torch.nn.Module.__init__(inst)
    object.__setattr__(inst, "_hparams_shim", easydict.EasyDict(hparams))
    cls.hparams = property(lambda self: self._hparams_shim)
    for key, value in components.items():
        setattr(inst, key, value)
    inst.has_bayer_filter = False
    inst.render_bkgd = components.get("render_bkgd_mode")
    inst.train_ray_sample_batch_size = train_ray_sample_batch_size
    logged = {}

    class _Dataset:
        batch_size = None

    class _Sampler:
        datasets = []

    class _DataModule:
        train_dataset = _Dataset()
        train_normalized_sampler = _Sampler()

    class _Trainer:
        accumulate_grad_batches = 1
        datamodule = _DataModule()

    object.__setattr__(inst, "trainer", _Trainer())
    cls.global_step = 0
    inst.log = lambda name, value, **kw: logged.__setitem__(name, value)
    inst.all_gather = lambda t: t[None]
    inst.logged = logged
    return inst
