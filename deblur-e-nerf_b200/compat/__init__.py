"""Stand-ins that let the REFERENCE'S OWN `scripts/run.py` and `deblur_e_nerf` package run unchanged on
top of the den_b200 operators: a `pytorch_lightning` 1.4.9-shaped package with the
part of the API `scripts/run.py:10,32,70-100` and the two Lightning classes of the reference use,
`easydict`, the handful of `roma` functions of `utils/tensor_ops.py` / `models/trajectories.py`, and
stand-ins for the evaluation-only dependencies (`torchmetrics.functional.psnr / ssim`: functional;
`lpips`: NaN with a warning; `pypose`: import-time only).

`install()` registers them in `sys.modules` under their upstream names — only where the real package
is not importable — together with the B1 drop-ins (`nerfacc`, `tinycudann` → deblur_e_nerf_b200);
`python -m deblur_e_nerf_b200.run_reference /path/to/reference/scripts/run.py train cfg.yaml` does that
and then runs the script with `runpy`."""

import importlib
import sys


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


def install(operators=True, force=False):
    """Register the stand-ins.  `operators=False` leaves `nerfacc` / `tinycudann` alone (a caller — the
    CPU test — has put its own there)."""
    from . import easydict as _easydict
    from . import lpips as _lpips
    from . import pypose as _pypose
    from . import pytorch_lightning as _pl
    from . import roma as _roma
    from . import torchmetrics as _torchmetrics
    table = {"easydict": _easydict, "roma": _roma, "pytorch_lightning": _pl, "pypose": _pypose,
             "torchmetrics": _torchmetrics, "lpips": _lpips}
    installed = []
    for name, module in table.items():
        if force or _missing(name):
            sys.modules[name] = module
            for sub in getattr(module, "SUBMODULES", ()):
                sys.modules[f"{name}.{sub}"] = getattr(module, sub.split(".")[0]) if "." not in sub \
                    else importlib.import_module(f"{module.__name__}.{sub}")
            installed.append(name)
    if operators:
        from .. import nerfacc, tinycudann
        sys.modules["nerfacc"] = nerfacc
        sys.modules["tinycudann"] = tinycudann
        installed += ["nerfacc", "tinycudann"]
    return installed
