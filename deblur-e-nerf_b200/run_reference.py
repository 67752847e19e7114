"""Run the REFERENCE'S OWN `scripts/run.py` unchanged on top of the den_b200 operators:

    python -m deblur_e_nerf_b200.run_reference /path/to/deblur-e-nerf/scripts/run.py train cfg.yaml

registers the stand-ins of `compat/` for the packages that are not installed (`pytorch_lightning`,
`easydict`, `roma`, evaluation-only `pypose` / `torchmetrics` / `lpips`), puts the B1 drop-ins in
`sys.modules` as `nerfacc` / `tinycudann` and then executes the script with `runpy`
as `__main__`, with `sys.argv` and `sys.path[0]` set as `python scripts/run.py ...` would set them
(`scripts/run.py:13-15` derives the project directory from `sys.path[0]`)."""

import os
import runpy
import sys


def main(argv=None, operators=True):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    from .compat import install
    install(operators=operators)
    old_argv, old_path0 = sys.argv, sys.path[0]
    sys.argv = [script] + argv[1:]
    sys.path[0] = os.path.dirname(script)
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv, sys.path[0] = old_argv, old_path0


if __name__ == "__main__":
    main()
