"""Import-time stand-in for `pypose`: the reference only touches it in the evaluation post-processing
(`external/optimizer.py`, `models/deblur_e_nerf.py:846-871`), which deblur_e_nerf_b200.eval_post
replaces on the device.  The classes exist so that `class LevenbergMarquardt(pp.optim.LevenbergMarquardt)`
can be defined; using them raises."""

import types


class _Unavailable:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "pypose is not installed: the Levenberg-Marquardt refinement of the reference's evaluation "
            "runs through deblur_e_nerf_b200.eval_post.evaluate instead")


class GaussNewton(_Unavailable):
    pass


class LevenbergMarquardt(_Unavailable):
    pass


optim = types.SimpleNamespace(
    GaussNewton=GaussNewton, LevenbergMarquardt=LevenbergMarquardt,
    solver=types.SimpleNamespace(LSTSQ=_Unavailable, Cholesky=_Unavailable),
    strategy=types.SimpleNamespace(TrustRegion=_Unavailable),
    functional=types.SimpleNamespace(modjac=_Unavailable))
