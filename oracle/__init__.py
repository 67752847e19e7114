"""CPU oracle for the Deblur e-NeRF event-supervised volumetric renderer.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the timed CPU baseline), never as the thing shipped.

Contents
--------
``easydict_ref``, ``roma_ref``, ``nerfacc_ref``, ``tcnn_ref``
    pure-PyTorch (CPU, fp32/fp64) restatements of the four third-party
    packages the reference's hot-path files import but which are absent from
    ``/root/reference`` and from this image: ``easydict``, ``roma==1.2.7``,
    ``nerfacc==0.3.1`` and ``tiny-cuda-nn`` (``tinycudann.Encoding``,
    ``HashGrid``/``Linear``, fp32).  Each function cites the upstream file it
    restates and the reference call site that pins its use.
``ref_shim``
    registers those four under their upstream names and imports the
    reference's OWN hot-path files, unmodified, from ``/root/reference`` (only
    possible in the build container; used to validate ``path_ref`` and to
    generate ``tests/golden/*.npz``).
``path_ref``
    a restatement, in this repo's own words, of the reference's hot-path files
    (``models/nerf.py``, ``external/{utils,vol_rendering,ngp,mlp,sh_encoder}.py``,
    ``models/{pixel_bandwidth,trajectories,event_generation_params}.py``,
    ``utils/control.py``, ``loss_metric/loss.py`` and the hot lines of
    ``models/deblur_e_nerf.py``) so the oracle travels to the GPU box, where
    ``/root/reference`` does not exist.

``eval_ref``
    the evaluation post-processing (``evaluation_epoch_end``: gain-exposure normalisation,
    log-space affine fit, offset-gamma Levenberg-Marquardt refinement, L1 / PSNR / SSIM);
    pinned against the reference's own method and its ``OffsetGammaCorrection``; the
    trust-region constants (pypose) and SSIM (torchmetrics 0.6.2) are restated from the
    published sources, parity unpinned.
``events_ref``
    the raw-event preprocessing (``Event.queue_raw_events``, ``extract_max_refractory_period``,
    ``colorize_events``): the literal loops and a vectorised form, pinned bit for bit against
    the reference's own classmethods (live and ``tests/golden/raw_events.npz``).

Parity status: the reference ships no tests, golden vectors or fixtures and the
two kernel packages are absent, so the THIRD-PARTY restatements are "parity
unpinned" against upstream binaries (closed-form known-answer tests only);
``path_ref`` IS pinned, against the reference's own files run here under
``ref_shim`` (``tests/test_oracle_vs_reference.py`` and ``tests/golden``).
"""
