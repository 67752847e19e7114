"""deblur-e-nerf_b200 — B200-native (sm_100a) event-supervised volumetric renderer,
the training hot path of Deblur e-NeRF behind the reference's own interfaces.

The directory name carries the upstream project's hyphens, so it is imported
through the alias package ``deblur_e_nerf_b200`` (a two-line ``__init__`` at the
repository root that points ``__path__`` here).

Sub-modules
-----------
``_build``      in-tree nvcc build of ``lib/libden_b200.so`` (C ABI, ``include/den_b200.h``)
``_lib``        ctypes binding; fails loudly when the library is missing (no fallback)
``ops``         tensor-level wrappers + autograd Functions over the C ABI
``nerfacc``     B1 drop-in for the seven nerfacc symbols the reference imports
``tinycudann``  B1 drop-in for ``tinycudann.Encoding`` (HashGrid, fp32)
``field``, ``nerf``, ``pixel_bandwidth``, ``trajectories``, ``event_generation_params``,
``loss``, ``renderer``   B2: host-side mirrors of the reference's hot-path classes
``ddp``         one-process-per-GPU data parallelism (flat-buffer NCCL all-reduce)
"""

__version__ = "0.1.0"
