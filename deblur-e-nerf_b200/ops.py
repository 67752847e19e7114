"""Tensor-level wrappers over the den_b200 C ABI (``include/den_b200.h``).

Each function validates its tensors (CUDA, dtype, contiguity), allocates the
outputs with torch (plumbing only) and launches the hand-written kernels on the
current CUDA stream through ctypes.  ``torch.autograd.Function`` subclasses glue
the forward/backward kernel pairs into autograd.  There is no CPU path: CPU
tensors raise ``NotImplementedError`` like the upstream packages do
(``nerfacc/ray_marching.py``: "Only support cuda inputs.").
"""

import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import HashGridDesc, LpfLossDesc, MarchParams, OccGridDesc

_LAUNCHES = 0           # kernels launched through the C ABI (bench.py reports it)


def launch_count():
    return _LAUNCHES


def _count(n=1):
    global _LAUNCHES
    _LAUNCHES += n


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


_ROW_QUANTUM = 1 << 18      # sample-sized buffers are allocated in multiples of 262 144 rows


def _rows(n, tail, dtype, device):
    """(n, *tail) buffer for per-sample data.  The sample count changes a little from one render
    call to the next (stratified jitter, fresh batch); allocating the exact size makes the caching
    allocator miss on every slightly larger request (a cudaMalloc + implicit device sync in the
    middle of the step).  Rounding the ROW count up to a quantum and handing out the leading view
    keeps the block sizes constant from step to step."""
    n = int(n)
    cap = n if n < (1 << 16) else -(-n // _ROW_QUANTUM) * _ROW_QUANTUM
    buf = torch.empty((cap, *tail), dtype=dtype, device=device)
    return buf if cap == n else buf[:n]


def _rows_like(t):
    return _rows(t.shape[0], tuple(t.shape[1:]), t.dtype, t.device)


def _req(t, dtype, name):
    if not t.is_cuda:
        raise NotImplementedError(f"{name}: only CUDA tensors are supported (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


_TIME_ALL = False
_TIMED = None           # {kernel name: [(start_event, end_event), ...]} while timing is on
_EVENT_POOL = []        # pre-created CUDA events (creation is kept out of the timed region)


def enable_kernel_timing(names, pool_size=4096):
    """Bracket launches of the named C-ABI entry points (None = all of them) with CUDA events
    recorded on the launching (current) stream; bench.py derives the roofline from them.  The
    events come from a pool created here, so the timed region only pays two cudaEventRecord
    calls per bracketed launch."""
    global _TIMED, _TIME_ALL
    _TIME_ALL = names is None
    _TIMED = {n: [] for n in (names or [])}
    while len(_EVENT_POOL) < pool_size:
        _EVENT_POOL.append(torch.cuda.Event(enable_timing=True))


def _event():
    return _EVENT_POOL.pop() if _EVENT_POOL else torch.cuda.Event(enable_timing=True)


def kernel_timings():
    """{name: (launches, total_ms)} — call after a device synchronise."""
    out = {}
    for name, pairs in (_TIMED or {}).items():
        out[name] = (len(pairs), sum(a.elapsed_time(b) for a, b in pairs))
    return out


def disable_kernel_timing():
    global _TIMED
    for pairs in (_TIMED or {}).values():
        for a, b in pairs:
            _EVENT_POOL.extend((a, b))
    _TIMED = None


def _call(name, *args, launches=1):
    if _TIMED is not None and (_TIME_ALL or name in _TIMED):
        _TIMED.setdefault(name, [])
        start, end = _event(), _event()
        start.record()
        _lib.lib().call(name, *args)
        end.record()
        _TIMED[name].append((start, end))
    else:
        _lib.lib().call(name, *args)
    _count(launches)


# --------------------------------------------------------------------------- #
# hash grid
# --------------------------------------------------------------------------- #
def grid_level_table(n_levels, base_resolution, per_level_scale, log2_hashmap_size):
    """tcnn grid.h level table: (scale fp32, resolution, entries, offset) per level.

    ``scale = exp2(l * log2(s)) * base - 1`` in IEEE fp32; ``res = ceil(scale) + 1``;
    entries ``min(round_up(res^3, 8), 2^T)`` (SURVEY.md §8(a) level table).
    """
    log2_s = np.float32(np.log2(np.float32(per_level_scale)))
    scales, ress, sizes, offsets = [], [], [], []
    offset = 0
    for level in range(n_levels):
        scale = np.float32(np.exp2(np.float32(np.float32(level) * log2_s))
                           * np.float32(base_resolution) - np.float32(1.0))
        res = int(math.ceil(float(scale))) + 1
        cap = 0xFFFFFFFF // 2
        dense = cap if float(res) ** 3 > float(cap) else res ** 3
        dense = ((dense + 7) // 8) * 8
        size = min(dense, 1 << log2_hashmap_size)
        scales.append(float(scale))
        ress.append(res)
        sizes.append(size)
        offsets.append(offset)
        offset += size
    return scales, ress, sizes, offsets, offset


def make_hashgrid_desc(n_levels, base_resolution, per_level_scale, log2_hashmap_size,
                       n_features=2, agg_max_resolution=800):
    """`agg_max_resolution`: levels up to this resolution merge runs of consecutive samples in the same
    cell before the scatter's atomics.  Measured on the synthetic.yaml shape (10.2 M samples,
    profiles/time_hashgrid.py): 256 -> 3.22 ms, 512 -> 3.21, 800 -> 3.08, 1100 -> 3.09, 2100 -> 3.12 alone;
    inside the training step (41 M samples, survivors of the visibility filter) 512 and 800 both take 11.8 ms."""
    if n_levels > _lib.DEN_MAX_LEVELS:
        raise ValueError("too many levels")
    scales, ress, sizes, offsets, total = grid_level_table(
        n_levels, base_resolution, per_level_scale, log2_hashmap_size)
    d = HashGridDesc()
    d.n_levels = n_levels
    d.n_features = n_features
    d.n_agg_levels = sum(1 for r in ress if r <= agg_max_resolution)
    for i in range(n_levels):
        d.scale[i] = scales[i]
        d.resolution[i] = ress[i]
        d.size[i] = sizes[i]
        d.offset[i] = offsets[i]
    return d, total


def hashgrid_fwd(desc, x, table, n_dev=None):
    """`n_dev` (here and in every per-sample wrapper below): optional device int32 with the true row
    count; the tensors then hold `x.shape[0]` rows of CAPACITY and the kernel works on the first
    min(capacity, n_dev) of them — no host read-back of the count (see include/den_b200.h)."""
    x = _req(x, torch.float32, "x")
    table = _req(table, torch.float32, "table")
    n = x.shape[0]
    out = _rows(n, (desc.n_levels * desc.n_features,), torch.float32, x.device)
    _call("den_hashgrid_fwd", ctypes.byref(desc), _ptr(x), _ptr(table), _ptr(out), n, _ptr(n_dev),
          _stream())
    return out


def hashgrid_bwd(desc, x, dout, table, need_dx, dtable=None, n_dev=None):
    x = _req(x, torch.float32, "x")
    dout = _req(dout, torch.float32, "dout")
    n = x.shape[0]
    if dtable is None:
        dtable = torch.zeros_like(table)
    dx = (torch.zeros_like(x) if n_dev is not None else _rows_like(x)) if need_dx else None
    _call("den_hashgrid_bwd", ctypes.byref(desc), _ptr(x), _ptr(dout), _ptr(table), _ptr(dtable),
          _ptr(dx), n, _ptr(n_dev), _stream())
    return dtable, dx


class _HashGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, desc):
        ctx.desc = desc
        ctx.save_for_backward(x, table)
        return hashgrid_fwd(desc, x, table)

    @staticmethod
    def backward(ctx, dout):
        x, table = ctx.saved_tensors
        need_dx = ctx.needs_input_grad[0]
        dtable, dx = hashgrid_bwd(ctx.desc, x, dout, table, need_dx)
        return dx, (dtable if ctx.needs_input_grad[1] else None), None


def hashgrid(x, table, desc):
    return _HashGridFn.apply(x, table, desc)


class _HashGridReuseFn(torch.autograd.Function):
    """Autograd node for an encoding that was ALREADY computed (the visibility pre-pass ran the
    gather on the same positions): forward hands it back, backward is the usual scatter."""

    @staticmethod
    def forward(ctx, x, table, desc, enc, n_dev):
        ctx.desc = desc
        ctx.n_dev = n_dev
        ctx.save_for_backward(x, table)
        return enc.view_as(enc)

    @staticmethod
    def backward(ctx, dout):
        x, table = ctx.saved_tensors
        dtable, dx = hashgrid_bwd(ctx.desc, x, dout, table, ctx.needs_input_grad[0], n_dev=ctx.n_dev)
        return dx, (dtable if ctx.needs_input_grad[1] else None), None, None, None


def hashgrid_reuse(x, table, desc, enc, n_dev=None):
    """With `enc_rows` (field.mlp_samples) `enc` may be the UN-compacted pre-pass encodings while `x`
    holds the survivors' positions: the node then stands for the survivors' rows, which only
    den_mlp_bwd reads (through the row map); both have the same capacity."""
    assert enc.shape[0] == x.shape[0]
    return _HashGridReuseFn.apply(x, table, desc, enc, n_dev)


# --------------------------------------------------------------------------- #
# scan / marching / visibility
# --------------------------------------------------------------------------- #
def exclusive_scan_i32(counts):
    """(n,) int32 -> (n+1,) int32 exclusive prefix, last element = total."""
    counts = _req(counts, torch.int32, "counts")
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int32, device=counts.device)
    ws_bytes = _lib.lib().raw("den_scan_workspace_bytes")(n)
    ws = torch.empty(max(ws_bytes // 4, 1), dtype=torch.int32, device=counts.device)
    _call("den_exclusive_scan_i32", _ptr(counts), _ptr(out), n, _ptr(ws), ws.numel() * 4,
          _stream(), launches=3 if n else 0)
    return out


def make_march_params(roi, res, contraction, step_size, cone_angle):
    p = MarchParams()
    for i in range(6):
        p.roi[i] = float(roi[i])
    for i in range(3):
        p.res[i] = int(res[i])
    p.contraction = int(contraction)
    p.step_size = float(np.float32(step_size))
    p.cone_angle = float(np.float32(cone_angle))
    return p


def ray_aabb_intersect(rays_o, rays_d, aabb):
    rays_o = _req(rays_o, torch.float32, "rays_o")
    rays_d = _req(rays_d, torch.float32, "rays_d")
    n = rays_o.shape[0]
    t_min = torch.empty(n, dtype=torch.float32, device=rays_o.device)
    t_max = torch.empty_like(t_min)
    host = (ctypes.c_float * 6)(*[float(v) for v in aabb])
    _call("den_ray_aabb_intersect", _ptr(rays_o), _ptr(rays_d), host, _ptr(t_min), _ptr(t_max), n,
          _stream())
    return t_min, t_max


def clamp_jitter_(t_min, t_max, jitter, near_plane, far_plane, step_size):
    n = t_min.numel()
    _call("den_clamp_jitter", _ptr(t_min), _ptr(t_max), _ptr(jitter),
          int(near_plane is not None), float(near_plane or 0.0),
          int(far_plane is not None), float(far_plane or 0.0), float(np.float32(step_size)), n,
          _stream())


MARCH_MAX_PER_RAY = 8192            # bound-kernel path: a ray whose bound exceeds this takes two passes
MARCH_ARENA_MAX = (1 << 31) - 1     # uniform-segment arena limit: int32 segment offsets (2 x 4 B per entry)
_ARENA = {}                         # device -> (t0, t1) grow-only scratch arena
_SEG_OFFSETS = {}                   # (device, n, seg_len) -> int32 arange(n + 1) * seg_len


def _arena(dev, n):
    cur = _ARENA.get(dev)
    if cur is None or cur[0].numel() < n:
        cur = (torch.empty(n, dtype=torch.float32, device=dev),
               torch.empty(n, dtype=torch.float32, device=dev))
        _ARENA[dev] = cur
    return cur


def march_segment_length(t_lo, t_hi, step_size):
    """Most samples a ray can emit between distances t_lo and t_hi (host arithmetic, + slack)."""
    if t_lo is None or t_hi is None or not math.isfinite(t_hi - t_lo):
        return None
    return int(max(t_hi - t_lo, 0.0) / float(step_size)) + 4


def march(params, rays_o, rays_d, t_min, t_max, binary, capacity=None, seg_len=None,
          single_pass=True, probe=None, overflow=None):
    """Occupancy march.  Returns (ray_indices i32, t_starts, t_ends, offsets (R+1)).

    Default: ONE marching pass into an upper-bound arena, then a coalesced pack — the sequential
    marching loop, the expensive part, runs once instead of twice.  With ``seg_len`` (a host-known
    bound of the samples per ray, e.g. (far - near) / step) every ray owns a fixed segment of a
    cached scratch arena and the only host read is the sample total; without it the per-ray bounds
    come from ``den_march_bound`` (one more host read for the arena size).
    ``single_pass=False`` (or a caller-given ``capacity``) selects the count + write scheme: with
    ``capacity=None`` the total is read back (one host sync, as upstream does) and the outputs are
    exact-size; with a capacity the arena is caller-bounded, nothing is synchronised and
    ``offsets[-1]`` (device) holds the true total.

    ``probe`` (int64 index tensor into ``offsets``): those entries ride along with the host read of
    the total and come back as a fifth return value (a list of ints) — per-group sample counts
    without a second synchronisation.

    ``overflow`` (device int32 scalar) selects the SYNC-FREE single pass: `capacity` rows are
    allocated from the caller's estimate, the offsets are clamped to it (``overflow`` is set if
    samples were lost) and NOTHING is read back — the true total stays on the device in
    ``offsets[-1]`` for the ``n_dev`` argument of the per-sample kernels.  Needs ``seg_len``.
    """
    rays_o = _req(rays_o, torch.float32, "rays_o")
    rays_d = _req(rays_d, torch.float32, "rays_d")
    t_min = _req(t_min, torch.float32, "t_min")
    t_max = _req(t_max, torch.float32, "t_max")
    if binary.dtype == torch.bool:
        binary = binary.contiguous().view(torch.uint8)
    binary = _req(binary, torch.uint8, "binary")
    n = rays_o.shape[0]
    dev = rays_o.device
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    if overflow is not None:
        if capacity is None or seg_len is None or n == 0 or n * seg_len > MARCH_ARENA_MAX:
            raise ValueError("sync-free march needs a capacity and a per-ray segment bound")
        key = (dev, n, seg_len)
        seg = _SEG_OFFSETS.get(key)
        if seg is None:
            seg = (torch.arange(n + 1, dtype=torch.int64, device=dev) * seg_len).to(torch.int32)
            if len(_SEG_OFFSETS) > 16:
                _SEG_OFFSETS.clear()
            _SEG_OFFSETS[key] = seg
        arena_t0, arena_t1 = _arena(dev, n * seg_len)
        _call("den_march_single", ctypes.byref(params), _ptr(rays_o), _ptr(rays_d), _ptr(t_min),
              _ptr(t_max), _ptr(binary), _ptr(seg), _ptr(counts), _ptr(arena_t0), _ptr(arena_t1), n,
              _stream())
        offsets = exclusive_scan_i32(counts)
        clamp_offsets_(offsets, capacity, overflow)
        ray_indices = _rows(capacity, (), torch.int32, dev)
        t_starts = _rows(capacity, (), torch.float32, dev)
        t_ends = _rows(capacity, (), torch.float32, dev)
        _call("den_march_pack", _ptr(seg), _ptr(offsets), _ptr(arena_t0), _ptr(arena_t1), n,
              _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends), _stream())
        return ray_indices, t_starts, t_ends, offsets
    if single_pass and capacity is None and n > 0:
        if seg_len is not None and n * seg_len <= MARCH_ARENA_MAX:
            key = (dev, n, seg_len)
            seg = _SEG_OFFSETS.get(key)
            if seg is None:
                seg = (torch.arange(n + 1, dtype=torch.int64, device=dev) * seg_len).to(torch.int32)
                if len(_SEG_OFFSETS) > 16:
                    _SEG_OFFSETS.clear()
                _SEG_OFFSETS[key] = seg
            bound = None
            arena_t0, arena_t1 = _arena(dev, n * seg_len)
        else:
            bound = torch.empty(n, dtype=torch.int32, device=dev)
            _call("den_march_bound", ctypes.byref(params), _ptr(t_min), _ptr(t_max),
                  MARCH_MAX_PER_RAY, _ptr(bound), n, _stream())
            seg = exclusive_scan_i32(bound)
            arena_t0, arena_t1 = _arena(dev, max(int(seg[-1].item()), 1))    # host read: arena size
        _call("den_march_single", ctypes.byref(params), _ptr(rays_o), _ptr(rays_d), _ptr(t_min),
              _ptr(t_max), _ptr(binary), _ptr(seg), _ptr(counts), _ptr(arena_t0), _ptr(arena_t1), n,
              _stream())
        offsets = exclusive_scan_i32(counts)
        over = (counts > (seg_len if bound is None else bound)).any().to(torch.int32)
        head = torch.stack((offsets[-1], over))
        if probe is not None:
            head = torch.cat((head, offsets[probe]))
        vals = [int(v) for v in head.tolist()]                                        # host read
        total, overflow, probed = vals[0], vals[1], vals[2:]
        if not overflow:
            ray_indices = _rows(total, (), torch.int32, dev)
            t_starts = _rows(total, (), torch.float32, dev)
            t_ends = _rows(total, (), torch.float32, dev)
            _call("den_march_pack", _ptr(seg), _ptr(offsets), _ptr(arena_t0), _ptr(arena_t1), n,
                  _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends), _stream())
            out = (ray_indices, t_starts, t_ends, offsets)
            return out if probe is None else out + (probed,)
        capacity = total            # a ray outgrew its segment: exact write pass instead
    else:
        _call("den_march_count", ctypes.byref(params), _ptr(rays_o), _ptr(rays_d), _ptr(t_min),
              _ptr(t_max), _ptr(binary), _ptr(counts), n, _stream())
        offsets = exclusive_scan_i32(counts)
        if capacity is None:
            capacity = int(offsets[-1].item())
    ray_indices = _rows(capacity, (), torch.int32, dev)
    t_starts = _rows(capacity, (), torch.float32, dev)
    t_ends = _rows(capacity, (), torch.float32, dev)
    _call("den_march_write", ctypes.byref(params), _ptr(rays_o), _ptr(rays_d), _ptr(t_min),
          _ptr(t_max), _ptr(binary), _ptr(offsets), _ptr(ray_indices), _ptr(t_starts),
          _ptr(t_ends), n, capacity, _stream())
    out = (ray_indices, t_starts, t_ends, offsets)
    return out if probe is None else out + ([int(v) for v in offsets[probe].tolist()],)


def alpha_from_sigma(sigmas, t_starts, t_ends, n_dev=None):
    sigmas = _req(sigmas.reshape(-1), torch.float32, "sigmas")
    out = _rows_like(sigmas)
    _call("den_alpha_from_sigma", _ptr(sigmas), _ptr(t_starts), _ptr(t_ends), _ptr(out),
          sigmas.numel(), _ptr(n_dev), _stream())
    return out


def visibility(alphas, offsets, early_stop_eps, alpha_thre):
    alphas = _req(alphas.reshape(-1), torch.float32, "alphas")
    offsets = _req(offsets, torch.int32, "offsets")
    n_rays = offsets.numel() - 1
    mask = _rows(alphas.numel(), (), torch.uint8, alphas.device)
    counts = torch.empty(n_rays, dtype=torch.int32, device=alphas.device)
    _call("den_visibility", _ptr(alphas), _ptr(offsets), n_rays, float(early_stop_eps),
          float(alpha_thre), _ptr(mask), _ptr(counts), _stream())
    return mask, counts


def compact(mask, offsets_in, offsets_out, ray_indices, t_starts, t_ends, capacity):
    dev = mask.device
    n_rays = offsets_in.numel() - 1
    ro = _rows(capacity, (), torch.int32, dev)
    t0 = _rows(capacity, (), torch.float32, dev)
    t1 = _rows(capacity, (), torch.float32, dev)
    _call("den_compact_samples", _ptr(mask), _ptr(offsets_in), _ptr(offsets_out),
          _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends), _ptr(ro), _ptr(t0), _ptr(t1), n_rays,
          _stream())
    return ro, t0, t1


def compact_ex(mask, offsets_in, offsets_out, ray_indices, t_starts, t_ends, capacity, sigmas, rgbs):
    """Compaction that also carries the survivors' pre-pass sigma / rgb along and records, per
    survivor, its row in the un-compacted arrays.  Capacity-sized outputs; no host read."""
    dev = mask.device
    n_rays = offsets_in.numel() - 1
    ro = _rows(capacity, (), torch.int32, dev)
    t0 = _rows(capacity, (), torch.float32, dev)
    t1 = _rows(capacity, (), torch.float32, dev)
    sig = _rows(capacity, (), torch.float32, dev)
    channels = rgbs.shape[-1]
    rgb = _rows(capacity, (channels,), torch.float32, dev)
    rows = _rows(capacity, (), torch.int32, dev)
    _call("den_compact_samples_ex", _ptr(mask), _ptr(offsets_in), _ptr(offsets_out),
          _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends), _ptr(ro), _ptr(t0), _ptr(t1), n_rays,
          _ptr(sigmas), _ptr(rgbs), channels, _ptr(sig), _ptr(rgb), _ptr(rows), _stream())
    return ro, t0, t1, sig, rgb, rows


def clamp_offsets_(offsets, capacity, overflow):
    """offsets <- min(offsets, capacity) in place; `overflow` (device int32) is set when the total
    exceeded the capacity."""
    _call("den_clamp_offsets", _ptr(offsets), offsets.numel(), int(capacity), _ptr(overflow), _stream())
    return offsets


def offsets_from_ray_indices(ray_indices, n_rays):
    """(R+1) int32 packing offsets of a sorted ``ray_indices`` (B1 callers pass indices)."""
    counts = torch.bincount(ray_indices.long(), minlength=n_rays).to(torch.int32)
    return exclusive_scan_i32(counts)


# --------------------------------------------------------------------------- #
# occupancy-grid update
# --------------------------------------------------------------------------- #
def make_occgrid_desc(roi, res, contraction):
    d = OccGridDesc()
    for i in range(6):
        d.roi[i] = float(roi[i])
    for i in range(3):
        d.res[i] = int(res[i])
    d.contraction = int(contraction)
    return d


def occgrid_cell_points(desc, indices, jitter, want_keep):
    """Cell indices (n,) int64 or None (= every cell in order) + jitter (n,3) -> world points (n,3)
    and, if asked, the mask of the points upstream keeps (sphere contraction: inside the unit ball)."""
    jitter = _req(jitter, torch.float32, "jitter")
    n = jitter.shape[0]
    if indices is not None:
        indices = _req(indices, torch.int64, "indices")
        assert indices.numel() == n
    world = torch.empty((n, 3), dtype=torch.float32, device=jitter.device)
    keep = torch.empty(n, dtype=torch.uint8, device=jitter.device) if want_keep else None
    _call("den_occgrid_cell_points", ctypes.byref(desc), _ptr(indices), _ptr(jitter), n, _ptr(world),
          _ptr(keep), _stream())
    return world, keep


def occgrid_occ(sigma, world, camera_ids, camera_pos, cone_angle, step_size, near_plane, far_plane):
    """density -> density * step size of the cell (models/nerf.py:175-198)."""
    sigma = _req(sigma.reshape(-1), torch.float32, "sigma")
    n = sigma.numel()
    occ = torch.empty_like(sigma)
    has_planes = near_plane is not None and far_plane is not None
    if cone_angle > 0.0:
        world = _req(world, torch.float32, "world")
        camera_ids = _req(camera_ids, torch.int64, "camera_ids")
        camera_pos = _req(camera_pos, torch.float32, "camera_pos")
    _call("den_occgrid_occ", _ptr(sigma), _ptr(world) if cone_angle > 0.0 else None,
          _ptr(camera_ids) if cone_angle > 0.0 else None,
          _ptr(camera_pos) if cone_angle > 0.0 else None, float(cone_angle),
          float(np.float32(step_size)), int(has_planes), float(near_plane or 0.0),
          float(far_plane or 0.0), n, _ptr(occ), _stream())
    return occ


_OCC_WS = {}        # (device, n_cells) -> initialised workspace


def occgrid_ema_update(indices, occ, occs, binary_u8, ema_decay, occ_thre):
    """occs[indices] = max(occs[indices] * decay, occ); binary = occs > min(mean(occs), thre).
    Returns mean(occs) as a device scalar."""
    occ = _req(occ.reshape(-1), torch.float32, "occ")
    occs = _req(occs, torch.float32, "occs")
    n_cells = occs.numel()
    if indices is not None:
        indices = _req(indices, torch.int64, "indices")
        assert indices.numel() == occ.numel()
    key = (occs.device, n_cells)
    ws = _OCC_WS.get(key)
    if ws is None:
        nbytes = _lib.lib().raw("den_occgrid_workspace_bytes")(n_cells)
        ws = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=occs.device)
        _call("den_occgrid_workspace_init", _ptr(ws), n_cells, _stream())
        _OCC_WS[key] = ws
    mean = torch.empty((), dtype=torch.float32, device=occs.device)
    _call("den_occgrid_ema_update", _ptr(indices), None, _ptr(occ), occ.numel(), float(ema_decay),
          float(np.float32(occ_thre)), _ptr(occs), n_cells, _ptr(binary_u8), _ptr(mean), _ptr(ws),
          _stream(), launches=3)
    return mean


# --------------------------------------------------------------------------- #
# weights / accumulation / fused compositor
# --------------------------------------------------------------------------- #
class _WeightFromDensityFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigmas, t_starts, t_ends, offsets):
        sigmas = _req(sigmas.reshape(-1), torch.float32, "sigmas")
        t_starts = _req(t_starts.reshape(-1), torch.float32, "t_starts")
        t_ends = _req(t_ends.reshape(-1), torch.float32, "t_ends")
        n_rays = offsets.numel() - 1
        w = _rows_like(sigmas)
        _call("den_weight_from_density_fwd", _ptr(sigmas), _ptr(t_starts), _ptr(t_ends),
              _ptr(offsets), n_rays, _ptr(w), _stream())
        ctx.save_for_backward(sigmas, t_starts, t_ends, offsets)
        return w

    @staticmethod
    def backward(ctx, dw):
        sigmas, t_starts, t_ends, offsets = ctx.saved_tensors
        dw = _req(dw.reshape(-1), torch.float32, "dweights")
        ds = _rows_like(sigmas)
        _call("den_weight_from_density_bwd", _ptr(sigmas), _ptr(t_starts), _ptr(t_ends),
              _ptr(offsets), offsets.numel() - 1, _ptr(dw), _ptr(ds), _stream())
        return ds, None, None, None


class _WeightFromAlphaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alphas, offsets):
        alphas = _req(alphas.reshape(-1), torch.float32, "alphas")
        w = _rows_like(alphas)
        _call("den_weight_from_alpha_fwd", _ptr(alphas), _ptr(offsets), offsets.numel() - 1,
              _ptr(w), _stream())
        ctx.save_for_backward(alphas, offsets)
        return w

    @staticmethod
    def backward(ctx, dw):
        alphas, offsets = ctx.saved_tensors
        dw = _req(dw.reshape(-1), torch.float32, "dweights")
        da = _rows_like(alphas)
        _call("den_weight_from_alpha_bwd", _ptr(alphas), _ptr(offsets), offsets.numel() - 1,
              _ptr(dw), _ptr(da), _stream())
        return da, None


class _AccumulateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, values, ray_indices, offsets):
        weights = _req(weights.reshape(-1), torch.float32, "weights")
        dim = 1
        if values is not None:
            values = _req(values, torch.float32, "values")
            dim = values.shape[-1]
        n_rays = offsets.numel() - 1
        out = torch.empty((n_rays, dim), dtype=torch.float32, device=weights.device)
        _call("den_accumulate_fwd", _ptr(weights), _ptr(values), _ptr(offsets), n_rays, dim,
              _ptr(out), _stream())
        ctx.dim = dim
        ctx.has_values = values is not None
        ctx.save_for_backward(weights, values if values is not None else weights, ray_indices)
        return out

    @staticmethod
    def backward(ctx, dout):
        weights, values, ray_indices = ctx.saved_tensors
        if not ctx.has_values:
            values = None
        dout = _req(dout, torch.float32, "dout")
        n = weights.numel()
        dw = _rows_like(weights) if ctx.needs_input_grad[0] else None
        dv = (_rows_like(values)
              if (values is not None and ctx.needs_input_grad[1]) else None)
        _call("den_accumulate_bwd", _ptr(weights), _ptr(values), _ptr(ray_indices), _ptr(dout), n,
              ctx.dim, _ptr(dw), _ptr(dv), _stream())
        return dw, dv, None, None


class _CompositeFn(torch.autograd.Function):
    """Fused weights + colour/opacity/depth accumulation + background blend."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, t_starts, t_ends, offsets, bkgd):
        sigmas = _req(sigmas.reshape(-1), torch.float32, "sigmas")
        rgbs = _req(rgbs, torch.float32, "rgbs")
        channels = rgbs.shape[-1]
        n_rays = offsets.numel() - 1
        dev = sigmas.device
        colour = torch.empty((n_rays, channels), dtype=torch.float32, device=dev)
        opacity = torch.empty(n_rays, dtype=torch.float32, device=dev)
        depth = torch.empty(n_rays, dtype=torch.float32, device=dev)
        bk = None if bkgd is None else _req(bkgd.detach(), torch.float32, "bkgd")
        _call("den_composite_fwd", _ptr(sigmas), _ptr(rgbs), _ptr(t_starts), _ptr(t_ends),
              _ptr(offsets), n_rays, channels, _ptr(bk), _ptr(colour), _ptr(opacity), _ptr(depth),
              _stream())
        ctx.channels = channels
        ctx.has_bkgd = bk is not None
        ctx.save_for_backward(sigmas, rgbs, t_starts, t_ends, offsets,
                              bk if bk is not None else opacity, opacity, colour, depth)
        return colour, opacity, depth

    @staticmethod
    def backward(ctx, d_colour, d_opacity, d_depth):
        sigmas, rgbs, t_starts, t_ends, offsets, bk, opacity, colour, depth = ctx.saved_tensors
        if not ctx.has_bkgd:
            bk = None
        n_rays = offsets.numel() - 1
        d_colour = _req(d_colour, torch.float32, "d_colour")
        d_opacity = _req(d_opacity, torch.float32, "d_opacity")
        d_depth = _req(d_depth, torch.float32, "d_depth")
        d_sigmas = _rows_like(sigmas)
        d_rgbs = _rows_like(rgbs)
        d_bk = torch.zeros_like(bk) if (bk is not None and ctx.needs_input_grad[5]) else None
        _call("den_composite_bwd", _ptr(sigmas), _ptr(rgbs), _ptr(t_starts), _ptr(t_ends),
              _ptr(offsets), n_rays, ctx.channels, _ptr(bk), _ptr(colour), _ptr(opacity), _ptr(depth),
              _ptr(d_colour),
              _ptr(d_opacity), _ptr(d_depth), _ptr(d_sigmas), _ptr(d_rgbs), _ptr(d_bk), _stream())
        return d_sigmas, d_rgbs, None, None, None, d_bk


def weight_from_density(sigmas, t_starts, t_ends, offsets):
    """(M,) weights; inputs of any (M,)/(M,1) shape."""
    return _WeightFromDensityFn.apply(sigmas.reshape(-1), t_starts.reshape(-1),
                                      t_ends.reshape(-1), offsets)


def weight_from_alpha(alphas, offsets):
    return _WeightFromAlphaFn.apply(alphas.reshape(-1), offsets)


def accumulate(weights, values, ray_indices, offsets):
    return _AccumulateFn.apply(weights.reshape(-1), values, ray_indices, offsets)


def composite(sigmas, rgbs, t_starts, t_ends, offsets, bkgd=None):
    """Fused `rendering()`: (colour (R,C), opacity (R,), depth (R,))."""
    t_starts = _req(t_starts.detach().reshape(-1), torch.float32, "t_starts")
    t_ends = _req(t_ends.detach().reshape(-1), torch.float32, "t_ends")
    return _CompositeFn.apply(sigmas.reshape(-1), rgbs, t_starts, t_ends, offsets, bkgd)


# --------------------------------------------------------------------------- #
# fused field
# --------------------------------------------------------------------------- #
def field_density_at(desc, params, positions):
    positions = _req(positions, torch.float32, "positions")
    n = positions.shape[0]
    sig = torch.empty(n, dtype=torch.float32, device=positions.device)
    _call("den_field_density_at", ctypes.byref(desc), ctypes.byref(params), _ptr(positions), n,
          _ptr(sig), _stream())
    return sig


def field_fwd(desc, params, rays_o, rays_d, ray_indices, t_starts, t_ends, channels, n_dev=None):
    """channels == 0 -> density only."""
    rays_o = _req(rays_o, torch.float32, "rays_o")
    rays_d = _req(rays_d, torch.float32, "rays_d")
    ray_indices = _req(ray_indices, torch.int32, "ray_indices")
    t_starts = _req(t_starts.reshape(-1), torch.float32, "t_starts")
    t_ends = _req(t_ends.reshape(-1), torch.float32, "t_ends")
    n = ray_indices.numel()
    dev = rays_o.device
    sig = _rows(n, (), torch.float32, dev)
    rgb = _rows(n, (channels,), torch.float32, dev) if channels else None
    _call("den_field_fwd", ctypes.byref(desc), ctypes.byref(params), _ptr(rays_o), _ptr(rays_d),
          _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends), n, _ptr(n_dev), _ptr(sig), _ptr(rgb),
          _stream())
    return sig, rgb


# --------------------------------------------------------------------------- #
# tensor-core MLP on pre-encoded samples
# --------------------------------------------------------------------------- #
def contract_samples(desc, rays_o, rays_d, ray_indices, t_starts, t_ends, n_dev=None):
    n = ray_indices.numel()
    out = _rows(n, (3,), torch.float32, rays_o.device)
    _call("den_contract_samples", ctypes.byref(desc), _ptr(rays_o), _ptr(rays_d),
          _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends), n, _ptr(n_dev), _ptr(out), _stream())
    return out


def mlp_fwd(desc, params, enc, rays_o, rays_d, ray_indices, t_starts, t_ends, channels, n_dev=None):
    enc = _req(enc, torch.float32, "enc")
    n = ray_indices.numel()
    dev = enc.device
    sig = _rows(n, (), torch.float32, dev)
    rgb = _rows(n, (channels,), torch.float32, dev) if channels else None
    _call("den_mlp_fwd", ctypes.byref(desc), ctypes.byref(params), _ptr(enc), _ptr(rays_o),
          _ptr(rays_d), _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends), n, _ptr(n_dev), _ptr(sig),
          _ptr(rgb), _stream())
    return sig, rgb


def mlp_bwd(desc, params, grads_struct, enc, rays_o, rays_d, ray_indices, t_starts, t_ends,
            d_sigmas, d_rgbs, need_d_dirs=False, n_dev=None, enc_rows=None):
    """dL/denc (M, L*2) [and dL/d(view dir) (M,3)]; weight gradients are accumulated into the
    buffers of `grads_struct`.  `enc_rows` (M) int32: sample i reads row enc_rows[i] of `enc`."""
    n = ray_indices.numel()
    d_enc = _rows(n, (enc.shape[1],), torch.float32, enc.device)
    d_dirs = torch.zeros((n, 3), dtype=torch.float32, device=enc.device) if need_d_dirs else None
    d_sigmas = _req(d_sigmas.reshape(-1), torch.float32, "d_sigmas")
    d_rgbs = _req(d_rgbs, torch.float32, "d_rgbs")
    _call("den_mlp_bwd", ctypes.byref(desc), ctypes.byref(params), ctypes.byref(grads_struct),
          _ptr(enc), _ptr(rays_o), _ptr(rays_d), _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends),
          _ptr(d_sigmas), _ptr(d_rgbs), n, _ptr(n_dev), _ptr(enc_rows), _ptr(d_enc), _ptr(d_dirs),
          _stream())
    return d_enc, d_dirs


def segment_sum(values, offsets):
    """(M,D) per-sample vectors -> (R,D) per-ray sums (den_accumulate_fwd with unit weights)."""
    values = _req(values, torch.float32, "values")
    n_rays = offsets.numel() - 1
    out = torch.empty((n_rays, values.shape[-1]), dtype=torch.float32, device=values.device)
    _call("den_accumulate_fwd", None, _ptr(values), _ptr(offsets), n_rays, values.shape[-1],
          _ptr(out), _stream())
    return out


class _ContractSamplesFn(torch.autograd.Function):
    """Unit-cube sample positions as a function of the rays (differentiable in rays_o / rays_d:
    the refractory-period gradient path)."""

    @staticmethod
    def forward(ctx, desc, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets, n_dev):
        ctx.desc = desc
        ctx.n_dev = n_dev
        ctx.save_for_backward(rays_o, rays_d, ray_indices, t_starts, t_ends, offsets)
        return contract_samples(desc, rays_o, rays_d, ray_indices, t_starts, t_ends, n_dev)

    @staticmethod
    def backward(ctx, d_unit):
        rays_o, rays_d, ray_indices, t_starts, t_ends, offsets = ctx.saved_tensors
        n = ray_indices.numel()
        d_unit = _req(d_unit, torch.float32, "d_unit")
        d_pos = _rows_like(d_unit)
        d_pos_t = _rows_like(d_unit)
        _call("den_contract_samples_bwd", ctypes.byref(ctx.desc), _ptr(rays_o), _ptr(rays_d),
              _ptr(ray_indices), _ptr(t_starts), _ptr(t_ends), _ptr(d_unit), n, _ptr(ctx.n_dev),
              _ptr(d_pos), _ptr(d_pos_t), _stream())
        return (None, segment_sum(d_pos, offsets), segment_sum(d_pos_t, offsets), None, None, None, None,
                None)


def contract_samples_autograd(desc, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets, n_dev=None):
    return _ContractSamplesFn.apply(desc, rays_o, rays_d, ray_indices, t_starts, t_ends, offsets, n_dev)


# --------------------------------------------------------------------------- #
# optimiser
# --------------------------------------------------------------------------- #
def adam_step(tensor_array, n_tensors, beta1, beta2, eps, step, grad_scale=1.0, step_dev=None,
              skip_flag=None):
    """One Adam step over a ctypes array of ``AdamTensor`` descriptors (see optim.FusedAdam).
    `step_dev`: device int64 step counter (CUDA-graph capture) instead of the host `step`;
    `skip_flag`: device int32, non-zero -> the update is skipped."""
    _call("den_adam_step", tensor_array, int(n_tensors), float(beta1), float(beta2), float(eps),
          int(step), _ptr(step_dev), float(grad_scale), _ptr(skip_flag), _stream(), launches=2)


# --------------------------------------------------------------------------- #
# trajectory + camera model -> rays
# --------------------------------------------------------------------------- #
class _RaysFn(torch.autograd.Function):
    """den_rays_from_trajectory with its reverse mode w.r.t. the timestamps (den_rays_from_trajectory_bwd)."""

    @staticmethod
    def forward(ctx, timestamps, pixels, pose_ts, pose_pos, pose_quat, kinv_host):
        o, d = rays_from_trajectory(timestamps.detach(), pixels, pose_ts, pose_pos, pose_quat, kinv_host)
        ctx.save_for_backward(timestamps.detach(), pose_ts, pose_pos, pose_quat, d)
        return o, d

    @staticmethod
    def backward(ctx, d_o, d_d):
        timestamps, pose_ts, pose_pos, pose_quat, d = ctx.saved_tensors
        ts = _req(timestamps.reshape(-1), torch.float64, "timestamps")
        n = ts.numel()
        g_o = None if d_o is None else _req(d_o.reshape(-1, 3), torch.float32, "d_rays_o")
        g_d = None if d_d is None else _req(d_d.reshape(-1, 3), torch.float32, "d_rays_d")
        out = torch.empty(n, dtype=torch.float64, device=ts.device)
        _call("den_rays_from_trajectory_bwd", _ptr(ts), _ptr(_req(pose_ts, torch.int64, "pose_ts")),
              _ptr(_req(pose_pos, torch.float32, "pose_pos")), _ptr(_req(pose_quat, torch.float32, "pose_quat")),
              pose_ts.numel(), _ptr(d.reshape(-1, 3)), _ptr(g_o), _ptr(g_d), _ptr(out), n, _stream())
        return out.view(timestamps.shape).to(timestamps.dtype), None, None, None, None, None


def rays_from_trajectory_grad(timestamps, pixels, pose_ts, pose_pos, pose_quat, kinv_host):
    """rays_from_trajectory for timestamps that carry a gradient (the refractory-period path)."""
    return _RaysFn.apply(timestamps, pixels, pose_ts, pose_pos, pose_quat, kinv_host)


def rays_from_trajectory(timestamps, pixels, pose_ts, pose_pos, pose_quat, kinv_host):
    """timestamps (..., N) f64 ns, pixels (N, 2) f32 -> rays_o, rays_d (..., N, 3) in one launch
    (no gradient; `rays_from_trajectory_grad` is the differentiable form).
    `kinv_host`: the 9 floats of K^-1, row-major, on the host."""
    shape = tuple(timestamps.shape)
    ts = _req(timestamps.detach().reshape(-1), torch.float64, "timestamps")
    pixels = _req(pixels.reshape(-1, 2), torch.float32, "pixels")
    pose_ts = _req(pose_ts, torch.int64, "pose_ts")
    pose_pos = _req(pose_pos, torch.float32, "pose_pos")
    pose_quat = _req(pose_quat, torch.float32, "pose_quat")
    n = ts.numel()
    assert pixels.shape[0] > 0 and n % pixels.shape[0] == 0 and shape[-1] == pixels.shape[0]
    o = torch.empty((n, 3), dtype=torch.float32, device=ts.device)
    d = torch.empty((n, 3), dtype=torch.float32, device=ts.device)
    host = (ctypes.c_float * 9)(*[float(v) for v in kinv_host])
    _call("den_rays_from_trajectory", _ptr(ts), _ptr(pixels), pixels.shape[0], _ptr(pose_ts),
          _ptr(pose_pos), _ptr(pose_quat), pose_ts.numel(), host, _ptr(o), _ptr(d), n, _stream())
    return o.view(*shape, 3), d.view(*shape, 3)


# --------------------------------------------------------------------------- #
# pixel-bandwidth low-pass filter
# --------------------------------------------------------------------------- #
class _LpfFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, intensity, sample_dt, coef, n_channels):
        intensity = _req(intensity, torch.float32, "intensity")
        sample_dt = _req(sample_dt, torch.float32, "sample_dt")
        coef = _req(coef, torch.float64, "coef")
        S, N = intensity.shape[0], intensity[0].numel()
        out = torch.empty((N, n_channels), dtype=torch.float32, device=intensity.device)
        _call("den_lpf_fwd", _ptr(intensity), _ptr(sample_dt), _ptr(coef), S, N, n_channels,
              _ptr(out), _stream())
        ctx.n_channels = n_channels
        ctx.save_for_backward(intensity, sample_dt, coef)
        return out

    @staticmethod
    def backward(ctx, d_out):
        intensity, sample_dt, coef = ctx.saved_tensors
        d_out = _req(d_out, torch.float32, "d_out")
        S, N = intensity.shape[0], intensity[0].numel()
        d_int = torch.empty_like(intensity)
        d_coef = torch.zeros_like(coef) if ctx.needs_input_grad[2] else None
        _call("den_lpf_bwd", _ptr(intensity), _ptr(sample_dt), _ptr(coef), S, N, ctx.n_channels,
              _ptr(d_out), _ptr(d_int), _ptr(d_coef), _stream())
        return d_int, None, d_coef, None


def lpf(intensity, sample_dt, coef, n_channels):
    """(S,N) intensities, (S-1,N) sample spacings [ns], 5 fp64 coefficients -> (N, n_channels)."""
    return _LpfFn.apply(intensity, sample_dt, coef, n_channels)


# --------------------------------------------------------------------------- #
# pixel-bandwidth filter fused with the event loss (all requests of a step, one launch)
# --------------------------------------------------------------------------- #
LPF_LOSS_MAX_S = 32
ERROR_KINDS = {"l1": 0, "mse": 1, "huber": 2, "mape": 3}
_LPF_LOSS_WS = {}       # (device, P, N) -> zero-initialised workspace (the kernel leaves it zeroed)


def _lpf_loss_workspace(dev, P, N):
    key = (dev, P, N)
    ws = _LPF_LOSS_WS.get(key)
    if ws is None:
        nbytes = _lib.lib().raw("den_lpf_loss_workspace_bytes")(P, N)
        ws = torch.zeros(nbytes // 8 + 1, dtype=torch.float64, device=dev)
        if len(_LPF_LOSS_WS) > 8:
            _LPF_LOSS_WS.clear()
        _LPF_LOSS_WS[key] = ws
    return ws


class _LpfLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, intensity, sample_dt, coef, reset_dt, target, inv_k, valid, desc):
        intensity = _req(intensity, torch.float32, "intensity")
        sample_dt = _req(sample_dt, torch.float32, "sample_dt")
        coef = _req(coef, torch.float64, "coef")
        inv_k = _req(inv_k, torch.float32, "inv_k")
        valid = _req(valid, torch.uint8, "valid")
        K, S, N = intensity.shape
        P = K // 2
        assert sample_dt.shape == (K, S - 1, N) and valid.shape == (P, N) and inv_k.shape == (P,)
        if reset_dt is not None:
            reset_dt = _req(reset_dt, torch.float64, "reset_dt")
            assert reset_dt.shape == (K, N)
        if target is not None:
            target = _req(target, torch.float32, "target")
            assert target.shape == (P, N)
        dev = intensity.device
        terms = torch.empty(P, dtype=torch.float32, device=dev)
        counts = torch.empty(P, dtype=torch.int32, device=dev)
        log_it = torch.empty((K, N), dtype=torch.float32, device=dev)
        _call("den_lpf_loss_fwd", ctypes.byref(desc), _ptr(intensity), _ptr(sample_dt), _ptr(coef),
              _ptr(reset_dt), _ptr(target), _ptr(inv_k), _ptr(valid), N, _ptr(terms), _ptr(counts),
              _ptr(log_it), _ptr(_lpf_loss_workspace(dev, P, N)), _stream())
        ctx.desc = desc
        ctx.has = (reset_dt is not None, target is not None)
        ctx.save_for_backward(intensity, sample_dt, coef, inv_k, valid, counts,
                              reset_dt if reset_dt is not None else coef,
                              target if target is not None else inv_k)
        ctx.mark_non_differentiable(counts, log_it)
        return terms, log_it, counts

    @staticmethod
    def backward(ctx, d_terms, _d_log_it, _d_counts):
        intensity, sample_dt, coef, inv_k, valid, counts, reset_dt, target = ctx.saved_tensors
        has_reset_dt, has_target = ctx.has
        reset_dt = reset_dt if has_reset_dt else None
        target = target if has_target else None
        N = intensity.shape[2]
        need = ctx.needs_input_grad
        d_terms = _req(d_terms, torch.float32, "d_terms")
        d_int = torch.empty_like(intensity)
        d_coef = torch.zeros_like(coef) if need[2] else None
        d_reset = torch.empty_like(reset_dt) if (has_reset_dt and need[3]) else None
        d_target = torch.zeros_like(target) if (has_target and need[4]) else None
        d_inv_k = torch.zeros(inv_k.shape, dtype=torch.float64, device=inv_k.device) if need[5] else None
        _call("den_lpf_loss_bwd", ctypes.byref(ctx.desc), _ptr(intensity), _ptr(sample_dt), _ptr(coef),
              _ptr(reset_dt), _ptr(target), _ptr(inv_k), _ptr(valid), N, _ptr(counts), _ptr(d_terms),
              _ptr(d_int), _ptr(d_coef), _ptr(d_reset), _ptr(d_target), _ptr(d_inv_k), _stream())
        return (d_int, None, d_coef, d_reset, d_target,
                d_inv_k.to(inv_k.dtype) if d_inv_k is not None else None, None, None)


def lpf_loss(intensity, sample_dt, coef, reset_dt, target, inv_k, valid, kinds, has_target,
             has_reset=True):
    """Filter + loss of the K = 2P render requests of a step.  intensity (K,S,N), sample_dt (K,S-1,N)
    [ns], coef (5) f64, reset_dt (K,N) f64 [ns] or None, target (P,N) or None, inv_k (P), valid (P,N)
    uint8/bool; kinds / has_target: per pair.  Returns (terms (P) — masked means of the per-event
    errors —, log_intensity (K,N) no-grad, counts (P) int32)."""
    K, S, _ = intensity.shape
    if S > LPF_LOSS_MAX_S or K % 2 or not 2 <= K <= 8:
        raise NotImplementedError("lpf_loss: it_sample_size <= 32 and an even number of requests <= 8")
    desc = LpfLossDesc()
    desc.it_sample_size, desc.n_requests, desc.has_reset = S, K, int(bool(has_reset))
    for p in range(K // 2):
        desc.error_kind[p] = ERROR_KINDS[kinds[p]]
        desc.has_target[p] = int(bool(has_target[p]))
    if valid.dtype == torch.bool:
        valid = valid.view(torch.uint8)
    return _LpfLossFn.apply(intensity, sample_dt, coef, reset_dt, target, inv_k, valid, desc)
