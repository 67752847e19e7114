"""B2 host mirror of the reference's ``NeRF`` (models/nerf.py:16-286) fused with its render
loop (external/utils.py:38-140 ``render_image``) and volume renderer
(external/vol_rendering.py:16-128 ``rendering``).

Same constructor arguments, same methods (``update_occ_grid``, ``pixel_params_to_ray``,
``forward``), same state-dict keys; ``forward(o, d)`` returns the reference's 4-tuple
``(radiance, opacity, depth, mean_num_samples_per_ray)``.

One render call is: ray/AABB slab test -> near/far clamp + stratified jitter -> two-pass
occupancy march into an arena -> fused density pre-pass -> sequential-T visibility ->
ballot compaction -> field evaluation -> fused compositor.  Every stage is a den_b200
kernel; torch only allocates, draws the RNG numbers in upstream order, and (for now)
carries the MLP gradient pass through autograd.
"""

import torch

from . import nerfacc, ops
from .field import NGPradianceField
from .nerfacc import ContractionType


class Softplus(torch.nn.Module):
    """utils/modules.py:58-75 — softplus parametrization with its right inverse."""

    def __init__(self, beta=1, threshold=20):
        super().__init__()
        self.beta, self.threshold = beta, threshold

    def forward(self, x):
        return torch.nn.functional.softplus(x, self.beta, self.threshold)

    def right_inverse(self, y):
        inv = torch.log(torch.exp(self.beta * y) - 1) / self.beta
        return torch.where(y * self.beta > self.threshold, y, inv)


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class NeRF(torch.nn.Module):
    def __init__(self, aabb, contraction_type, occ_grid_config, near_plane, far_plane,
                 render_step_size, render_bkgd, cone_angle, early_stop_eps, alpha_thre,
                 test_chunk_size, arch, arch_config, num_dim, radiance_dim, opacity_eps=1e-10):
        super().__init__()
        assert all(r > 0 for r in ([_get(occ_grid_config, "resolution")]
                                   if isinstance(_get(occ_grid_config, "resolution"), int)
                                   else _get(occ_grid_config, "resolution")))
        assert 0 <= _get(occ_grid_config, "occ_thre") <= 1
        assert 0 <= _get(occ_grid_config, "ema_decay") <= 1
        assert _get(occ_grid_config, "warmup_steps") > 0 and _get(occ_grid_config, "n") > 0
        if near_plane is not None and far_plane is not None:
            assert 0 <= near_plane <= far_plane
        assert render_step_size > 0
        assert render_bkgd is None or render_bkgd == "parameter" \
            or isinstance(render_bkgd, torch.Tensor)
        assert cone_angle >= 0 and 0 <= early_stop_eps <= 1 and 0 <= alpha_thre <= 1
        assert test_chunk_size > 0 and num_dim == 3 and radiance_dim > 0 and opacity_eps > 0
        if arch != "ngp":
            raise NotImplementedError("only `arch: ngp` is on the hot path (no shipped config "
                                      "selects `mlp`, configs/train/synthetic.yaml:75)")

        self.register_buffer("aabb", torch.tensor(aabb), persistent=False)
        self._aabb_host = [float(v) for v in aabb]
        self.contraction_type = contraction_type
        self.occ_grid_config = occ_grid_config
        self.near_plane, self.far_plane = near_plane, far_plane
        self.register_buffer("render_step_size", torch.tensor(render_step_size),
                             persistent=False)
        self._step_host = float(render_step_size)
        if render_bkgd is None:
            self.render_bkgd = None
        elif isinstance(render_bkgd, str):
            self.render_bkgd = torch.nn.parameter.Parameter(torch.ones(radiance_dim))
            torch.nn.utils.parametrize.register_parametrization(self, "render_bkgd", Softplus())
        else:
            self.register_buffer("render_bkgd", render_bkgd, persistent=False)
        self.cone_angle = cone_angle
        self.early_stop_eps = early_stop_eps
        self.alpha_thre = alpha_thre
        self.test_chunk_size = test_chunk_size
        self.opacity_eps = opacity_eps

        self.occupancy_grid = nerfacc.OccupancyGrid(
            roi_aabb=aabb, resolution=_get(occ_grid_config, "resolution"),
            contraction_type=contraction_type)
        base = dict(_get(arch_config, "mlp_base"))
        head = dict(_get(arch_config, "mlp_head"))
        head["output_dim"] = radiance_dim
        self.radiance_field = NGPradianceField(
            aabb=aabb, num_dim=num_dim, use_viewdirs=True, contraction_type=contraction_type,
            pos_encoding_config=dict(_get(arch_config, "pos_encoding")),
            dir_encoding_config=dict(_get(arch_config, "dir_encoding")),
            mlp_base_config=base, mlp_head_config=head)
        self.last_num_samples = None        # device int32 scalar of the latest render call
        self.last_marched = None            # marched samples of the latest synchronising render call
        # eval mode: the reference renders `test_chunk_size` rays at a time to bound memory
        # (external/utils.py:99-103); rays are independent and the eval march is deterministic, so any
        # chunking gives the same image — on a 180 GB part the chunks are merged up to this many rays
        # (40 launch sequences per 800x800 view become three).  The merged chunk is sized from the
        # samples per ray seen so far so that one chunk stays below `eval_chunk_samples` samples
        # (~0.4 KB of per-sample buffers each): the first chunk is `test_chunk_size` rays.
        self.eval_chunk_rays = 1 << 20
        self.eval_chunk_samples = 1 << 25
        # Training without host read-backs: the sample buffers of a render
        # call are sized from the samples per ray of the previous calls (x `capacity_margin`, whole
        # allocation quanta), the true counts stay on the device (`n_dev` of every per-sample kernel) and
        # reach the host one call late through an asynchronous copy.  The first call, and any call after an
        # overflow, takes the synchronising path (which also teaches the estimate).
        self.sync_free = True
        self.capacity_margin = 1.1           # + a term for the batch-to-batch spread, see _capacity
        # set by den_clamp_offsets when a sync-free render call ran out of capacity; cleared by the
        # renderer at the start of every optimizer step; FusedAdam skips the update while it is set
        self.register_buffer("overflow_flag", torch.zeros((), dtype=torch.int32), persistent=False)
        self._spr_estimate = None            # marched samples per ray, max of the recent calls
        self._stats = None                   # LaggedReadback of the latest sync-free call
        self.overflow_count = 0

    # ---------------------------------------------------------------- occupancy ------
    def update_occ_grid(self, step, T_wc_position=None):
        """models/nerf.py:170-204: density (fused gather + MLP kernel) times the cone-aware step
        size per drawn cell (`den_occgrid_occ`), handed to the grid's EMA-max / threshold kernels."""
        def occ_eval_fn(x):
            camera_ids = None
            if self.cone_angle > 0.0:
                # the reference's draw (models/nerf.py:177-180), after the grid's own draws
                camera_ids = torch.randint(0, len(T_wc_position), (x.shape[0],),
                                           device=T_wc_position.device)
            sigma = self.radiance_field.density_at(x)
            return ops.occgrid_occ(sigma, x, camera_ids, T_wc_position, self.cone_angle,
                                   self._step_host, self.near_plane, self.far_plane)[:, None]

        self.occupancy_grid.every_n_step(
            step, occ_eval_fn, _get(self.occ_grid_config, "occ_thre"),
            _get(self.occ_grid_config, "ema_decay"), _get(self.occ_grid_config, "warmup_steps"),
            _get(self.occ_grid_config, "n"))

    # --------------------------------------------------------------------- rays ------
    @staticmethod
    def pixel_params_to_ray(intrinsics_inverse, pixel_position, T_wc_position, T_wc_orientation):
        """models/nerf.py:206-228 (small elementwise prologue; the fused trajectory->ray
        kernel `den_rays_from_trajectory` replaces it on the training path)."""
        homog = torch.cat((pixel_position, torch.ones_like(pixel_position[..., :1])), dim=-1)
        d = (T_wc_orientation @ (intrinsics_inverse @ homog.unsqueeze(-1))).squeeze(-1)
        d = d / torch.linalg.vector_norm(d, dim=-1, keepdim=True)
        return T_wc_position, d

    # ------------------------------------------------------------------- render ------
    def _march(self, o, d, jitter, probe=None, capacity=None, overflow=None):
        grid = self.occupancy_grid
        if self.contraction_type == ContractionType.AABB:
            t_min, t_max = ops.ray_aabb_intersect(o, d, self._aabb_host)
        else:
            t_min = torch.zeros_like(o[:, 0])
            t_max = torch.full_like(o[:, 0], 1e10)
        stratified = self.radiance_field.training
        if stratified and jitter is None:
            jitter = torch.rand_like(t_min)
        if not stratified:
            jitter = None
        ops.clamp_jitter_(t_min, t_max, jitter, self.near_plane, self.far_plane, self._step_host)
        params = ops.make_march_params(grid._roi_host, grid._res_host,
                                       self.contraction_type.to_cpp_version(), self._step_host,
                                       self.cone_angle)
        # samples per ray are bounded by the marched span: near..far when both planes are set (the
        # jitter only moves t_min forward), else the AABB diagonal
        if self.near_plane is not None and self.far_plane is not None:
            seg_len = ops.march_segment_length(self.near_plane, self.far_plane, self._step_host)
        elif self.contraction_type == ContractionType.AABB:
            a = self._aabb_host
            diag = sum((a[i + 3] - a[i]) ** 2 for i in range(3)) ** 0.5
            seg_len = ops.march_segment_length(0.0, diag, self._step_host)
        else:
            seg_len = None
        if overflow is not None:
            return ops.march(params, o, d, t_min, t_max, grid.binary, capacity=capacity,
                             seg_len=seg_len, overflow=overflow)
        return ops.march(params, o, d, t_min, t_max, grid.binary, seg_len=seg_len, probe=probe)

    def _segment_bound_known(self):
        return (self.near_plane is not None and self.far_plane is not None) \
            or self.contraction_type == ContractionType.AABB

    # ---------------------------------------------------- sync-free bookkeeping ------
    def _consume_stats(self):
        """Fold the (lagged) counts of the previous sync-free call into the capacity estimate."""
        if self._stats is None:
            return
        vals = self._stats.pop()
        self._stats = None
        if vals is None:
            return
        marched, n_rays, overflow = vals[0], vals[1], vals[2]
        spr = marched / max(n_rays, 1.0)
        if overflow:
            self.overflow_count += 1
            self._spr_estimate = None        # next call: the synchronising path, exact sizes
        elif self._spr_estimate is not None:
            self._spr_estimate = max(0.98 * self._spr_estimate, spr)

    def _capacity(self, n_rays):
        if self._spr_estimate is None:
            return None
        # samples per ray vary from batch to batch like the share of rays that hit the scene: rays of one
        # event (4 render calls x S pixel-bandwidth samples) move together, so a batch holds roughly
        # n_rays / 120 independent draws; five standard deviations of a ~170 % per-event spread on top
        margin = self.capacity_margin + 8.5 / max(n_rays / 120.0, 1.0) ** 0.5
        need = int(self._spr_estimate * margin * n_rays) + 1
        q = ops._ROW_QUANTUM
        return (-(-need // q) + 1) * q

    def render_chunk_sync_free(self, o, d, jitter, groups, capacity):
        """`render_chunk` for training with NO host read: capacity-sized sample buffers, device-side
        counts.  Returns colour, opacity, depth and the per-group sample counts as a DEVICE tensor."""
        field = self.radiance_field
        n_rays = o.shape[0]
        dev = o.device
        overflow = self.overflow_flag
        ray_idx, t0, t1, offsets = self._march(o, d, jitter, capacity=capacity, overflow=overflow)
        n_dev = offsets[-1:]
        marched_total = n_dev
        sig, rgb, enc = field.eval_samples_tc(o, d, ray_idx, t0, t1, full=True, n_dev=n_dev)
        enc_rows = None
        if self.early_stop_eps > 0.0:
            alphas = ops.alpha_from_sigma(sig, t0, t1, n_dev)
            mask, counts = ops.visibility(alphas, offsets, self.early_stop_eps, 0.0)
            offsets_out = ops.exclusive_scan_i32(counts)
            # survivors keep their order; their pre-pass sigma / rgb travel with them (8 B per sample)
            # and the 128-byte encodings are read in place through `enc_rows`
            ray_idx, t0, t1, sig, rgb, enc_rows = ops.compact_ex(
                mask, offsets, offsets_out, ray_idx, t0, t1, capacity, sig, rgb)
            offsets = offsets_out
            n_dev = offsets[-1:]
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in field.parameters())
        if needs_grad:
            enc_node = field.encode_samples(o, d, ray_idx, t0, t1, offsets, enc=enc, n_dev=n_dev)
            sig, rgb = field.mlp_samples(enc_node, o, d, ray_idx, t0, t1, offsets, precomputed=(sig, rgb),
                                         n_dev=n_dev, enc_rows=enc_rows)
        colour, opacity, depth = ops.composite(sig, rgb, t0, t1, offsets, self.render_bkgd)
        per = n_rays // groups
        counts = offsets[::per].diff()                 # (groups,) samples per render call
        # the counts reach the host one call late (capacity estimate, overflow check)
        from .lagged import LaggedReadback
        stats = torch.stack((marched_total[0].double(),
                             torch.full((), float(n_rays), device=dev, dtype=torch.float64),
                             overflow.double()))
        self._stats = LaggedReadback(stats)
        return colour, opacity, depth, counts


    def render_chunk(self, o, d, jitter=None, groups=1):
        """One chunk of rays (R,3),(R,3) -> colour (R,C), opacity (R,), depth (R,), M.
        With ``groups`` > 1 the rays are `groups` equal consecutive blocks (the render calls of one
        training step batched into a single launch sequence) and M is the list of per-block sample
        counts; the counts ride along with the host read of the total (no extra sync)."""
        field = self.radiance_field
        n_rays = o.shape[0]
        probe = None
        if groups > 1:
            assert n_rays % groups == 0
            probe = torch.arange(0, groups + 1, device=o.device) * (n_rays // groups)
        marched = self._march(o, d, jitter, probe)
        ray_idx, t0, t1, offsets = marched[:4]
        self.last_marched = ray_idx.numel()
        bounds = marched[4] if groups > 1 else None

        needs_grad = torch.is_grad_enabled() and any(
            p.requires_grad for p in field.parameters())
        pre = None                      # (sigma, rgb, enc) of the current sample set, if known
        if (self.alpha_thre > 0.0 or self.early_stop_eps > 0.0) and ray_idx.numel() > 0:
            alpha_thre = self.alpha_thre
            if alpha_thre > 0.0:
                alpha_thre = min(alpha_thre, self.occupancy_grid.occs.mean().item())
            pre = field.eval_samples_tc(o, d, ray_idx, t0, t1, full=True)
            alphas = ops.alpha_from_sigma(pre[0], t0, t1)
            mask, counts = ops.visibility(alphas, offsets, self.early_stop_eps, alpha_thre)
            offsets_out = ops.exclusive_scan_i32(counts)
            if groups > 1:
                bounds = [int(v) for v in offsets_out[probe].tolist()]          # host read
                total = bounds[-1]
            else:
                total = int(offsets_out[-1].item())                             # host read
            if total < ray_idx.numel():
                ray_idx, t0, t1 = ops.compact(mask, offsets, offsets_out, ray_idx, t0, t1, total)
                offsets = offsets_out
                # survivors keep their order: the pre-pass outputs stay valid.  The survivor count is
                # already on the host, so the row gather needs no further synchronisation (boolean
                # indexing would read the count back once per tensor)
                keep = torch.nonzero_static(mask.reshape(-1), size=total).reshape(-1)
                pre = tuple(t.index_select(0, keep) for t in pre)
        if pre is None:
            pre = field.eval_samples_tc(o, d, ray_idx, t0, t1, full=True) if ray_idx.numel() \
                else (torch.empty(0, device=o.device),
                      torch.empty(0, field.radiance_dim, device=o.device),
                      torch.empty(0, field.encoding.n_output_dims, device=o.device))
        sigma, rgb, enc = pre
        if needs_grad:
            enc = field.encode_samples(o, d, ray_idx, t0, t1, offsets, enc=enc)
            sigma, rgb = field.mlp_samples(enc, o, d, ray_idx, t0, t1, offsets,
                                           precomputed=(sigma, rgb))
        bkgd = self.render_bkgd
        colour, opacity, depth = ops.composite(sigma, rgb, t0, t1, offsets, bkgd)
        if groups > 1:
            return colour, opacity, depth, [bounds[g + 1] - bounds[g] for g in range(groups)]
        return colour, opacity, depth, ray_idx.numel()

    def forward(self, ray_origin, ray_direction, jitter=None, groups=1):
        """``groups`` > 1 (training only): the leading dimension of the rays indexes `groups`
        independent render calls evaluated as one; the fourth return value is then the list of
        their mean samples per ray."""
        shape = ray_origin.shape
        o = ray_origin.reshape(-1, 3).float().contiguous()
        d = ray_direction.reshape(-1, 3).float().contiguous()
        n_rays = o.shape[0]
        if groups > 1:
            assert self.radiance_field.training and shape[0] == groups
            per = n_rays // groups
            if jitter is None:
                # one draw per render call, in call order: the RNG stream of `groups` separate calls
                jitter = torch.cat([torch.rand(per, device=o.device) for _ in range(groups)])
            jitter = jitter.reshape(-1).contiguous()
            self._consume_stats()
            capacity = self._capacity(n_rays) if (
                self.sync_free and o.is_cuda and self.alpha_thre == 0.0
                and self._segment_bound_known()) else None
            if capacity is not None:
                col, opa, dep, counts = self.render_chunk_sync_free(o, d, jitter, groups, capacity)
                means = list((counts.to(torch.float64) / max(per, 1)).unbind(0))      # device scalars
            else:
                col, opa, dep, counts = self.render_chunk(o, d, jitter, groups)
                means = [c / max(per, 1) for c in counts]
                if self.last_marched is not None:
                    self._spr_estimate = self.last_marched / max(n_rays, 1)
            radiance = col.view(*shape[:-1], -1).squeeze(dim=-1)
            opacity = opa.view(*shape[:-1])
            depth = dep.view(*shape[:-1]) / (opacity + self.opacity_eps)
            return radiance, opacity, depth, means
        training = self.radiance_field.training
        chunk = n_rays if training else self.test_chunk_size
        cols, opas, deps, total = [], [], [], 0
        i = 0
        while i < max(n_rays, 1):
            jit = None if jitter is None else jitter.reshape(-1)[i:i + chunk].contiguous()
            col, opa, dep, m = self.render_chunk(o[i:i + chunk], d[i:i + chunk], jit)
            cols.append(col)
            opas.append(opa)
            deps.append(dep)
            total += m
            done = min(i + chunk, n_rays) - i
            i += max(chunk, 1)
            if not training:
                per_ray = max(m / max(done, 1), 1.0)
                chunk = int(min(max(self.eval_chunk_samples / per_ray, self.test_chunk_size),
                                max(self.eval_chunk_rays, self.test_chunk_size)))
        colour = torch.cat(cols).view(*shape[:-1], -1)
        opacity = torch.cat(opas).view(*shape[:-1])
        depth = torch.cat(deps).view(*shape[:-1])
        radiance = colour.squeeze(dim=-1)
        depth = depth / (opacity + self.opacity_eps)
        return radiance, opacity, depth, total / max(n_rays, 1)
