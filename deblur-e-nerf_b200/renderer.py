"""B2 host of the hot path: the event-supervised training step of the reference's
LightningModule (models/deblur_e_nerf.py:396-586) and its render helpers
(``render_log_intensity`` :1129-1160, ``render_train_pixels`` :1162-1183,
``render_pixels`` :1185-1221, ``update_train_batch_size`` :1252-1308), without Lightning.

``EventRenderer.training_step(batch)`` takes the reference's batch dict
(``{"event": {position, start_ts, end_ts, num_pos, num_neg}, "normalized": {ts_diff,
diff_start_ts, ts_subdiff, subdiff_start_ts[, interval_gen]}}``, with or without the
DataLoader's leading dim of 1) and returns the scalar loss; ``self.logged`` holds what the
reference passes to ``self.log``.  Mono sensors (``channel_idx is None``: every shipped
config, SURVEY.md conventions table C = 1).
"""

import functools

import torch


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class EventRenderer(torch.nn.Module):
    def __init__(self, nerf, trajectory, contrast_threshold, refractory_period, pixel_bandwidth,
                 loss, train_intrinsics_inv, min_modeled_intensity=0.001,
                 train_ray_sample_batch_size=131072, accumulate_grad_batches=1,
                 world_size=1):
        super().__init__()
        self.nerf = nerf
        self.trajectory = trajectory
        self.contrast_threshold = contrast_threshold
        self.refractory_period = refractory_period
        self.pixel_bandwidth = pixel_bandwidth          # None <=> pixel_bandwidth.enable: false
        self.loss = loss
        self.register_buffer("train_intrinsics_inv", train_intrinsics_inv, persistent=False)
        self.min_modeled_intensity = min_modeled_intensity
        # models/deblur_e_nerf.py:72-75 — the sample budget is split evenly over the GPUs
        self.train_ray_sample_batch_size = train_ray_sample_batch_size // world_size
        self.accumulate_grad_batches = accumulate_grad_batches
        self.render_bkgd = "parameter" if nerf.render_bkgd is not None else None
        self.logged = {}
        self.next_train_batch_size = None
        self._jitters = None
        self.mean_samples_reduce_fn = None              # set by ddp.attach(): all-reduce mean
        # evaluate the (up to) four render calls of a training step as ONE march / field /
        # compositing launch sequence over 4x the rays (same arithmetic per ray, same RNG draws in
        # the same order): a quarter of the launches and host synchronisations per step
        self.batch_render_calls = True
        # with the pixel-bandwidth model on, filter + reset + loss of all requests run as ONE kernel per
        # direction (den_lpf_loss_fwd / _bwd) instead of a filter launch per request + eager loss ops
        self.fuse_lpf_loss = True
        # reverse mode of the ray generation w.r.t. the timestamps (the tau path): "kernel" = closed form
        # per pose interval (den_rays_from_trajectory_bwd), "autograd" = the reference's torch form
        self.rays_reverse_mode = "kernel"
        self.batch_bytes_per_sample = 448        # encodings + their gradient + the per-sample scalars
        self._last_mean_samples = None           # samples per ray of the previous step (memory guard)

    # ------------------------------------------------------------- render helpers ----
    def _kinv_host(self):
        """K^-1 as 9 host floats (read back once; the intrinsics are a constant buffer)."""
        if getattr(self, "_kinv_cache", None) is None:
            self._kinv_cache = [float(v) for v in self.train_intrinsics_inv.detach().cpu().reshape(-1)]
        return self._kinv_cache

    def rays(self, timestamp, pixel_position):
        """Camera rays of `pixel_position` (N,2) at `timestamp` (..., N): one fused kernel
        (trajectory interpolation + pinhole model); when the timestamps carry a gradient (the
        refractory-period path) its reverse mode is one kernel too (dL/dt in closed form).
        `rays_reverse_mode = "autograd"` keeps the term-by-term torch form of the reference for that
        path (its fp32 rounding differs from the closed form's: tests/test_gpu_b2_renderer.py)."""
        if pixel_position.requires_grad or not timestamp.is_cuda or (
                timestamp.requires_grad and self.rays_reverse_mode == "autograd"):
            pos, rot = self.trajectory(timestamp)
            return self.nerf.pixel_params_to_ray(self.train_intrinsics_inv, pixel_position, pos, rot)
        from . import ops
        tr = self.trajectory
        fn = ops.rays_from_trajectory_grad if timestamp.requires_grad else ops.rays_from_trajectory
        return fn(timestamp, pixel_position, tr.T_wc_timestamp, tr.T_wc_position,
                  tr.T_wc_orientation_quat, self._kinv_host())

    def render_pixels(self, intrinsics_inverse, pixel_position, T_wc_position, T_wc_orientation):
        o, d = self.nerf.pixel_params_to_ray(intrinsics_inverse, pixel_position, T_wc_position,
                                             T_wc_orientation)
        jitter = self._jitters.pop(0) if self._jitters else None
        intensity, opacity, depth, mean_samples = self.nerf(o, d, jitter=jitter)
        if intensity.dim() > opacity.dim():                       # colour: channels first (:1200-1201)
            intensity = intensity.permute(-1, *range(opacity.dim()))
        intensity = intensity + self.min_modeled_intensity
        if self.render_bkgd is None:
            is_valid = opacity > 0
        else:
            is_valid = torch.ones_like(opacity, dtype=torch.bool)
        depth = depth * torch.sum(d * T_wc_orientation[..., 2], dim=-1)
        return intensity, opacity, depth, mean_samples, is_valid

    # ----------------------------------------------------------------- evaluation ----
    @staticmethod
    def image_pixel_positions(height, width, device=None):
        """(H, W, 2) pixel positions (x = column, y = row) of an image: the reference's
        `val_img_pixel_pos` / `test_img_pixel_pos` buffers (models/deblur_e_nerf.py:120-127,153-160)."""
        xs, ys = torch.meshgrid(torch.arange(width, device=device), torch.arange(height, device=device),
                                indexing="xy")
        return torch.stack((xs, ys), dim=2).to(torch.get_default_dtype())

    @torch.no_grad()
    def evaluation_step(self, batch, intrinsics_inv, img_pixel_pos):
        """models/deblur_e_nerf.py:604-652 (validation_step / test_step): render the posed view
        `batch` = {img ([3,] H, W), T_wc_position (3), T_wc_orientation (3, 3)[, sample_id, exposure_time,
        gain]} at every pixel of `img_pixel_pos` (H, W, 2) and return the reference's output dict.  A
        DataLoader's leading dim of 1 on every entry is removed like upstream (:609-613).  With the field
        in eval mode the march is deterministic and runs `test_chunk_size` rays per chunk."""
        batch = dict(batch)
        if batch["T_wc_position"].dim() == 2:
            assert len(batch["img"]) == 1
            batch = {k: v.squeeze(0) if torch.is_tensor(v) else v for k, v in batch.items()}
        target = batch["img"]
        H, W = img_pixel_pos.shape[:2]
        assert tuple(target.shape[-2:]) == (H, W), "the target image and the pixel grid differ in size"
        device = img_pixel_pos.device
        pos = batch["T_wc_position"].to(device).view(1, 1, 3).expand(H, W, -1)
        rot = batch["T_wc_orientation"].to(device).view(1, 1, 3, 3).expand(H, W, -1, -1)
        pred, _, _, _, _ = self.render_pixels(intrinsics_inv, img_pixel_pos, pos, rot)
        one = torch.ones((), device=device)
        return {
            "sample_id": batch.get("sample_id"),
            "pred_intensity_img": pred,
            "target_intensity_img": target.to(device),
            "exposure_time": batch.get("exposure_time", one.to(torch.int64)),      # :636-643: unity defaults
            "gain": batch.get("gain", one.to(torch.get_default_dtype())),
        }

    @torch.no_grad()
    def evaluation_epoch_end(self, outputs, min_normalized_pixel_value, max_normalized_pixel_value,
                             black_level_offset=False, per_channel_log_it_scale=True, init_correction=None,
                             max_steps=10, radius=1e6, stage="test"):
        """models/deblur_e_nerf.py:674-969 on the device (`eval_post.evaluate`): stack the views, remove
        the affine ambiguity of the predicted log intensity (with the optional black-level-offset
        refinement) and average L1 / PSNR / SSIM over the views.  The images stay in HBM (the reference
        moves them to the host, :714-718); LPIPS is not computed (no lpips network in this image).
        Returns ({"<stage>/l1", "<stage>/psnr", "<stage>/ssim"}, corrected predictions (B, C, H, W))."""
        from . import eval_post
        pred = torch.stack([o["pred_intensity_img"] for o in outputs])
        target = torch.stack([o["target_intensity_img"] for o in outputs]).to(pred.device, torch.float32)
        exposure = torch.stack([torch.as_tensor(o["exposure_time"]).reshape(()) for o in outputs]).to(pred.device)
        gain = torch.stack([torch.as_tensor(o["gain"]).reshape(()) for o in outputs]).to(pred.device)
        res = eval_post.evaluate(pred.float(), target, exposure, gain, float(min_normalized_pixel_value),
                                 float(max_normalized_pixel_value), black_level_offset=black_level_offset,
                                 init=init_correction, max_steps=max_steps, radius=radius,
                                 per_channel_scale=per_channel_log_it_scale)
        metrics = {f"{stage}/{k}": res[k] for k in ("l1", "psnr", "ssim") if res[k] is not None}
        self.logged.update(metrics)
        return metrics, res["pred"]

    @staticmethod
    def bayering(intensity, channel_idx):
        """models/deblur_e_nerf.py:1223-1234: a colour sensor behind a Bayer filter sees, at pixel n, only
        channel `channel_idx[n]` of the rendered radiance.  `intensity` (..., N, 3) -> (..., N)."""
        idx = channel_idx.reshape((1,) * (intensity.dim() - 2) + (-1, 1)).expand(*intensity.shape[:-1], 1)
        return intensity.gather(-1, idx).squeeze(-1)

    def render_train_pixels(self, timestamp, pixel_position, pixel_channel_idx=None):
        """models/deblur_e_nerf.py:1162-1183 (training never uses the depth, so the camera-z
        correction of render_pixels is skipped and the rays come from the fused kernel)."""
        o, d = self.rays(timestamp, pixel_position)
        jitter = self._jitters.pop(0) if self._jitters else None
        intensity, opacity, _, mean_samples = self.nerf(o, d, jitter=jitter)
        intensity = intensity + self.min_modeled_intensity
        if pixel_channel_idx is not None:
            intensity = self.bayering(intensity, pixel_channel_idx)
        hit = opacity > 0
        is_valid = hit if self.render_bkgd is None else torch.ones_like(hit)
        occ_rate = torch.mean(hit, dtype=torch.get_default_dtype())
        return intensity, occ_rate, mean_samples, is_valid

    def render_log_intensity(self, timestamp, pixel_position, pixel_channel_idx=None,
                             normalized_interval_gen=None, reset_diff=False):
        if self.pixel_bandwidth is not None:
            fn = functools.partial(self.render_train_pixels, pixel_position=pixel_position,
                                   pixel_channel_idx=pixel_channel_idx)
            log_it, aux = self.pixel_bandwidth(normalized_interval_gen, timestamp, fn, reset_diff)
            occ_rate, mean_samples, is_valid = aux
            return log_it, occ_rate, mean_samples, is_valid.any(dim=0)
        intensity, occ_rate, mean_samples, is_valid = self.render_train_pixels(
            timestamp, pixel_position, pixel_channel_idx)
        return intensity.log(), occ_rate, mean_samples, is_valid

    def render_log_intensity_batched(self, requests, pixel_position, normalized_interval_gen,
                                     pixel_channel_idx=None):
        """`render_log_intensity` for several (timestamp, reset_diff) requests of the same pixels
        at once.  Returns one (log_intensity, occ_rate, mean_samples, is_valid) tuple per request,
        in order (the pixel-bandwidth reset state is carried from request to request exactly as
        the sequential calls do)."""
        pb = self.pixel_bandwidth
        K = len(requests)
        life = coeff = None
        if pb is not None:
            # the same for every request of the step: evaluated once (host time, not device time:
            # at the reference's batch sizes the step is launch-bound after its last host read)
            life = pb.sample_lifetimes(normalized_interval_gen)
            coeff = pb.coefficients()
            ts_all = torch.stack([(ts - life).clamp(min=pb.min_ts) for ts, _ in requests])   # (K, S, N)
        else:
            ts_all = torch.stack([ts for ts, _ in requests])             # (K, N)
        o, d = self.rays(ts_all, pixel_position)
        jitter = None
        if self._jitters:
            jitter = torch.cat([self._jitters.pop(0).reshape(-1) for _ in range(K)])
        intensity, opacity, _, means = self.nerf(o, d, jitter=jitter, groups=K)
        intensity = intensity + self.min_modeled_intensity
        if pixel_channel_idx is not None:
            intensity = self.bayering(intensity, pixel_channel_idx)
        hit = opacity > 0
        is_valid = hit if self.render_bkgd is None else torch.ones_like(hit)
        occ = hit.reshape(K, -1).to(torch.get_default_dtype()).mean(dim=1)
        out = []
        for k, (ts, reset_diff) in enumerate(requests):
            if pb is not None:
                cached = (intensity[k], occ[k], means[k], is_valid[k])
                log_it, aux = pb(normalized_interval_gen, ts, lambda _ts, c=cached: c, reset_diff,
                                 lifetimes=life, coefficients=coeff)
                out.append((log_it, aux[0], aux[1], aux[2].any(dim=0)))
            else:
                out.append((intensity[k].log(), occ[k], means[k], is_valid[k]))
        return out

    def render_and_loss_fused(self, event, segs, pixel_position, normalized_interval_gen, mean_ct,
                              pixel_channel_idx=None):
        """All render requests of the step as one launch sequence, then the pixel-bandwidth filter of
        every request, the reset carried from the first one, the pair differences and the masked loss
        means in ONE kernel (`den_lpf_loss_fwd`; reverse mode `den_lpf_loss_bwd`).  `segs`: the
        supervision intervals [(seg dict, is_diff)], at most two.  Returns (terms {name: mean},
        mean samples per ray per request, occupancy rate per request)."""
        from . import ops
        pb, loss = self.pixel_bandwidth, self.loss
        life = pb.sample_lifetimes(normalized_interval_gen)                    # (S, N) f64, no grad
        stamps = []
        for seg, _ in segs:
            stamps += [seg["start_ts"], seg["end_ts"]]
        out_ts = torch.stack(stamps)                                           # (K, N) f64
        K = out_ts.shape[0]
        sample_ts = out_ts[:, None, :] - life[None]                            # (K, S, N)
        o, d = self.rays(sample_ts.clamp(min=pb.min_ts), pixel_position)
        jitter = None
        if self._jitters:
            jitter = torch.cat([self._jitters.pop(0).reshape(-1) for _ in range(K)])
        intensity, opacity, _, means = self.nerf(o, d, jitter=jitter, groups=K)
        intensity = intensity + self.min_modeled_intensity                    # (K, S, N [, 3])
        if pixel_channel_idx is not None:
            intensity = self.bayering(intensity, pixel_channel_idx).contiguous()
        hit = opacity > 0
        occ = hit.reshape(K, -1).to(torch.get_default_dtype()).mean(dim=1)
        P = K // 2
        if self.render_bkgd is None:
            per_request = hit.any(dim=1)                                       # (K, N)
            valid = per_request[0::2] | per_request[1::2]                      # (P, N)
        else:
            valid = torch.ones((P, out_ts.shape[1]), dtype=torch.bool, device=out_ts.device)
        sample_dt = sample_ts.detach().diff(dim=1).to(intensity.dtype)         # (K, S-1, N) ns
        has_reset = segs[0][1]
        reset_dt = out_ts - out_ts[0] if has_reset else None
        names, kinds, has_target, inv_k, rows = [], [], [], [], []
        for seg, is_diff in segs:
            name = "log_intensity_diff" if is_diff else "log_intensity_tv"
            names.append(name)
            kinds.append(_get(loss.error_fn, name))
            k = mean_ct if _get(loss.normalize, name) else torch.ones_like(mean_ct)
            inv_k.append(1.0 / k)
            has_target.append(is_diff)
            if is_diff:         # loss_metric/loss.py:44-60: target = ts_diff * dL/dt of the event / k
                grad = event["log_intensity_diff"] / (event["end_ts"] - event["start_ts"])
                event["log_intensity_grad"] = grad
                rows.append((seg["ts_diff"] * grad / k).to(intensity.dtype))
            else:               # :82-94: total variation, target zero
                rows.append(None)
        target = None
        if any(r is not None for r in rows):
            zero = torch.zeros_like(next(r for r in rows if r is not None))
            target = torch.stack([zero if r is None else r for r in rows])
        terms, log_it, _ = ops.lpf_loss(intensity, sample_dt, pb.coefficients(), reset_dt, target,
                                        torch.stack(inv_k).to(intensity.dtype), valid, kinds,
                                        has_target, has_reset)
        for p, (seg, _) in enumerate(segs):
            seg["log_intensity_diff"] = log_it[2 * p + 1] - log_it[2 * p]
            seg["is_valid"] = valid[p]
        return {name: terms[p] for p, name in enumerate(names)}, list(means), list(occ.unbind(0))

    # ------------------------------------------------------------------ the step -----
    @staticmethod
    def supervision_timestamps(event, normalized, use_diff, use_tv):
        """models/deblur_e_nerf.py:419-455 (float64 ns)."""
        diff = subdiff = None
        tv_s, tv_e = event["start_ts"], event["end_ts"]
        if use_diff:
            ts_diff = (event["end_ts"] - event["start_ts"]) * normalized["ts_diff"]
            start = torch.lerp(event["start_ts"],
                               torch.max(event["end_ts"] - ts_diff, event["start_ts"]),
                               normalized["diff_start_ts"])
            end = torch.min(start + ts_diff, event["end_ts"])
            diff = {"ts_diff": ts_diff, "start_ts": start, "end_ts": end}
            tv_s, tv_e = start, end
        if use_tv:
            ts_sub = (tv_e - tv_s) * normalized["ts_subdiff"]
            start = torch.lerp(tv_s, torch.max(tv_e - ts_sub, tv_s), normalized["subdiff_start_ts"])
            end = torch.min(start + ts_sub, tv_e)
            subdiff = {"ts_diff": ts_sub, "start_ts": start, "end_ts": end}
        return diff, subdiff

    def training_step(self, batch, batch_index=0, global_step=0, jitters=None):
        # every parametrised tensor (softplus of C_p, tau, Omega, the background) is evaluated once
        # per step instead of once per access: ~30 fewer tiny launches on the host's critical path
        with torch.nn.utils.parametrize.cached():
            return self._training_step(batch, batch_index, global_step, jitters)

    def _training_step(self, batch, batch_index, global_step, jitters):
        self._jitters = list(jitters) if jitters is not None else None
        event = {k: (v.squeeze(0) if v.dim() > 1 and v.shape[0] == 1 and k != "position"
                     else v) for k, v in batch["event"].items()}
        if event["position"].dim() == 3:
            event["position"] = event["position"].squeeze(0)
        normalized = {k: (v.squeeze(0) if v.shape[0] == 1 and v.dim() > 1 else v)
                      for k, v in batch["normalized"].items()}
        size = event["start_ts"].numel()
        for value in normalized.values():
            assert value.shape[-1] == size

        event = self.contrast_threshold(event)
        event = self.refractory_period(event)
        weight = self.loss.loss_weight
        use_diff = _get(weight, "log_intensity_diff") > 0
        use_tv = _get(weight, "log_intensity_tv") > 0
        diff, subdiff = self.supervision_timestamps(event, normalized, use_diff, use_tv)
        gen = normalized.get("interval_gen")

        if batch_index % self.accumulate_grad_batches == 0:
            if self.nerf.overflow_flag.is_cuda:
                self.nerf.overflow_flag.zero_()         # a new optimizer step: nothing has overflowed yet
            self.nerf.update_occ_grid(step=global_step,
                                      T_wc_position=self.trajectory.T_wc_position)

        mean_samples, occ_rates, valid_rates = [], [], []
        segs = [(seg, is_diff) for seg, is_diff in ((diff, True), (subdiff, False))
                if seg is not None]
        batched = None
        fits = self.batch_render_calls and self.nerf.radiance_field.training and segs \
            and self._batch_fits(2 * len(segs) * size, gen)
        # models/deblur_e_nerf.py:409-412: a Bayer sensor's events carry the colour channel of their pixel
        channel_idx = event.get("channel_idx")
        if channel_idx is not None:
            channel_idx = channel_idx.reshape(-1).to(torch.int64)
        if fits and self.fuse_lpf_loss and self.pixel_bandwidth is not None \
                and event["position"].is_cuda and gen.shape[0] + 1 <= 32 and segs[0][1]:
            terms, mean_samples, occ_rates = self.render_and_loss_fused(
                event, segs, event["position"], gen, self.contrast_threshold.mean_contrast_threshold,
                channel_idx)
            return self._finish_step(terms, mean_samples, occ_rates, weight, size, batch_index)
        if fits:
            requests = []
            for seg, is_diff in segs:
                requests += [(seg["start_ts"], is_diff), (seg["end_ts"], False)]
            batched = self.render_log_intensity_batched(requests, event["position"], gen, channel_idx)
        for seg, is_diff in segs:
            if batched is not None:
                (a, occ_a, ms_a, va), (b, occ_b, ms_b, vb) = batched[0], batched[1]
                batched = batched[2:]
            else:
                a, occ_a, ms_a, va = self.render_log_intensity(
                    seg["start_ts"], event["position"], channel_idx, gen, reset_diff=is_diff)
                b, occ_b, ms_b, vb = self.render_log_intensity(
                    seg["end_ts"], event["position"], channel_idx, gen)
            seg["log_intensity_diff"] = b - a
            seg["is_valid"] = va | vb
            mean_samples += [ms_a, ms_b]
            occ_rates += [occ_a, occ_b]
            valid_rates += [va, vb]

        terms = self.loss.compute(event, diff, subdiff,
                                  self.contrast_threshold.mean_contrast_threshold)
        return self._finish_step(terms, mean_samples, occ_rates, weight, size, batch_index)

    def _finish_step(self, terms, mean_samples, occ_rates, weight, size, batch_index):
        if mean_samples and torch.is_tensor(mean_samples[0]):
            # sync-free render calls: the sample counts are device scalars.  The batch controller and the
            # memory guard consume them one step late (an asynchronous copy started now, read when the
            # next step reaches this point), so the step never waits for the device.
            from .lagged import LaggedReadback
            self._apply_lagged_controller()
            stacked = torch.stack(mean_samples).double()
            mean_dev = stacked.mean()
            if self.mean_samples_reduce_fn is not None:
                mean_dev = self.mean_samples_reduce_fn(mean_dev)
            self._pending_controller = (LaggedReadback(torch.stack((mean_dev, stacked.max()))), batch_index)
            mean_samples = mean_dev
        else:
            self._last_mean_samples = max(mean_samples) if mean_samples else None
            mean_samples = self.update_train_batch_size(mean_samples, batch_index)
        loss = sum(v * _get(weight, k) for k, v in terms.items())

        self.logged = {"train/loss": loss.detach(), "train/batch_size": size,
                       "train/mean_num_samples_per_ray": mean_samples}
        for key, value in terms.items():
            self.logged[f"train/{key}"] = value.detach()
        self.logged["train/mean_ray_occ_rate"] = sum(occ_rates) / len(occ_rates)
        return loss

    def _apply_lagged_controller(self):
        pending = getattr(self, "_pending_controller", None)
        if pending is None:
            return
        readback, batch_index = pending
        self._pending_controller = None
        mean, largest = readback.pop()
        self._last_mean_samples = largest
        acc = self.accumulate_grad_batches
        if acc > 1 and (batch_index % acc) != (acc - 2):
            return
        self.next_train_batch_size = int(self.train_ray_sample_batch_size / max(mean, 1e-9))

    def _batch_fits(self, n_render_rays, gen):
        """Memory guard of the batched render calls: estimated per-sample buffers (from the samples
        per ray of the previous step; 128 before the first) against 70 % of the free device memory."""
        if not torch.cuda.is_available():
            return True
        rays = n_render_rays * (gen.shape[0] + 1 if (gen is not None and self.pixel_bandwidth is not None)
                                else 1)
        per_ray = self._last_mean_samples if self._last_mean_samples else 128.0
        # device capacity minus what this process holds (allocator statistics: no driver call — the
        # cudaMemGetInfo behind torch.cuda.mem_get_info costs milliseconds per step)
        if getattr(self, "_device_bytes", None) is None:
            self._device_bytes = torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory
        free = self._device_bytes - torch.cuda.memory_allocated()
        return rays * per_ray * self.batch_bytes_per_sample < 0.7 * free

    def step_fits_batched(self, batch):
        """Would `training_step` evaluate this batch's render calls as ONE launch sequence (the only form
        without host read-backs, i.e. the only one a CUDA graph can record)?  The memory guard of
        `_training_step`, from the batch alone: 2 render calls per supervised interval."""
        weight = self.loss.loss_weight
        n_segs = int(_get(weight, "log_intensity_diff") > 0) + int(_get(weight, "log_intensity_tv") > 0)
        size = batch["event"]["start_ts"].numel()
        gen = batch["normalized"].get("interval_gen")
        if gen is not None and gen.dim() == 3 and gen.shape[0] == 1:
            gen = gen.squeeze(0)
        return self.batch_render_calls and n_segs > 0 and self._batch_fits(2 * n_segs * size, gen)

    def update_train_batch_size(self, mean_samples_per_call, batch_index):
        """models/deblur_e_nerf.py:1252-1308: N_next = int(budget / mean samples per ray)."""
        mean = sum(mean_samples_per_call) / len(mean_samples_per_call)
        if self.mean_samples_reduce_fn is not None:
            mean = self.mean_samples_reduce_fn(mean)
        acc = self.accumulate_grad_batches
        if acc > 1 and (batch_index % acc) != (acc - 2):
            return mean
        self.next_train_batch_size = int(self.train_ray_sample_batch_size / max(mean, 1e-9))
        return mean
