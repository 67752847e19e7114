// Timestamps + pixels -> rays in one kernel (sm_100a).
//
// Replaces, on the path where the timestamps carry no gradient, the ~40 elementwise / gather / bmm
// launches of LinearTrajectory.forward (models/trajectories.py:30-90: searchsorted over the pose
// stamps, float64 interpolation weight, LERP of the position, shortest-path SLERP of the orientation
// through the rotation vector — utils/tensor_ops.py:118-184 with RoMa 1.2.7's rotvec <-> quaternion
// maps, scripts/preprocess_esim.py:390-393 XYZW convention — and quaternion -> rotation matrix) and of
// NeRF.pixel_params_to_ray (models/nerf.py:206-228: d = R K^-1 [u, v, 1], normalised; o = position).
// One thread per ray; the pose table (C ~ 1 000 poses, 36 KB) stays L1/L2 resident.  The formulas are
// the torch host path's, term by term (deblur_e_nerf_b200/trajectories.py), so both paths agree to
// fp32 rounding.
//
// Reverse mode (the refractory-period path: timestamps = event time - tau carry a gradient, config 4,
// models/deblur_e_nerf.py:465-469): poses are buffers, so the only gradient is dL/dt per ray, and inside a
// pose interval it has a closed form.  With w = (t - t_left) / width: position = lerp(p0, p1, w) gives
// d o / d w = p1 - p0; orientation R(w) = R0 Exp(w r), r = rotation vector of q0^-1 q1 (body frame),
// gives d R / d w = R(w) [r]x, and since the direction is d = R(w) c / |c| with a fixed camera-frame c,
// d d / d w = (R(w) r) x d = (R0 r) x d  (Exp(w r) leaves r in place).  Hence
//     dL/dt = (dL/do . (p1 - p0) + dL/dd . ((R0 r) x d)) / width
// — one thread per ray, no saved intermediates but the forward direction.  The torch autograd form
// (trajectories.py) differentiates the same functions term by term; tests compare the two.
#include "den_common.cuh"

namespace den {

__device__ __forceinline__ void quat_mul(const float p[4], const float q[4], float out[4]) {   // xyzw
    const float pw = p[3], qw = q[3];
    out[0] = pw * q[0] + qw * p[0] + (p[1] * q[2] - p[2] * q[1]);
    out[1] = pw * q[1] + qw * p[1] + (p[2] * q[0] - p[0] * q[2]);
    out[2] = pw * q[2] + qw * p[2] + (p[0] * q[1] - p[1] * q[0]);
    out[3] = pw * qw - (p[0] * q[0] + p[1] * q[1] + p[2] * q[2]);
}

__global__ void __launch_bounds__(256)
rays_from_trajectory_kernel(const double* __restrict__ ts, const float* __restrict__ pix, int64_t n_pix,
                            const int64_t* __restrict__ pose_ts, const float* __restrict__ pose_pos,
                            const float* __restrict__ pose_quat, int32_t n_poses, float k00, float k01,
                            float k02, float k10, float k11, float k12, float k20, float k21, float k22,
                            float* __restrict__ rays_o, float* __restrict__ rays_d, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double t = ts[i];
        // torch.searchsorted(T, t) (right = False): first index with T[idx] >= t
        int lo = 0, hi = n_poses;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((double)pose_ts[mid] < t) lo = mid + 1; else hi = mid;
        }
        int right = lo;
        int left = (t == (double)pose_ts[0]) ? right : right - 1;
        left = min(max(left, 0), n_poses - 2);
        right = min(max(right, 1), n_poses - 1);
        const double t_l = (double)pose_ts[left];
        const double width = (double)(pose_ts[left + 1] - pose_ts[left]);
        const float w = (float)((t - t_l) / width);

        // position: torch.lerp(a, b, w) = |w| < 0.5 ? a + w (b - a) : b - (b - a) (1 - w)
        float pos[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float a = pose_pos[3 * left + d], b = pose_pos[3 * right + d];
            const float diff = b - a;
            pos[d] = fabsf(w) < 0.5f ? a + w * diff : b - diff * (1.f - w);
        }
        // orientation: q = q0 * exp(w * log(q0^-1 q1)), shortest path
        float q0[4], q1[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) { q0[d] = pose_quat[4 * left + d]; q1[d] = pose_quat[4 * right + d]; }
        const float dot = q0[0] * q1[0] + q0[1] * q1[1] + q0[2] * q1[2] + q0[3] * q1[3];
        if (dot < 0.f) {
#pragma unroll
            for (int d = 0; d < 4; ++d) q1[d] = -q1[d];
        }
        const float q0c[4] = {-q0[0], -q0[1], -q0[2], q0[3]};
        float rel[4];
        quat_mul(q0c, q1, rel);
        // full rotation vector of `rel` (RoMa unitquat_to_rotvec, angle in [0, 2 pi))
        const float vnorm = sqrtf(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]);
        const float angle = 2.f * atan2f(vnorm, rel[3]);
        const bool small_a = fabsf(angle) <= 1e-3f;
        const float a2 = angle * angle;
        const float scale_v = small_a ? 2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f : angle / sinf(angle * 0.5f);
        float rv[3] = {w * scale_v * rel[0], w * scale_v * rel[1], w * scale_v * rel[2]};
        // rotation vector -> quaternion (RoMa rotvec_to_unitquat)
        const float theta = sqrtf(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
        const bool small_t = theta <= 1e-3f;
        const float th2 = theta * theta;
        const float scale_q = small_t ? 0.5f - th2 / 48.f + th2 * th2 / 3840.f : sinf(theta * 0.5f) / theta;
        const float dq[4] = {scale_q * rv[0], scale_q * rv[1], scale_q * rv[2], cosf(theta * 0.5f)};
        float q[4];
        quat_mul(q0, dq, q);
        // quaternion -> rotation matrix (no normalisation, as the reference)
        const float x = q[0], y = q[1], z = q[2], qw = q[3];
        const float x2 = x * x, y2 = y * y, z2 = z * z, w2 = qw * qw;
        const float xy = x * y, zw = z * qw, xz = x * z, yw = y * qw, yz = y * z, xw = x * qw;
        const float R[3][3] = {{x2 - y2 - z2 + w2, 2.f * (xy - zw), 2.f * (xz + yw)},
                               {2.f * (xy + zw), -x2 + y2 - z2 + w2, 2.f * (yz - xw)},
                               {2.f * (xz - yw), 2.f * (yz + xw), -x2 - y2 + z2 + w2}};
        // direction: R (K^-1 [u, v, 1]), normalised
        const int64_t pi = i % n_pix;
        const float u = pix[2 * pi], v = pix[2 * pi + 1];
        const float c0 = k00 * u + k01 * v + k02, c1 = k10 * u + k11 * v + k12, c2 = k20 * u + k21 * v + k22;
        float dir[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) dir[d] = R[d][0] * c0 + R[d][1] * c1 + R[d][2] * c2;
        const float inv = 1.f / sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            rays_o[3 * i + d] = pos[d];
            rays_d[3 * i + d] = dir[d] * inv;
        }
    }
}

__global__ void __launch_bounds__(256)
rays_from_trajectory_bwd_kernel(const double* __restrict__ ts, const int64_t* __restrict__ pose_ts,
                                const float* __restrict__ pose_pos, const float* __restrict__ pose_quat,
                                int32_t n_poses, const float* __restrict__ rays_d,
                                const float* __restrict__ d_rays_o, const float* __restrict__ d_rays_d,
                                double* __restrict__ d_ts, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double t = ts[i];
        int lo = 0, hi = n_poses;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((double)pose_ts[mid] < t) lo = mid + 1; else hi = mid;
        }
        int right = lo;
        int left = (t == (double)pose_ts[0]) ? right : right - 1;
        left = min(max(left, 0), n_poses - 2);
        right = min(max(right, 1), n_poses - 1);
        const double width = (double)(pose_ts[left + 1] - pose_ts[left]);
        float g = 0.f;                                   // dL/dw
        if (d_rays_o != nullptr) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
                g = fmaf(d_rays_o[3 * i + d], pose_pos[3 * right + d] - pose_pos[3 * left + d], g);
        }
        if (d_rays_d != nullptr) {
            float q0[4], q1[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) { q0[d] = pose_quat[4 * left + d]; q1[d] = pose_quat[4 * right + d]; }
            const float dot = q0[0] * q1[0] + q0[1] * q1[1] + q0[2] * q1[2] + q0[3] * q1[3];
            if (dot < 0.f) {
#pragma unroll
                for (int d = 0; d < 4; ++d) q1[d] = -q1[d];
            }
            const float q0c[4] = {-q0[0], -q0[1], -q0[2], q0[3]};
            float rel[4];
            quat_mul(q0c, q1, rel);
            const float vnorm = sqrtf(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]);
            const float angle = 2.f * atan2f(vnorm, rel[3]);
            const bool small_a = fabsf(angle) <= 1e-3f;
            const float a2 = angle * angle;
            const float scale_v = small_a ? 2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f : angle / sinf(angle * 0.5f);
            // body-frame rotation vector r, turned into the world frame by q0: q0 (r, 0) q0^-1
            const float r[4] = {scale_v * rel[0], scale_v * rel[1], scale_v * rel[2], 0.f};
            float tmp[4], om[4];
            quat_mul(q0, r, tmp);
            quat_mul(tmp, q0c, om);
            const float inv_n2 = 1.f / (q0[0] * q0[0] + q0[1] * q0[1] + q0[2] * q0[2] + q0[3] * q0[3]);
            const float ox = om[0] * inv_n2, oy = om[1] * inv_n2, oz = om[2] * inv_n2;
            const float dx = rays_d[3 * i], dy = rays_d[3 * i + 1], dz = rays_d[3 * i + 2];
            g = fmaf(d_rays_d[3 * i], oy * dz - oz * dy, g);
            g = fmaf(d_rays_d[3 * i + 1], oz * dx - ox * dz, g);
            g = fmaf(d_rays_d[3 * i + 2], ox * dy - oy * dx, g);
        }
        d_ts[i] = (double)g / width;
    }
}

}  // namespace den

extern "C" int den_rays_from_trajectory_bwd(const double* timestamps, const int64_t* pose_ts,
                                            const float* pose_pos, const float* pose_quat, int32_t n_poses,
                                            const float* rays_d, const float* d_rays_o, const float* d_rays_d,
                                            double* d_timestamps, int64_t n_rays, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0, "bad sizes");
    DEN_CHECK_ARG(n_poses >= 2, "at least two poses");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(timestamps && pose_ts && pose_pos && pose_quat && d_timestamps, "null pointer");
    DEN_CHECK_ARG(d_rays_d == nullptr || rays_d != nullptr, "the direction gradient needs the forward directions");
    rays_from_trajectory_bwd_kernel<<<grid_for(n_rays, 256, 8), 256, 0, as_stream(stream)>>>(
        timestamps, pose_ts, pose_pos, pose_quat, n_poses, rays_d, d_rays_o, d_rays_d, d_timestamps, n_rays);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}

extern "C" int den_rays_from_trajectory(const double* timestamps, const float* pixels, int64_t n_pixels,
                                        const int64_t* pose_ts, const float* pose_pos,
                                        const float* pose_quat, int32_t n_poses, const float* kinv9_host,
                                        float* rays_o, float* rays_d, int64_t n_rays, void* stream) {
    using namespace den;
    DEN_CHECK_ARG(n_rays >= 0 && n_pixels > 0, "bad sizes");
    DEN_CHECK_ARG(n_poses >= 2, "at least two poses");
    if (n_rays == 0) return DEN_OK;
    DEN_CHECK_ARG(timestamps && pixels && pose_ts && pose_pos && pose_quat && kinv9_host && rays_o && rays_d,
                  "null pointer");
    const float* k = kinv9_host;
    rays_from_trajectory_kernel<<<grid_for(n_rays, 256, 8), 256, 0, as_stream(stream)>>>(
        timestamps, pixels, n_pixels, pose_ts, pose_pos, pose_quat, n_poses, k[0], k[1], k[2], k[3], k[4], k[5],
        k[6], k[7], k[8], rays_o, rays_d, n_rays);
    DEN_CHECK_LAUNCH();
    return DEN_OK;
}
