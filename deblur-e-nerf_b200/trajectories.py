"""B2 host mirror of ``LinearTrajectory`` (models/trajectories.py:8-90) with the
reference's fixed SLERP (utils/tensor_ops.py:87-184) and the five RoMa 1.2.7 functions it
uses: timestamp (float64 ns) -> camera position (LERP) and rotation matrix (shortest-path
SLERP through the rotation vector).

The autograd (torch) form below carries the tau-gradient path (timestamps require grad
only when the refractory period is optimised, config 4); when they do not, the fused
``den_rays_from_trajectory`` kernel turns timestamps + pixels straight into rays."""

import torch


def _quat_conj(q):
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def _quat_mul(p, q):
    pv, pw, qv, qw = p[..., :3], p[..., 3:], q[..., :3], q[..., 3:]
    return torch.cat((pw * qv + qw * pv + torch.cross(pv, qv, dim=-1),
                      pw * qw - torch.sum(pv * qv, dim=-1, keepdim=True)), dim=-1)


def _full_rotvec(q):
    vec = q[..., :3]
    angle = 2 * torch.atan2(torch.norm(vec, dim=-1), q[..., 3])
    small = angle.abs() <= 1e-3
    safe = torch.where(small, torch.ones_like(angle), angle)
    scale = torch.where(small, 2 + angle ** 2 / 12 + 7 * angle ** 4 / 2880,
                        safe / torch.sin(safe / 2))
    return scale[..., None] * vec


def _rotvec_to_quat(r):
    theta = torch.norm(r, dim=-1)
    small = theta <= 1e-3
    safe = torch.where(small, torch.ones_like(theta), theta)
    scale = torch.where(small, 0.5 - theta ** 2 / 48 + theta ** 4 / 3840,
                        torch.sin(safe / 2) / safe)
    return torch.cat((scale[..., None] * r, torch.cos(theta / 2)[..., None]), dim=-1)


def quat_to_rotmat(q):
    x, y, z, w = q.unbind(dim=-1)
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    return torch.stack((
        torch.stack((x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw)), dim=-1),
        torch.stack((2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw)), dim=-1),
        torch.stack((2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2), dim=-1)), dim=-2)


def slerp(q0, q1, w):
    q1 = torch.where(torch.sum(q0 * q1, dim=-1, keepdim=True) < 0, -q1, q1)
    rel = _quat_mul(_quat_conj(q0), q1)
    return _quat_mul(q0, _rotvec_to_quat(w[..., None] * _full_rotvec(rel)))


class LinearTrajectory(torch.nn.Module):
    def __init__(self, camera_poses):
        """`camera_poses`: the reference's CameraPose dataset (has `.camera_poses` with
        T_wc_position / T_wc_orientation / T_wc_timestamp) or a (pos, quat, ts) tuple."""
        super().__init__()
        if isinstance(camera_poses, (tuple, list)):
            pos, quat, ts = camera_poses
        else:
            cp = camera_poses.camera_poses
            pos, quat, ts = cp["T_wc_position"], cp["T_wc_orientation"], cp["T_wc_timestamp"]
        self.register_buffer("T_wc_position", pos, persistent=False)
        self.register_buffer("T_wc_orientation_quat", quat, persistent=False)
        self.register_buffer("T_wc_timestamp", ts.contiguous(), persistent=False)
        self.register_buffer("bin_width", self.T_wc_timestamp.diff(), persistent=False)

    def forward(self, input_timestamp):
        right = torch.searchsorted(self.T_wc_timestamp, input_timestamp.contiguous())
        left = torch.where(input_timestamp == self.T_wc_timestamp[0], right, right - 1)
        # the reference asserts the bins are in range (a host sync); the kernels clamp
        left = left.clamp(0, len(self.T_wc_timestamp) - 2)
        right = right.clamp(1, len(self.T_wc_timestamp) - 1)
        w = ((input_timestamp - self.T_wc_timestamp[left]) / self.bin_width[left]).to(
            self.T_wc_position.dtype)
        pos = torch.lerp(self.T_wc_position[left], self.T_wc_position[right], w[..., None])
        quat = slerp(self.T_wc_orientation_quat[left], self.T_wc_orientation_quat[right], w)
        return pos, quat_to_rotmat(quat)
