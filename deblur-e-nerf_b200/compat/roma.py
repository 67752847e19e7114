"""The `roma` (RoMa 1.2.7) functions the reference calls — `utils/tensor_ops.py:76,98,115,169-179`,
`models/trajectories.py:86` — on XYZW unit quaternions."""

import types

import torch


def _flatten(t, end_dim):
    lead = t.shape[:t.dim() + end_dim + 1]
    return t.reshape(-1, *t.shape[t.dim() + end_dim + 1:]), lead


def _unflatten(t, lead):
    return t.reshape(*lead, *t.shape[1:])


internal = types.SimpleNamespace(flatten_batch_dims=_flatten, unflatten_batch_dims=_unflatten)


def quat_conjugation(q):
    return q * q.new_tensor([-1.0, -1.0, -1.0, 1.0])


def quat_product(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack((aw * bx + bw * ax + (ay * bz - az * by),
                        aw * by + bw * ay + (az * bx - ax * bz),
                        aw * bz + bw * az + (ax * by - ay * bx),
                        aw * bw - (ax * bx + ay * by + az * bz)), dim=-1)


def rotvec_to_unitquat(rotvec):
    angle = rotvec.norm(dim=-1, keepdim=True)
    tiny = angle <= 1e-3
    a2 = angle * angle
    series = 0.5 - a2 / 48 + a2 * a2 / 3840                 # sin(a / 2) / a around 0
    ratio = torch.where(tiny, series, torch.sin(angle / 2) / torch.where(tiny, torch.ones_like(angle), angle))
    return torch.cat((ratio * rotvec, torch.cos(angle / 2)), dim=-1)


def unitquat_to_rotvec(q):
    q = torch.where(q[..., 3:] < 0, -q, q)
    s = q[..., :3].norm(dim=-1, keepdim=True)
    angle = 2 * torch.atan2(s, q[..., 3:])
    tiny = s <= 1e-3
    a2 = angle * angle
    series = 2 + a2 / 12 + 7 * a2 * a2 / 2880
    ratio = torch.where(tiny, series, angle / torch.where(tiny, torch.ones_like(s), torch.sin(angle / 2)))
    return ratio * q[..., :3]


def unitquat_to_rotmat(q):
    x, y, z, w = q.unbind(-1)
    return torch.stack((
        torch.stack((w * w + x * x - y * y - z * z, 2 * (x * y - z * w), 2 * (x * z + y * w)), -1),
        torch.stack((2 * (x * y + z * w), w * w - x * x + y * y - z * z, 2 * (y * z - x * w)), -1),
        torch.stack((2 * (x * z - y * w), 2 * (y * z + x * w), w * w - x * x - y * y + z * z), -1)), -2)


def unitquat_slerp(q0, q1, steps, shortest_path=False):
    """steps (S,) -> (S, ..., 4): q0 * exp(step * log(q0^-1 q1))."""
    rel = quat_product(quat_conjugation(q0), q1)
    if shortest_path:
        rel = torch.where(rel[..., 3:] < 0, -rel, rel)
    rotvec = unitquat_to_rotvec(rel)
    shape = (-1,) + (1,) * rotvec.dim()
    return quat_product(q0.expand(len(steps), *q0.shape),
                        rotvec_to_unitquat(steps.reshape(shape) * rotvec.unsqueeze(0)))
