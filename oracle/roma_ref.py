"""Oracle restatement of the five ``roma==1.2.7`` functions the hot path uses
(TEST INFRASTRUCTURE).  Quaternions are XYZW.

Reference call sites: ``utils/tensor_ops.py:98,115,169,176,179`` and
``models/trajectories.py:86``.  Upstream: ``roma/utils.py`` (``quat_product``,
``quat_conjugation``), ``roma/mappings.py`` (``rotvec_to_unitquat``,
``unitquat_to_rotmat``), ``roma/internal.py`` (batch-dim helpers).
"""

import types

import torch


def _flatten_batch_dims(tensor, end_dim):
    """Collapse all leading dims up to ``end_dim`` (inclusive, negative)."""
    batch_shape = tensor.shape[: end_dim + 1]
    flat = tensor.reshape((-1,) + tuple(tensor.shape[end_dim + 1:]))
    return flat, batch_shape


def _unflatten_batch_dims(tensor, batch_shape):
    return tensor.reshape(tuple(batch_shape) + tuple(tensor.shape[1:]))


internal = types.SimpleNamespace(
    flatten_batch_dims=_flatten_batch_dims,
    unflatten_batch_dims=_unflatten_batch_dims,
)


def quat_conjugation(quat):
    return torch.cat((-quat[..., :3], quat[..., 3:]), dim=-1)


def quat_product(p, q):
    pv, pw = p[..., :3], p[..., 3:]
    qv, qw = q[..., :3], q[..., 3:]
    vector = pw * qv + qw * pv + torch.cross(pv, qv, dim=-1)
    scalar = pw * qw - torch.sum(pv * qv, dim=-1, keepdim=True)
    return torch.cat((vector, scalar), dim=-1)


def rotvec_to_unitquat(rotvec):
    rotvec, batch_shape = _flatten_batch_dims(rotvec, end_dim=-2)
    theta = torch.norm(rotvec, dim=-1)
    small = theta <= 1e-3
    safe = torch.where(small, torch.ones_like(theta), theta)
    scale = torch.where(
        small,
        0.5 - theta ** 2 / 48 + theta ** 4 / 3840,
        torch.sin(safe / 2) / safe,
    )
    quat = torch.cat((scale[:, None] * rotvec, torch.cos(theta / 2)[:, None]), dim=-1)
    return _unflatten_batch_dims(quat, batch_shape)


def unitquat_to_rotmat(quat):
    x, y, z, w = quat.unbind(dim=-1)
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    rows = (
        torch.stack((x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw)), dim=-1),
        torch.stack((2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw)), dim=-1),
        torch.stack((2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2), dim=-1),
    )
    return torch.stack(rows, dim=-2)
