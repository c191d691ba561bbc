"""TEST INFRASTRUCTURE — CPU (torch, float32, autograd) restatement of PartGS's
superquadric -> surfel parameterisation.  Checker only; never imported by partgs_b200/.

Follows games/block_mesh_splatting/scene/block_gaussian_model.py:
  get_verts :189-193, prepare_scaling_rot :198-256, get_opacity :106-109,
  utils/superquadric.py:10-14 (parametric_sq), :93-101 (quaternion_to_rotation_matrix),
  utils/pytorch.py:28-29 (signed_pow), utils/general_utils.py:10-87 (rot_to_quat_batch).
Pinned against golden vectors produced by the reference Python itself
(tests/golden/sq2surfel_*.npz, made by tools/make_golden_sq.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def signed_pow(t, e):
    return torch.sign(t) * torch.abs(t).pow(e)


def quat_to_rotmat(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([
        1 - 2 * y ** 2 - 2 * z ** 2, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w,
        2 * x * y + 2 * z * w, 1 - 2 * x ** 2 - 2 * z ** 2, 2 * y * z - 2 * x * w,
        2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x ** 2 - 2 * y ** 2], dim=-1).view(-1, 3, 3)


def rotmat_to_quat(rot):
    m = rot.reshape(-1, 9)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
    x = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], dim=-1)
    q_abs = torch.zeros_like(x)
    pos = x > 0
    q_abs[pos] = torch.sqrt(x[pos])
    cand = torch.stack([
        torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].max(torch.tensor(0.1)))
    pick = F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    out = cand[pick, :].reshape(-1, 4)
    return torch.where(out[:, 0:1] < 0, -out, out)


def sq_to_surfels(sq_r, sq_s, sq_t, sq_eps, sq_occ, alpha, scale_raw, eta, omega, faces, ratio=0.25, smin=0.2,
                  eps=1e-8):
    """Returns (vertices, xyz, _scaling, _rotation, opacity) like the fused op."""
    B, Fn, _ = faces.shape
    K = alpha.shape[1]
    e = torch.sigmoid(sq_eps) * 1.8 + 0.1
    e1, e2 = e[:, 0:1], e[:, 1:2]
    ce, se = signed_pow(torch.cos(eta), e1), signed_pow(torch.sin(eta), e1)
    co, so = signed_pow(torch.cos(omega), e2), signed_pow(torch.sin(omega), e2)
    verts = torch.stack([ce * so, se, ce * co], dim=-1) * ratio
    S = torch.exp(sq_s) + smin
    Rm = quat_to_rotmat(F.normalize(sq_r))
    vertices = torch.bmm(verts * S.unsqueeze(1), Rm) + sq_t.unsqueeze(1)
    tri = vertices[torch.arange(B)[:, None, None], faces]            # [B,F,3,3]
    xyz = torch.matmul(alpha, tri.reshape(-1, 3, 3)).reshape(-1, 3)  # (b,f,k) row-major
    n = torch.linalg.cross(tri[:, :, 1] - tri[:, :, 0], tri[:, :, 2] - tri[:, :, 0], dim=2)
    v0 = n / (torch.linalg.vector_norm(n, dim=-1, keepdim=True) + eps)
    m = tri.mean(dim=2)
    a = tri[:, :, 1] - m
    la = torch.linalg.vector_norm(a, dim=-1, keepdim=True) + eps
    v1 = a / la
    b = tri[:, :, 2] - m
    w = b - (b * v0).sum(-1, keepdim=True) * v0 - (b * v1).sum(-1, keepdim=True) * v1
    v2 = w / (torch.linalg.vector_norm(w, dim=-1, keepdim=True) + eps)
    s1 = la / 2.0
    s2 = (b * v2).sum(-1, keepdim=True) / 2.0
    sc = torch.cat([s1, s2], dim=2).unsqueeze(2).expand(B, Fn, K, 2).reshape(B, Fn * K, 2)
    scaling = torch.log(torch.relu(scale_raw.reshape(B, Fn * K, 1) * sc) + eps).reshape(-1, 2)
    rot = torch.stack([v1, v2, v0], dim=2).unsqueeze(2).expand(B, Fn, K, 3, 3).reshape(-1, 3, 3).transpose(-2, -1)
    rotation = rotmat_to_quat(rot)
    opacity = torch.sigmoid(sq_occ).reshape(B, 1, 1).expand(B, Fn * K, 1).reshape(-1, 1)
    return vertices, xyz, scaling, rotation, opacity
