"""Deterministic synthetic scenes and cameras for parity tests and bench.py.

Follows SURVEY.md §8(d): superquadric "blocks" inside the unit sphere, point-level
surfels sampled on their surfaces, DTU-shaped pinhole cameras looking at the origin.
Camera matrices use the reference conventions (scene/cameras.py:54-63,
utils/graphics_utils.py:29-62): ``viewmatrix`` / ``projmatrix`` are the *transposed*
(row-vector) world->view and world->clip matrices.

Deviation from §8(d): the surfel scale is seeded from the *expected* 3-NN squared
distance of a uniform surface density (2A / (pi N)) instead of calling distCUDA2, so the
generator is pure CPU torch and produces identical tensors on every box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

SEED_BASE = 20251017

# name -> (P, W, H, views, S)   (BASELINE.json configs / SURVEY.md §8 table)
CONFIGS = {
    "C1": dict(P=20480, W=400, H=300, views=1, S=0),
    "C2": dict(P=300_000, W=400, H=300, views=49, S=0),
    "C3": dict(P=1_000_000, W=1600, H=1200, views=49, S=0),
    "C4": dict(P=500_000, W=800, H=600, views=1, S=16),
    "C5": dict(P=3_000_000, W=1920, H=1080, views=64, S=0),
}
CONFIG_INDEX = {"C1": 0, "C2": 1, "C3": 2, "C4": 3, "C5": 4}

C0 = 0.28209479177387814


def RGB2SH(rgb):
    return (rgb - 0.5) / C0


def spow(t, e):
    return torch.sign(t) * torch.abs(t).pow(e)


@dataclass
class Blocks:
    eps: torch.Tensor    # [B,2] raw (pre-sigmoid)
    s: torch.Tensor      # [B,3] raw log-scale
    r: torch.Tensor      # [B,4] raw quaternion (w,x,y,z)
    t: torch.Tensor      # [B,3]


def make_blocks(B: int, gen: torch.Generator) -> Blocks:
    eps = torch.rand(B, 2, generator=gen) * 4 - 2
    s = math.log(0.25) + 0.3 * torch.randn(B, 3, generator=gen)
    r = torch.randn(B, 4, generator=gen)
    t = torch.rand(B, 3, generator=gen) - 0.5
    return Blocks(eps, s, r, t)


def quat_to_mat(q):
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w,
        2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w,
        2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y], dim=-1).view(-1, 3, 3)


def mat_to_quat(R):
    """Rotation matrices [N,3,3] -> quaternions (w,x,y,z), w >= 0."""
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    qw = torch.sqrt(torch.clamp(1 + m00 + m11 + m22, min=0)) / 2
    qx = torch.sqrt(torch.clamp(1 + m00 - m11 - m22, min=0)) / 2
    qy = torch.sqrt(torch.clamp(1 - m00 + m11 - m22, min=0)) / 2
    qz = torch.sqrt(torch.clamp(1 - m00 - m11 + m22, min=0)) / 2
    qx = torch.copysign(qx, m21 - m12)
    qy = torch.copysign(qy, m02 - m20)
    qz = torch.copysign(qz, m10 - m01)
    q = torch.stack([qw, qx, qy, qz], dim=-1)
    return q / q.norm(dim=-1, keepdim=True)


def make_point_scene(P: int, seed: int, S: int = 0, n_blocks: int = 8, sh_degree: int = 3, device="cpu"):
    """Point-level surfel cloud on superquadric surfaces (C2/C3/C4 style)."""
    gen = torch.Generator().manual_seed(seed)
    blk = make_blocks(n_blocks, gen)
    eps = torch.sigmoid(blk.eps) * 1.8 + 0.1
    Sc = torch.exp(blk.s) + 0.05
    Rm = quat_to_mat(blk.r)

    bid = torch.randint(0, n_blocks, (P,), generator=gen)
    eta = (torch.rand(P, generator=gen) - 0.5) * math.pi
    omega = (torch.rand(P, generator=gen) * 2 - 1) * math.pi
    e1, e2 = eps[bid, 0], eps[bid, 1]
    ce, se = torch.cos(eta), torch.sin(eta)
    co, so = torch.cos(omega), torch.sin(omega)
    local = torch.stack([spow(ce, e1) * spow(so, e2), spow(se, e1), spow(ce, e1) * spow(co, e2)], dim=-1) * Sc[bid]
    nloc = torch.stack([spow(ce, 2 - e1) * spow(so, 2 - e2), spow(se, 2 - e1), spow(ce, 2 - e1) * spow(co, 2 - e2)],
                       dim=-1) / Sc[bid]
    nloc = nloc / (nloc.norm(dim=-1, keepdim=True) + 1e-12)
    Rb = Rm[bid]
    xyz = torch.einsum("nij,nj->ni", Rb, local) + blk.t[bid] + 1e-3 * torch.randn(P, 3, generator=gen)
    nrm = torch.einsum("nij,nj->ni", Rb, nloc)

    # tangent frame + random in-plane angle
    helper = torch.where(nrm[:, :1].abs() < 0.9, torch.tensor([[1.0, 0.0, 0.0]]), torch.tensor([[0.0, 1.0, 0.0]]))
    t1 = torch.linalg.cross(nrm, helper.expand_as(nrm))
    t1 = t1 / (t1.norm(dim=-1, keepdim=True) + 1e-12)
    t2 = torch.linalg.cross(nrm, t1)
    ang = torch.rand(P, generator=gen) * 2 * math.pi
    a1 = torch.cos(ang)[:, None] * t1 + torch.sin(ang)[:, None] * t2
    a2 = torch.linalg.cross(nrm, a1)
    Rs = torch.stack([a1, a2, nrm], dim=-1)  # columns: two tangents, normal
    rot = mat_to_quat(Rs) * torch.exp((torch.rand(P, 1, generator=gen) * 2 - 1) * math.log(2.0))

    # expected mean 3-NN squared distance for a uniform surface density
    area = float((4 * math.pi * (((Sc[:, 0] * Sc[:, 1]) ** 1.6 + (Sc[:, 0] * Sc[:, 2]) ** 1.6 +
                                  (Sc[:, 1] * Sc[:, 2]) ** 1.6) / 3).pow(1 / 1.6)).sum())
    dist2 = 2.0 * area / (math.pi * P)
    base = math.sqrt(dist2)
    scales = base * torch.exp((torch.rand(P, 2, generator=gen) * 2 - 1) * math.log(2.0))

    opacity = torch.sigmoid(1.0 + 1.5 * torch.randn(P, 1, generator=gen))
    M = (sh_degree + 1) ** 2
    shs = torch.zeros(P, M, 3)
    shs[:, 0] = RGB2SH(torch.rand(P, 3, generator=gen))
    if M > 1:
        shs[:, 1:] = 0.05 * torch.randn(P, M - 1, 3, generator=gen)
    out = dict(means3D=xyz.float(), scales=scales.float(), rotations=rot.float(), opacities=opacity.float(),
               shs=shs.float(), block_id=bid)
    if S > 0:
        out["semantics"] = torch.nn.functional.one_hot(bid % S, S).float()
    return {k: v.to(device).contiguous() for k, v in out.items()}


@dataclass
class Camera:
    image_width: int
    image_height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor    # [4,4] transposed world->view
    projmatrix: torch.Tensor    # [4,4] transposed world->clip
    campos: torch.Tensor        # [3]

    def to(self, device):
        return Camera(self.image_width, self.image_height, self.tanfovx, self.tanfovy,
                      self.viewmatrix.to(device), self.projmatrix.to(device), self.campos.to(device))

    # attribute names of the reference's scene/cameras.py:Camera (what render() and point_utils read)
    @property
    def world_view_transform(self):
        return self.viewmatrix

    @property
    def full_proj_transform(self):
        return self.projmatrix

    @property
    def camera_center(self):
        return self.campos

    @property
    def FoVx(self):
        return 2.0 * math.atan(self.tanfovx)

    @property
    def FoVy(self):
        return 2.0 * math.atan(self.tanfovy)


def projection_matrix(znear, zfar, tanx, tany):
    """utils/graphics_utils.py:42-62 (getProjectionMatrix) restated from tan(fov/2)."""
    top = tany * znear
    right = tanx * znear
    Pm = torch.zeros(4, 4, dtype=torch.float64)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def make_cameras(n: int, W: int, H: int, seed: int, radius: float = 2.5, device="cpu"):
    gen = torch.Generator().manual_seed(seed + 7919)
    fx = fy = 2892.0 * W / 1600.0
    tanx = W / (2 * fx)
    tany = H / (2 * fy)
    cams = []
    for i in range(n):
        az = (i + float(torch.rand(1, generator=gen))) / n * 2 * math.pi
        el = math.radians(-30 + 75 * float(torch.rand(1, generator=gen)))
        C = torch.tensor([radius * math.cos(el) * math.sin(az), -radius * math.sin(el),
                          radius * math.cos(el) * math.cos(az)], dtype=torch.float64)
        z = -C / C.norm()                      # forward: look at the origin
        up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)   # image y points down
        x = torch.linalg.cross(up, z)
        x = x / x.norm()
        y = torch.linalg.cross(z, x)
        Rcw = torch.stack([x, y, z], dim=0)    # world -> camera rotation
        W2C = torch.eye(4, dtype=torch.float64)
        W2C[:3, :3] = Rcw
        W2C[:3, 3] = -Rcw @ C
        view = W2C.float().t().contiguous()
        proj = projection_matrix(0.01, 100.0, tanx, tany).float().t().contiguous()
        full = (view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
        campos = view.inverse()[3, :3].contiguous()
        cams.append(Camera(W, H, tanx, tany, view, full, campos).to(device))
    return cams


def make_config(name: str, device="cpu", P: int | None = None, views: int | None = None):
    cfg = dict(CONFIGS[name])
    if P is not None:
        cfg["P"] = P
    if views is not None:
        cfg["views"] = views
    seed = SEED_BASE + CONFIG_INDEX[name]
    scene = make_point_scene(cfg["P"], seed, S=cfg["S"], device=device)
    cams = make_cameras(cfg["views"], cfg["W"], cfg["H"], seed, device=device)
    return cfg, scene, cams


def upstream_grads(W: int, H: int, seed: int, n_aux: int = 7, S: int = 0, device="cpu"):
    """Fixed dL/dcolor, dL/dallmap (and dL/dsemantic) ~ N(0,1)/Npix."""
    gen = torch.Generator().manual_seed(seed + 104729)
    npix = W * H
    g = dict(color=torch.randn(3, H, W, generator=gen) / npix, allmap=torch.randn(n_aux, H, W, generator=gen) / npix)
    if S > 0:
        g["semantic"] = torch.randn(S, H, W, generator=gen) / npix
    return {k: v.to(device) for k, v in g.items()}
