"""Synthetic SK_GS-shaped scenes: skeleton, canonical Gaussians, skinning parameters and cameras.

Everything is generated on the CPU with a seeded torch.Generator (seed 20241017 + config index) so that the tests, the
CPU oracle and every GPU rank see identical inputs.  Conventions follow the reference:
  * quaternions xyzw, `_scaling` is log-scale, `_opacity` a logit, SH as f_dc[P,1,3] | f_rest[P,15,3]
    (/root/reference/networks/gaussian_splatting.py:155-160,185-200),
  * cameras are OpenCV/colmap (x right, y down, z forward), `Tv2c` as my_ext/ops_3d/coord_trans_opencv.py:203-239,
    raster settings as networks/gaussian_splatting.py:271-284 (viewmatrix = Tw2v.T, projmatrix = (Tv2c @ Tw2v).T).
The binary-lifting joint table is built host-side exactly like `find_root` (networks/sk_gs.py:50-103).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

SMPL24_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


@dataclass
class SceneConfig:
    name: str
    P: int
    M: int
    W: int
    H: int
    views: int
    backward: bool = True
    s0: float = 0.012  # median Gaussian scale (world units); tuned so mean tiles_touched is about 8
    smpl: bool = False
    idx: int = 0


# BASELINE.json `configs`, in order.
CONFIGS: Dict[str, SceneConfig] = {
    'c1': SceneConfig('c1-dnerf-hook-10k-400', 10_000, 16, 400, 400, 1, s0=0.024, idx=0),
    'c2': SceneConfig('c2-dnerf-100k-800', 100_000, 32, 800, 800, 1, s0=0.012, idx=1),
    'c3': SceneConfig('c3-wim512-200k-512x8', 200_000, 32, 512, 512, 8, s0=0.019, idx=2),
    'c4': SceneConfig('c4-zju-300k-1024x4', 300_000, 24, 1024, 1024, 4, s0=0.0095, smpl=True, idx=3),
    'c5': SceneConfig('c5-repose-3m-1080p-x64', 3_000_000, 64, 1920, 1080, 64, backward=False, s0=0.0058, idx=4),
    # the north-star headline shape: 300K Gaussians at 800x800 (BASELINE.json north_star "Target")
    'ns': SceneConfig('ns-300k-800', 300_000, 32, 800, 800, 1, s0=0.012, idx=5),
}


def find_root_table(parent: List[int]):
    """Host-side joint table: (parents[M, L] int32 binary lifting, depth[M] int32, root).

    Same construction as the reference's `find_root` (networks/sk_gs.py:50-103): peel leaves to find the tree centre,
    re-root there, L = ceil(log2(radius + 1)) levels, parents[root, :] = root.  Radius-1 trees get L = 1 (the
    reference raises IndexError there)."""
    M = len(parent)
    edges = [[] for _ in range(M)]
    for i, j in enumerate(parent):
        if j >= 0:
            edges[i].append(j)
            edges[j].append(i)
    deg = [len(e) for e in edges]
    level = [0] * M
    que = [i for i in range(M) if deg[i] == 1] or [0]
    for n in que:
        level[n] = 1
    i = 0
    while i < len(que):
        now = que[i]
        i += 1
        for nb in edges[now]:
            if deg[nb] > 1:
                deg[nb] -= 1
                level[nb] = max(level[nb], level[now] + 1)
                if deg[nb] == 1:
                    que.append(nb)
    root = que[-1]
    max_depth = max(level) if M > 1 else 1
    L = 0
    while 2 ** L < max_depth:
        L += 1
    L = max(L, 1)
    parents = torch.full((M, L), root, dtype=torch.int32)
    depth = torch.zeros(M, dtype=torch.int32)
    seen = [False] * M
    seen[root] = True
    que = [root]
    i = 0
    while i < len(que):
        now = que[i]
        i += 1
        for nb in edges[now]:
            if not seen[nb]:
                seen[nb] = True
                parents[nb, 0] = now
                depth[nb] = depth[now] + 1
                que.append(nb)
    for lv in range(1, L):
        for j in range(M):
            parents[j, lv] = parents[int(parents[j, lv - 1]), lv - 1]
    return parents, depth, root


def perspective_opencv(fovy: float, W: int, H: int, n: float = 0.01, f: float = 1000.0) -> torch.Tensor:
    y = math.tan(fovy * 0.5)
    x = y * (W / H)
    T = torch.zeros(4, 4)
    T[0, 0] = 1.0 / x
    T[1, 1] = 1.0 / y
    T[3, 2] = 1.0
    T[2, 2] = (f + n) / (f - n)
    T[2, 3] = -(2 * f * n) / (f - n)
    return T


def look_at_opencv(eye: torch.Tensor, at: torch.Tensor) -> torch.Tensor:
    """World->view, OpenCV axes (x right, y down, z forward)."""
    z = torch.nn.functional.normalize(at - eye, dim=0)
    up = torch.tensor([0.0, 0.0, 1.0])
    x = torch.nn.functional.normalize(torch.linalg.cross(z, up), dim=0)
    y = torch.linalg.cross(z, x)
    R = torch.stack([x, y, z], 0)
    T = torch.eye(4)
    T[:3, :3] = R
    T[:3, 3] = -R @ eye
    return T


@dataclass
class Camera:
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    Tw2v: torch.Tensor
    Tv2c: torch.Tensor
    campos: torch.Tensor
    bg: torch.Tensor

    @property
    def viewmatrix(self):
        return self.Tw2v.t().contiguous()

    @property
    def projmatrix(self):
        return (self.Tv2c @ self.Tw2v).t().contiguous()


@dataclass
class Scene:
    cfg: SceneConfig
    # skeleton
    joints: torch.Tensor  # [M,3]
    parents: torch.Tensor  # [M,L] int32
    joint_depth: torch.Tensor
    root: int
    sk_r: torch.Tensor  # [M,4] unit xyzw
    sk_d_rot: torch.Tensor  # [M,4]
    sk_d_scale: torch.Tensor  # [M,3]
    g_tr: torch.Tensor  # [7] (t, q xyzw)
    # canonical Gaussians
    xyz: torch.Tensor
    scaling: torch.Tensor  # log-scale
    rotation: torch.Tensor
    opacity: torch.Tensor  # logit [P,1]
    f_dc: torch.Tensor  # [P,1,3]
    f_rest: torch.Tensor  # [P,15,3]
    # skinning
    sp_W: torch.Tensor  # [P,M]
    sp_radius: torch.Tensor  # [M] log radius
    sp_weight: torch.Tensor  # [M]
    cameras: List[Camera] = field(default_factory=list)
    K: int = 5
    sh_degree: int = 3


def make_scene(cfg, views: Optional[int] = None, seed: Optional[int] = None, P: Optional[int] = None) -> Scene:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    g = torch.Generator().manual_seed(20241017 + cfg.idx if seed is None else seed)
    M = cfg.M
    P = cfg.P if P is None else P
    V = cfg.views if views is None else views

    def randn(*s):
        return torch.randn(*s, generator=g)

    def rand(*s):
        return torch.rand(*s, generator=g)

    # ---- skeleton
    if cfg.smpl:
        parent = list(SMPL24_PARENTS)
    else:
        parent = [-1] + [int(torch.randint(0, j, (1,), generator=g)) for j in range(1, M)]
    joints = torch.zeros(M, 3)
    for j in range(1, M):
        joints[j] = joints[parent[j]] + 0.25 * randn(3)
    joints = joints.clamp(-1.3, 1.3)
    parents, depth, root = find_root_table(parent)
    aa = randn(M, 3) * math.radians(20.0)
    th = aa.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    sk_r = torch.cat([aa / th * torch.sin(0.5 * th), torch.cos(0.5 * th)], -1)
    sk_d_rot = 0.01 * randn(M, 4)
    sk_d_scale = 0.001 * randn(M, 3)
    gq = torch.nn.functional.normalize(torch.tensor([0.0, 0.0, 0.0, 1.0]) + 0.05 * randn(4), dim=0)
    g_tr = torch.cat([0.05 * randn(3), gq])
    # ---- Gaussians
    owner = torch.randint(0, M, (P,), generator=g)
    xyz = joints[owner] + 0.15 * randn(P, 3)
    scaling = math.log(cfg.s0) + 0.5 * randn(P, 3)
    rotation = torch.nn.functional.normalize(randn(P, 4), dim=-1)
    opacity = 2.0 * randn(P, 1)
    f_dc = ((rand(P, 1, 3) - 0.5) / 0.28209479177387814)
    f_rest = 0.05 * randn(P, 15, 3)
    sp_W = randn(P, M)
    rng = float((joints.max(0).values - joints.min(0).values).max())
    sp_radius = torch.full((M,), math.log(0.1 * max(rng, 1e-3)))
    sp_weight = torch.zeros(M)
    # ---- cameras
    cams = []
    fovx = 0.6911
    tanx = math.tan(0.5 * fovx)
    tany = tanx * cfg.H / cfg.W
    fovy = 2 * math.atan(tany)
    for _ in range(V):
        az = float(rand(1)) * 2 * math.pi
        el = math.radians(float(rand(1)) * 60.0 - 30.0)
        eye = 4.0 * torch.tensor([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
        Tw2v = look_at_opencv(eye, torch.zeros(3))
        cams.append(Camera(cfg.W, cfg.H, tanx, tany, Tw2v, perspective_opencv(fovy, cfg.W, cfg.H), eye.clone(),
                           torch.ones(3)))
    return Scene(cfg, joints, parents, depth, root, sk_r, sk_d_rot, sk_d_scale, g_tr, xyz, scaling, rotation, opacity,
                 f_dc, f_rest, sp_W, sp_radius, sp_weight, cams)
