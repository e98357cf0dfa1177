"""oracle/fk_lbs.py - pure-torch CPU restatement of SK_GS's skeleton FK + linear-blend skinning.

TEST INFRASTRUCTURE ONLY (see oracle/raster_oracle.c header): only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg import this module.  The product (sk_gs_b200/) never does.

Parity status: "parity unpinned" at the lietorch / pytorch3d boundary (both un-vendored).  What IS
pinned: FK against the reference's own matrix formulation `skeleton_warp` / `skeleton_warp_v0` and `find_root`
(imported from /root/reference by tests/golden/make_golden.py -> tests/golden/fk_*.npz).

Everything is written in plain quaternion algebra, quaternions are (x, y, z, w) as inside SK_GS.
All /root/reference citations are `networks/sk_gs.py` unless stated.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor


# ----------------------------------------------------------------------------------------------- quaternion algebra
def q_normalize(q: Tensor) -> Tensor:
    """lietorch normalises in every SO3 constructor (my_ext/_C/include/lie.h:45-47)."""
    return q / q.norm(dim=-1, keepdim=True)


def q_mul(a: Tensor, b: Tensor) -> Tensor:
    """Hamilton product, xyzw (my_ext/ops_3d/quaternion.py:44-49)."""
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ], dim=-1)


def q_rotate(q: Tensor, p: Tensor) -> Tensor:
    """p + w*2(v x p) + v x 2(v x p) (lie.h:59-64); q must be unit."""
    v, w = q[..., :3], q[..., 3:]
    v, p = torch.broadcast_tensors(v, p)
    uv = torch.cross(v, p, dim=-1)
    uv = uv + uv
    return p + w * uv + torch.cross(v, uv, dim=-1)


def so3_exp(phi: Tensor) -> Tensor:
    """lie.h:142-159, Taylor branch for theta < 1e-6."""
    theta2 = (phi * phi).sum(-1, keepdim=True)
    theta = theta2.sqrt()
    small = theta < 1e-6
    theta_safe = torch.where(small, torch.ones_like(theta), theta)
    imag = torch.where(small, 0.5 - theta2 / 48.0 + theta2 * theta2 / 3840.0, torch.sin(0.5 * theta_safe) / theta_safe)
    real = torch.where(small, 1.0 - theta2 / 8.0 + theta2 * theta2 / 384.0, torch.cos(0.5 * theta))
    return q_normalize(torch.cat([imag * phi, real], dim=-1))


def se3_mul(a: Tensor, b: Tensor) -> Tensor:
    """(t1,q1)o(t2,q2) = (t1 + R(q1) t2, q1 q2) with the product quaternion re-normalised (lie.h:242-244, 45, 55-57)."""
    qa, qb = q_normalize(a[..., 3:]), q_normalize(b[..., 3:])
    return torch.cat([a[..., :3] + q_rotate(qa, b[..., :3]), q_normalize(q_mul(qa, qb))], dim=-1)


def se3_act(T: Tensor, p: Tensor) -> Tensor:
    """lie.h:246."""
    return q_rotate(q_normalize(T[..., 3:]), p) + T[..., :3]


# --------------------------------------------------------------------------------------------------------- tree
def build_tree(parent: Tensor) -> Tuple[Tensor, Tensor, int]:
    """Restatement of `find_root` (:50-103): re-root at the tree centre, binary-lifting table parents[M, L].

    parent[j] = index of j's parent in the input tree, -1 for its root.  Unlike the reference (IndexError for L == 0,
    SURVEY App. A.1) a tree of radius 1 yields L = 1.
    """
    M = parent.shape[0]
    par = [int(v) for v in parent.tolist()]
    edges = [[] for _ in range(M)]
    for i, j in enumerate(par):
        if j >= 0:
            edges[i].append(j)
            edges[j].append(i)
    num_edges = [len(e) for e in edges]
    visited = [0] * M
    que = [i for i in range(M) if num_edges[i] == 1]
    for n in que:
        visited[n] = 1
    if not que:  # single joint
        que = [0]
    i = 0
    while i < len(que):
        now = que[i]
        i += 1
        for node in edges[now]:
            if num_edges[node] > 1:
                num_edges[node] -= 1
                visited[node] = max(visited[node], visited[now] + 1)
                if num_edges[node] == 1:
                    que.append(node)
    root = que[-1]
    max_depth = max(visited) if M > 1 else 1
    L = 0
    while 2 ** L < max_depth:
        L += 1
    L = max(L, 1)
    parents = torch.full((M, L), root, dtype=torch.int64)
    depth = torch.zeros(M, dtype=torch.int64)
    seen = [False] * M
    seen[root] = True
    que = [root]
    i = 0
    while i < len(que):
        now = que[i]
        i += 1
        for node in edges[now]:
            if not seen[node]:
                parents[node, 0] = now
                depth[node] = depth[now] + 1
                que.append(node)
                seen[node] = True
    for lv in range(1, L):
        for j in range(M):
            parents[j, lv] = parents[parents[j, lv - 1], lv - 1]
    return parents, depth, root


# ----------------------------------------------------------------------------------------------------------- FK
def local_transforms(joints: Tensor, sk_r: Tensor, sk_r_delta: Optional[Tensor] = None) -> Tensor:
    """`kinematic` post-MLP part (:1086-1091): L_a = (j_a - R(r_a) j_a, r_a), optional repose delta (:1087-1088)."""
    r = q_normalize(sk_r)
    if sk_r_delta is not None:
        d = so3_exp(sk_r_delta) if sk_r_delta.shape[-1] == 3 else q_normalize(sk_r_delta)
        r = q_normalize(q_mul(d, r))
    t = joints + q_rotate(r, -joints)
    return torch.cat([t, r], dim=-1)


def skeleton_warp_jump(local_T: Tensor, g_tr: Optional[Tensor], parents: Tensor, root: int) -> Tensor:
    """`skeleton_warp_SE3` (:193-206): L rounds of pointer jumping, then left-multiply by the global transform."""
    out = local_T.clone()
    ident = out.new_tensor([0, 0, 0, 0, 0, 0, 1.0])
    mask = torch.zeros(out.shape[0], 1, dtype=torch.bool, device=out.device)
    mask[root] = True
    out = torch.where(mask, ident, out)
    for lv in range(parents.shape[1]):
        out = se3_mul(out[parents[:, lv]], out)
    if g_tr is None:
        return out
    return se3_mul(g_tr.view(1, 7).expand_as(out), out)


def skeleton_warp_serial(local_T: Tensor, g_tr: Optional[Tensor], parent0: Tensor, root: int) -> Tensor:
    """Serial recurrence T_root = g, T_a = T_parent(a) o L_a (cf. `skeleton_warp_v0` :167-179); cross-check only."""
    M = local_T.shape[0]
    ident = local_T.new_tensor([0, 0, 0, 0, 0, 0, 1.0])
    out = [None] * M
    out[root] = ident if g_tr is None else se3_mul(g_tr.view(7), ident)
    order, seen = [root], {root}
    children = {i: [] for i in range(M)}
    for j in range(M):
        if j != root:
            children[int(parent0[j])].append(j)
    i = 0
    while i < len(order):
        for c in children[order[i]]:
            if c not in seen:
                seen.add(c)
                order.append(c)
        i += 1
    for a in order[1:]:
        out[a] = se3_mul(out[int(parent0[a])], local_T[a])
    return torch.stack(out, 0)


# ------------------------------------------------------------------------------------------------- skinning weights
def knn_joints(points: Tensor, joints: Tensor, K: int) -> Tuple[Tensor, Tensor]:
    """pytorch3d knn_points semantics (:757): K smallest SQUARED distances, ascending; ties -> lower index."""
    diff = points[:, None, :] - joints[None, :, :]
    d2 = diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1] + diff[..., 2] * diff[..., 2]
    order = torch.argsort(d2, dim=1, stable=True)[:, :K]
    return torch.gather(d2, 1, order), order


def lbs_weights(points: Tensor, joints: Tensor, K: int, mode: str, sp_W: Optional[Tensor] = None,
                sp_radius: Optional[Tensor] = None, sp_weight: Optional[Tensor] = None, temperature: float = 1.0):
    """`calc_LBS_weight` (:751-774).  points are detached (:1113); joints are not.

    mode 'W': softmax(gather(sp_W, idx)) (:767-768); 'kernel' / 'weighted_kernel': exp(-d2 / (2 r^2)) [* sigmoid(w)] + 1e-7,
    L1-normalised (:760-766) with r = exp(_sp_radius), (:548-553); 'dist': softmax(-d2 / temperature) (:769-770).
    """
    points = points.detach()
    with torch.no_grad():
        _, idx = knn_joints(points, joints, K)
    diff = points[:, None, :] - joints[idx]
    d2 = (diff * diff).sum(-1)
    if mode in ('kernel', 'weighted_kernel'):
        radius = torch.exp(sp_radius)[idx]
        w = torch.exp(-d2 / (2 * radius ** 2))
        if mode == 'weighted_kernel':
            w = w * torch.sigmoid(sp_weight)[idx]
        w = w + 1e-7
        w = w / w.sum(dim=-1, keepdim=True)
    elif mode == 'W':
        w = torch.gather(sp_W, 1, idx).softmax(dim=-1)
    elif mode == 'dist':
        w = torch.softmax(-d2 / temperature, dim=-1)
    else:
        raise ValueError(mode)
    return w, idx


# -------------------------------------------------------------------------------------------------------- sk_stage
def sk_stage(xyz: Tensor, joints: Tensor, sk_r: Tensor, sk_d_rot: Tensor, sk_d_scale: Tensor, g_tr: Optional[Tensor],
             parents: Tensor, root: int, K: int = 5, mode: str = 'W', sp_W: Optional[Tensor] = None,
             sp_radius: Optional[Tensor] = None, sp_weight: Optional[Tensor] = None, temperature: float = 1.0,
             sk_r_delta: Optional[Tensor] = None):
    """`sk_stage` (:1109-1150) minus the MLP: returns the same 9-tuple
    (d_xyz, d_rot, d_scale, sk_T[M,7], sk_d_rot, sk_d_scale, g_tr, weights, indices)."""
    points = xyz.detach()
    sk_T = skeleton_warp_jump(local_transforms(joints, sk_r, sk_r_delta), g_tr, parents, root)
    w, idx = lbs_weights(points, joints, K, mode, sp_W, sp_radius, sp_weight, temperature)
    d_xyz = (se3_act(sk_T[idx], points[:, None, :]) * w[..., None]).sum(dim=1) - points
    d_rot = (sk_d_rot[idx] * w[..., None]).sum(dim=1)
    d_scale = (sk_d_scale[idx] * w[..., None]).sum(dim=1)
    return d_xyz, d_rot, d_scale, sk_T, sk_d_rot, sk_d_scale, g_tr, w, idx


# -------------------------------------------------------------------------------------------------------- sp_stage
def sp_warp(points: Tensor, sp_points: Tensor, sp_t: Tensor, sp_r: Tensor, sp_rot: Optional[Tensor],
            sp_scale: Optional[Tensor], weights: Tensor, indices: Tensor, method: str = 'LBS'):
    """`warp` (:776-828), tensor branch of sp_t (:794-799), SE3 branch of spT (:807-812).

    lietorch (un-vendored, "parity unpinned" for its internals) supplies SE3.act(p) = R(q) p + t (lie.h:246, q used as
    is) and returns gradients w.r.t. the 7-vector projected onto the tangent space at (t, q) (FromVec backward: tangent
    gradient times pinv of the orthogonal projector, whose in-tree copy is my_ext/_C/include/lie.h:82-90,303-311).  Both are reproduced by evaluating the action with q / |q|: equal
    values for the unit quaternions `sp_stage` passes (:847), and autograd through the normalisation IS that projection.
    Pinned on the reference's own `warp` code run with a functional lietorch stand-in: tests/golden/sp_stage.npz."""
    q = q_normalize(sp_r)
    t = sp_t
    if method == 'LBS_c':
        t = sp_t + sp_points + q_rotate(q, -sp_points)
    spT = torch.cat([t, q], dim=-1)
    if method in ('LBS', 'LBS_c'):
        d_points = (se3_act(spT[indices], points[:, None, :]) * weights[..., None]).sum(dim=1) - points
    elif method == 'largest':
        p2sp = torch.gather(indices, -1, weights.argmax(dim=-1, keepdim=True))[:, 0]  # :850-851
        d_points = se3_act(spT[p2sp], points) - points
    else:
        raise ValueError(method)
    rot = sp_rot if sp_rot is not None else sp_r
    d_rotation = (rot[indices] * weights[..., None]).sum(dim=1)
    d_scales = (sp_scale[indices] * weights[..., None]).sum(dim=1) if sp_scale is not None else None
    return d_points, d_rotation, d_scales, spT


def sp_stage(points: Tensor, sp_points: Tensor, sp_t: Tensor, sp_r: Tensor, sp_rot: Optional[Tensor],
             sp_scale: Optional[Tensor], K: int = 5, mode: str = 'W', sp_W: Optional[Tensor] = None,
             sp_radius: Optional[Tensor] = None, sp_weight: Optional[Tensor] = None, temperature: float = 1.0,
             method: str = 'LBS'):
    """`sp_stage` (:830-856) minus the deformation network: weights (:843) then warp (:852-855).
    -> (d_points, d_rotation, d_scales, spT, weights, indices)."""
    points = points.detach()
    w, idx = lbs_weights(points, sp_points, K, mode, sp_W, sp_radius, sp_weight, temperature)
    return sp_warp(points, sp_points, sp_t, sp_r, sp_rot, sp_scale, w, idx, method) + (w, idx)


def assemble(_xyz: Tensor, _scaling: Tensor, _rotation: Tensor, _opacity: Tensor, f_dc: Tensor, f_rest: Tensor,
             d_xyz: Tensor, d_rot: Tensor, d_scale: Tensor):
    """Output assembly of `forward` (:1162-1163,1192,1202-1203) with the activations of
    networks/gaussian_splatting.py:155-160: points, scales, rotations (xyzw, F.normalize eps 1e-12), opacity, sh."""
    points = _xyz + d_xyz
    scales = torch.exp(_scaling) + d_scale
    rotations = torch.nn.functional.normalize(_rotation + d_rot)
    opacity = torch.sigmoid(_opacity)
    sh = torch.cat((f_dc, f_rest), dim=1)
    return points, scales, rotations, opacity, sh


__all__ = [name for name in dir() if not name.startswith('_') and name not in ('math', 'torch', 'Tensor', 'Optional', 'Tuple')]
