"""Independent dense float64 torch restatement of the rasterizer FORWARD, differentiable by
autograd.  Used only to check the hand-derived backward formulas of the C oracle (tests/test_oracle_selfcheck.py).
Small scenes only: it materialises [pixels, Gaussians] tensors.  Discrete decisions (culling, radius, tile rects) are
taken from float64 arithmetic of the same formulas."""

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def sh_rgb(deg, sh, d):
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6] + \
            SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8]
    if deg > 2:
        res = res + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10] + \
            SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12] + \
            SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14] + \
            SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15]
    return torch.clamp_min(res + 0.5, 0.0)


def render(means, scales, rots_xyzw, opac, shs, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy, bg, deg=3,
           mod=1.0):
    """All inputs float64 tensors.  Returns color [3,H,W], depth [H,W], alpha [H,W]."""
    P = means.shape[0]
    V, Pm = viewmatrix, projmatrix  # stored transposed: p_view = V^T [x;1]
    ones = torch.ones(P, 1, dtype=means.dtype)
    hom = torch.cat([means, ones], 1)
    pv = hom @ V[:, :3]
    ph = hom @ Pm
    pw = 1.0 / (ph[:, 3] + 1e-7)
    ndc = ph[:, :2] * pw[:, None]
    x, y, z, r = rots_xyzw.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).view(P, 3, 3)
    S = torch.diag_embed(mod * scales)
    Sigma = R @ S @ S @ R.transpose(1, 2)
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    tz = pv[:, 2]
    tx = torch.clamp(pv[:, 0] / tz, -1.3 * tanfovx, 1.3 * tanfovx) * tz
    ty = torch.clamp(pv[:, 1] / tz, -1.3 * tanfovy, 1.3 * tanfovy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * tx / (tz * tz), zero, fy / tz, -fy * ty / (tz * tz)], -1).view(P, 2, 3)
    Rv = V[:3, :3].t()  # world -> view rotation
    A = J @ Rv
    cov = A @ Sigma @ A.transpose(1, 2)
    c00, c01, c11 = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = c00 * c11 - c01 * c01
    conA, conB, conC = c11 / det, -c01 / det, c00 / det
    mid = 0.5 * (c00 + c11)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3 * torch.sqrt(lam)).detach()
    pix = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], -1)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pd = pix.detach()
    x0 = torch.clamp(torch.trunc((pd[:, 0] - radius) / 16), 0, gx)
    y0 = torch.clamp(torch.trunc((pd[:, 1] - radius) / 16), 0, gy)
    x1 = torch.clamp(torch.trunc((pd[:, 0] + radius + 15) / 16), 0, gx)
    y1 = torch.clamp(torch.trunc((pd[:, 1] + radius + 15) / 16), 0, gy)
    visible = (tz.detach() > 0.2) & ((x1 - x0) * (y1 - y0) > 0)
    dirs = means - campos[None]
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    rgb = sh_rgb(deg, shs, dirs)  # shs [P,16,3]
    order = torch.argsort(tz.detach(), stable=True)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=means.dtype), torch.arange(W, dtype=means.dtype), indexing='ij')
    px, py = xs.reshape(-1, 1), ys.reshape(-1, 1)
    tpx, tpy = torch.floor(px / 16), torch.floor(py / 16)
    o = order
    in_tile = visible[o][None] & (tpx >= x0[o][None]) & (tpx < x1[o][None]) & (tpy >= y0[o][None]) & (tpy < y1[o][None])
    dx, dy = pix[o, 0][None] - px, pix[o, 1][None] - py
    power = -0.5 * (conA[o][None] * dx * dx + conC[o][None] * dy * dy) - conB[o][None] * dx * dy
    alpha = torch.clamp_max(opac[o][None] * torch.exp(power), 0.99)
    contrib = in_tile & (power <= 0) & (alpha >= 1.0 / 255.0)
    a = torch.where(contrib, alpha, torch.zeros_like(alpha))
    T_after = torch.cumprod(1 - a, dim=1)
    T_before = torch.cat([torch.ones_like(T_after[:, :1]), T_after[:, :-1]], 1)
    stop = contrib & (T_after.detach() < 1e-4)
    stopped = torch.cumsum(stop.to(torch.int32), dim=1) > 0  # the stopping Gaussian itself is not blended
    wgt = torch.where(stopped, torch.zeros_like(a), a * T_before)
    T_final = 1 - wgt.sum(1)  # = product of (1 - alpha) over the blended ones
    # exact final T (product form) for the background term
    one_m = torch.where(stopped, torch.ones_like(a), 1 - a)
    T_final = torch.prod(one_m, dim=1)
    color = wgt @ rgb[o] + T_final[:, None] * bg[None]
    depth = wgt @ tz[o]
    return color.t().reshape(3, H, W), depth.reshape(H, W), (1 - T_final).reshape(H, W), radius * visible
