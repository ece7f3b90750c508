"""Independent dense PyTorch restatement of the rasterizer (test infrastructure).

Written against SURVEY.md Appendix A with plain torch ops (fp64 by default) and NO tiling data
structures: every pixel visits every Gaussian in depth order, masked by the Gaussian's tile
rectangle.  Gradients come from autograd, so agreement with oracle/raster_oracle.c checks both the
oracle's forward and its hand-derived backward.
"""
from __future__ import annotations

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def _sh_rgb(deg, dirs, sh):
    x, y, z = dirs.unbind(-1)
    x, y, z = x[:, None], y[:, None], z[:, None]
    res = C0 * sh[:, 0]
    if deg > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
                   + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
                       + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
                       + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
                       + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return torch.clamp_min(res + 0.5, 0.0)


def cov6_from_scale_rot(scales, rots, mod=1.0):
    r, x, y, z = rots.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = R * (mod * scales)[:, None, :]
    S = M @ M.transpose(1, 2)
    return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1)


def render(*, H, W, tanfovx, tanfovy, bg, viewmatrix, projmatrix, campos, means3D, opacities, shs=None,
           colors_precomp=None, cov6=None, scales=None, rotations=None, sh_degree=0, scale_modifier=1.0,
           dtype=torch.float64):
    """Returns (color[3,H,W], depth[H,W], final_T[H,W], radii[P]).  viewmatrix/projmatrix: flat 16
    (transposed convention of the reference: m[c*4+r] = M[r][c])."""
    cvt = lambda t: None if t is None else (t.to(dtype) if isinstance(t, torch.Tensor) else torch.tensor(t, dtype=dtype))
    means3D, opacities, shs, colors_precomp, cov6, scales, rotations = map(
        cvt, (means3D, opacities, shs, colors_precomp, cov6, scales, rotations))
    vm = cvt(viewmatrix).reshape(4, 4).T   # math matrix M[r][c]
    pm = cvt(projmatrix).reshape(4, 4).T
    campos = cvt(campos).reshape(3); bg = cvt(bg).reshape(3)
    opacities = opacities.reshape(-1)
    P = means3D.shape[0]
    if cov6 is None:
        cov6 = cov6_from_scale_rot(scales, rotations, scale_modifier)
    ones = torch.ones((P, 1), dtype=dtype)
    mh = torch.cat([means3D, ones], -1)
    pv = mh @ vm.T
    ph = mh @ pm.T
    t = pv[:, :3]
    pw = 1.0 / (ph[:, 3] + 0.0000001)
    pproj = ph[:, :2] * pw[:, None]
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tz = t[:, 2]
    txtz, tytz = t[:, 0] / tz, t[:, 1] / tz
    cl_x = (txtz < -limx) | (txtz > limx)
    cl_y = (tytz < -limy) | (tytz > limy)
    # upstream treats the clamped coordinate as a constant in the backward pass
    tx = torch.where(cl_x, (txtz.clamp(-limx, limx) * tz).detach(), t[:, 0])
    ty = torch.where(cl_y, (tytz.clamp(-limy, limy) * tz).detach(), t[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], -1).reshape(P, 2, 3)
    Rw = vm[:3, :3]
    Sig = torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2], cov6[:, 1], cov6[:, 3], cov6[:, 4],
                       cov6[:, 2], cov6[:, 4], cov6[:, 5]], -1).reshape(P, 3, 3)
    Tm = J @ Rw
    c2 = Tm @ Sig @ Tm.transpose(1, 2)
    a = c2[:, 0, 0] + 0.3; b = c2[:, 0, 1]; c = c2[:, 1, 1] + 0.3
    det = a * c - b * b
    conx, cony, conz = c / det, -b / det, a / det
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    px = ((pproj[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((pproj[:, 1] + 1.0) * H - 1.0) * 0.5
    gx, gy = (W + 15) // 16, (H + 15) // 16
    trunc = lambda v: torch.trunc(v).long()
    x0 = trunc((px - radius) / 16).clamp(0, gx); y0 = trunc((py - radius) / 16).clamp(0, gy)
    x1 = trunc((px + radius + 15) / 16).clamp(0, gx); y1 = trunc((py + radius + 15) / 16).clamp(0, gy)
    visible = (tz > 0.2) & (det != 0) & ((x1 - x0) * (y1 - y0) > 0)
    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - campos
        d = d / d.norm(dim=-1, keepdim=True)
        rgb = _sh_rgb(sh_degree, d, shs)
    order = torch.argsort(tz.detach().float(), stable=True)   # keys are the fp32 depth bits
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    tyi, txi = (ys / 16).long(), (xs / 16).long()
    T = torch.ones((H, W), dtype=dtype); C = torch.zeros((3, H, W), dtype=dtype); D = torch.zeros((H, W), dtype=dtype)
    done = torch.zeros((H, W), dtype=torch.bool)
    for i in order.tolist():
        if not bool(visible[i]):
            continue
        inrect = (txi >= x0[i]) & (txi < x1[i]) & (tyi >= y0[i]) & (tyi < y1[i])
        if not bool(inrect.any()):
            continue
        dx, dy = px[i] - xs, py[i] - ys
        power = -0.5 * (conx[i] * dx * dx + conz[i] * dy * dy) - cony[i] * dx * dy
        alpha_raw = opacities[i] * torch.exp(power)
        alpha = alpha_raw + (torch.clamp_max(alpha_raw, 0.99) - alpha_raw).detach()
        valid = inrect & (power <= 0) & (alpha >= 1.0 / 255.0) & ~done
        test_T = T * (1 - alpha)
        newly = valid & (test_T < 0.0001)
        done = done | newly
        valid = valid & ~newly
        w = torch.where(valid, alpha * T, torch.zeros_like(T))
        C = C + rgb[i][:, None, None] * w[None]
        D = D + tz[i] * w
        T = torch.where(valid, test_T, T)
    color = C + T[None] * bg[:, None, None]
    radii = torch.where(visible, radius, torch.zeros_like(radius)).long()
    return color, D, T, radii
