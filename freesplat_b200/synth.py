"""Seeded synthetic inputs shaped like the reference's workloads.

Everything is generated on the CPU with a torch.Generator so that the same seed gives the same
tensors here, on the GPU box, in tests and in bench.py.  Conventions follow the reference:
extrinsics are camera-to-world (OpenCV: x right, y down, z forward), intrinsics are normalised
by image size (src/dataset/dataset_scannet.py), near/far = 0.5/15
(config/experiment/scannet/2views.yaml:16-17).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

FX, FY, CX, CY = 0.90, 1.20, 0.5, 0.5      # ScanNet 577 px focal at 640x480
NEAR, FAR = 0.5, 15.0
SH_MASK = (1.0, 0.025, 0.00625, 0.0015625)  # gaussian_adapter.py:127-133: 0.1 * 0.25**degree


def intrinsics(n: int) -> torch.Tensor:
    K = torch.tensor([[FX, 0.0, CX], [0.0, FY, CY], [0.0, 0.0, 1.0]], dtype=torch.float32)
    return K[None].repeat(n, 1, 1)


def _rot_y(a: float) -> torch.Tensor:
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=torch.float32)


def _rot_x(a: float) -> torch.Tensor:
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=torch.float32)


def camera_path(n: int, spacing: float = 0.25, yaw_deg: float = 12.0, t0: float = 0.0) -> torch.Tensor:
    """n camera-to-world poses on a gentle arc looking down +z (0.15-0.4 m apart, <=15 deg yaw)."""
    out = []
    for i in range(n):
        u = (i + t0) - 0.5 * (n - 1)
        yaw = math.radians(yaw_deg) * u / max(n - 1, 1) * 2.0
        pitch = math.radians(2.0) * math.sin(1.3 * u)
        E = torch.eye(4, dtype=torch.float32)
        E[:3, :3] = _rot_y(-yaw) @ _rot_x(pitch)
        E[:3, 3] = torch.tensor([u * spacing, 0.02 * math.cos(u), 0.05 * math.sin(0.7 * u)])
        out.append(E)
    return torch.stack(out)


def smooth_depth(h: int, w: int, g: torch.Generator, lo: float = 1.0, hi: float = 4.0) -> torch.Tensor:
    """Piecewise-smooth depth map in [lo,hi]: low-resolution uniform noise, bicubic upsample."""
    coarse = lo + (hi - lo) * torch.rand((1, 1, 5, 7), generator=g)
    d = F.interpolate(coarse, size=(h, w), mode="bicubic", align_corners=True)[0, 0]
    return d.clamp(lo, hi).contiguous()


def backproject(depth: torch.Tensor, K_norm: torch.Tensor, c2w: torch.Tensor) -> torch.Tensor:
    """Pixel-aligned world points, integer pixel coordinates without +0.5
    (Create_from_depth_map.project, gaussian_adapter.py:36-79)."""
    h, w = depth.shape
    fx, fy, cx, cy = K_norm[0, 0] * w, K_norm[1, 1] * h, K_norm[0, 2] * w, K_norm[1, 2] * h
    ii, jj = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    x = (jj - cx) / fx * depth
    y = (ii - cy) / fy * depth
    cam = torch.stack([x, y, depth, torch.ones_like(depth)], dim=0).reshape(4, -1)
    return (c2w @ cam)[:3].T.contiguous()


def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    """(r,x,y,z) unit quaternions [N,4] -> [N,3,3]."""
    r, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)


@dataclass
class Scene:
    """One scene: a Gaussian set + target cameras, in the reference's Gaussians layout."""
    means: torch.Tensor          # [G,3]
    covariances: torch.Tensor    # [G,3,3]
    harmonics: torch.Tensor      # [G,3,d_sh]
    opacities: torch.Tensor      # [G]
    scales: torch.Tensor         # [G,3]
    rotations: torch.Tensor      # [G,4]
    extrinsics: torch.Tensor     # [V,4,4] target c2w
    intrinsics: torch.Tensor     # [V,3,3]
    near: torch.Tensor           # [V]
    far: torch.Tensor            # [V]
    image_shape: tuple
    context_extrinsics: torch.Tensor

    def to(self, device):
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.to(device) if isinstance(v, torch.Tensor) else v
        return Scene(**kw)


def pixel_aligned_scene(seed: int = 0, h: int = 480, w: int = 640, n_context: int = 2, n_target: int = 3,
                        keep: int | None = 307200, sh_degree: int = 2) -> Scene:
    """Config 2-5 style: pixel-aligned Gaussians of `n_context` views, thinned to `keep`."""
    g = torch.Generator().manual_seed(seed)
    K = intrinsics(1)[0]
    ctx = camera_path(n_context)
    fx_px, fy_px = FX * w, FY * h
    d_sh = (sh_degree + 1) ** 2
    means, scales, depths_all = [], [], []
    for v in range(n_context):
        d = smooth_depth(h, w, g)
        means.append(backproject(d, K, ctx[v]))
        dflat = d.reshape(-1)
        s = (0.5 + 14.5 * torch.sigmoid(torch.randn((h * w, 3), generator=g) - 3.0))
        s = s * dflat[:, None] * 0.1 * (1.0 / fx_px + 1.0 / fy_px)
        scales.append(s)
        depths_all.append(dflat)
    means = torch.cat(means); scales = torch.cat(scales)
    G = means.shape[0]
    if keep is not None and keep < G:
        idx = torch.randperm(G, generator=g)[:keep].sort().values
        means, scales = means[idx], scales[idx]
        G = keep
    q = torch.randn((G, 4), generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    R = quat_to_rot(q)
    cov = R @ torch.diag_embed(scales ** 2) @ R.transpose(1, 2)
    opac = torch.sigmoid(2.0 * torch.randn((G,), generator=g))
    mask = torch.cat([torch.full((2 * l + 1,), SH_MASK[l]) for l in range(sh_degree + 1)])
    sh = torch.randn((G, 3, d_sh), generator=g) * mask
    # targets interpolated between / slightly beyond the contexts
    tgt = camera_path(n_target, spacing=0.25 * max(n_context - 1, 1) / max(n_target, 1), t0=0.37)
    V = n_target
    return Scene(means.contiguous(), cov.contiguous(), sh.contiguous(), opac, scales, q, tgt, intrinsics(V),
                 torch.full((V,), NEAR), torch.full((V,), FAR), (h, w), ctx)


def random_scene(seed: int = 0, h: int = 256, w: int = 256, P: int = 10000, n_target: int = 1,
                 sh_degree: int = 2, sigma_px=(0.5, 8.0)) -> Scene:
    """Config 1 style: P iid Gaussians in the frustum of the first target, sigma log-uniform in pixels."""
    g = torch.Generator().manual_seed(seed)
    tgt = camera_path(n_target, spacing=0.2)
    K = intrinsics(1)[0]
    fx_px = FX * w
    z = 0.8 + 4.0 * torch.rand((P,), generator=g)
    u = (torch.rand((P,), generator=g) * 1.2 - 0.1) * w
    v = (torch.rand((P,), generator=g) * 1.2 - 0.1) * h
    x = (u - CX * w) / (FX * w) * z
    y = (v - CY * h) / (FY * h) * z
    cam = torch.stack([x, y, z, torch.ones_like(z)], dim=0)
    means = (tgt[0] @ cam)[:3].T.contiguous()
    lo, hi = math.log(sigma_px[0]), math.log(sigma_px[1])
    sig = torch.exp(lo + (hi - lo) * torch.rand((P, 1), generator=g)) * (0.5 + torch.rand((P, 3), generator=g))
    scales = sig * z[:, None] / fx_px
    q = torch.randn((P, 4), generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    R = quat_to_rot(q)
    cov = R @ torch.diag_embed(scales ** 2) @ R.transpose(1, 2)
    opac = torch.sigmoid(2.0 * torch.randn((P,), generator=g))
    d_sh = (sh_degree + 1) ** 2
    mask = torch.cat([torch.full((2 * l + 1,), (1.0, 0.3, 0.15, 0.1)[l]) for l in range(sh_degree + 1)])
    sh = torch.randn((P, 3, d_sh), generator=g) * mask
    V = n_target
    return Scene(means, cov.contiguous(), sh.contiguous(), opac, scales, q, tgt, intrinsics(V),
                 torch.full((V,), NEAR), torch.full((V,), FAR), (h, w), tgt)


# ------------------------------------------------------------------------------------------------
# cost-volume inputs, built the way EncoderFreeSplat.forward does (encoder_freesplat.py:218-277)
def cost_volume_inputs(seed: int = 0, n_views: int = 3, K: int | None = None, C: int = 48, Hf: int = 120, Wf: int = 160,
                       spacing: float = 0.25):
    """Returns dict(cur_feats [V,C,Hf,Wf], src_feats [V,K,C,Hf,Wf], src_extrinsics [V,K,4,4]
    (src_cam_T_cur_cam), src_poses [V,K,4,4], src_Ks [V,K,4,4], cur_invK [V,4,4], min_depth, max_depth)."""
    g = torch.Generator().manual_seed(seed)
    V = n_views
    K = V - 1 if K is None else K
    ext = camera_path(V, spacing=spacing)                       # c2w
    Kn = intrinsics(V).clone()
    Kn[:, 0] *= Wf                                              # feature-resolution pixel intrinsics
    Kn[:, 1] *= Hf
    feats = torch.randn((V, C, Hf, Wf), generator=g)
    # low-pass a little so that neighbouring pixels correlate (post-BN-like magnitude ~1)
    feats = 0.6 * feats + 0.4 * F.avg_pool2d(feats, 3, stride=1, padding=1)
    src_idx = []
    for b in range(V):
        others = [j for j in range(V) if j != b]
        others.sort(key=lambda j: abs(j - b))
        src_idx.append(others[:K])
    src_idx = torch.tensor(src_idx)                             # [V,K]
    src_ext = ext[src_idx]                                      # [V,K,4,4]
    src_cam_T_cur_cam = src_ext.inverse() @ ext[:, None]
    cur_cam_T_src_cam = ext.inverse()[:, None] @ src_ext
    src_K = torch.eye(4)[None, None].repeat(V, K, 1, 1)
    src_K[:, :, :3, :3] = Kn[src_idx]
    cur_invK = torch.eye(4)[None].repeat(V, 1, 1)
    cur_invK[:, :3, :3] = Kn.inverse()
    return dict(cur_feats=feats.contiguous(), src_feats=feats[src_idx].contiguous(),
                src_extrinsics=src_cam_T_cur_cam.contiguous(), src_poses=cur_cam_T_src_cam.contiguous(),
                src_Ks=src_K, cur_invK=cur_invK,
                min_depth=torch.tensor(NEAR).view(1, 1, 1, 1), max_depth=torch.tensor(FAR).view(1, 1, 1, 1))


def cost_volume_mlp(seed: int = 0, C: int = 48):
    """Seeded MLP weights in nn.Linear layout: (W0 [32,C+1], b0, W1 [32,32], b1, W2 [1,32], b2)."""
    g = torch.Generator().manual_seed(1000 + seed)
    def lin(o, i):
        bound = 1.0 / math.sqrt(i)
        return (torch.rand((o, i), generator=g) * 2 - 1) * bound, (torch.rand((o,), generator=g) * 2 - 1) * bound
    W0, b0 = lin(32, C + 1); W1, b1 = lin(32, 32); W2, b2 = lin(1, 32)
    return [W0, b0, W1, b1, W2, b2]


# ------------------------------------------------------------------------------------------------
# Pixel-wise Triplet Fusion inputs, shaped as EncoderFreeSplat.forward passes them to fuse_gaussians
# (encoder_freesplat.py:299-368): a smooth analytic surface seen by V cameras, so that neighbouring
# views really do observe the same points (depth-consistent matches) -- plus noise so some do not.
def _surface_depth(c2w: torch.Tensor, K_norm: torch.Tensor, h: int, w: int, iters: int = 12) -> torch.Tensor:
    fx, fy, cx, cy = K_norm[0, 0] * w, K_norm[1, 1] * h, K_norm[0, 2] * w, K_norm[1, 2] * h
    ii, jj = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    dcam = torch.stack([(jj - cx) / fx, (ii - cy) / fy, torch.ones_like(ii)], -1).reshape(-1, 3)   # z = 1 rays
    dw = dcam @ c2w[:3, :3].T
    o = c2w[:3, 3]
    g = lambda x, y: 2.6 + 0.35 * torch.sin(1.7 * x + 0.3) * torch.cos(1.3 * y) + 0.15 * x
    t = torch.full((h * w,), 2.6)
    for _ in range(iters):
        pt = o + t[:, None] * dw
        t = (g(pt[:, 0], pt[:, 1]) - o[2]) / dw[:, 2]
    return t.reshape(h, w)            # camera-space depth (ray has z = 1 in camera coordinates)


def ptf_inputs(seed: int = 0, n_views: int = 3, h: int = 24, w: int = 32, feat_dim: int = 64, noise: float = 0.04,
               spacing: float = 0.2):
    g = torch.Generator().manual_seed(seed)
    V = n_views
    ext = camera_path(V, spacing=spacing)
    Kn = intrinsics(V)
    depths, coords = [], []
    for v in range(V):
        d = _surface_depth(ext[v], Kn[v], h, w)
        d = d + noise * torch.randn((h, w), generator=g) * (torch.rand((h, w), generator=g) < 0.5)
        depths.append(d)
        coords.append(backproject(d, Kn[v], ext[v]))
    depths = torch.stack(depths)                                    # [V,h,w]
    coords = torch.stack(coords)                                    # [V,hw,3]
    feats = torch.randn((1, V, h * w, feat_dim), generator=g) * 0.5
    dens = torch.sigmoid(torch.randn((1, V, h * w, 1, 1), generator=g))
    wemb = torch.rand((1, V, h * w, 1, 1), generator=g)
    return dict(gaussians=[feats], coords=[coords[None, :, :, None, None, :].contiguous()], densities=dens,
                weight_emb=wemb, depths=depths[:, None].contiguous(), extrinsics=ext[None].contiguous(),
                intrinsics=Kn[None].contiguous(), image_shape=(h, w))


def gru_state(seed: int = 0):
    """Seeded weights of networks.GRU (176->64->64 x2, 152->64->64), as a state_dict."""
    g = torch.Generator().manual_seed(2000 + seed)
    sd = {}
    for name, din in (("mlp_z", 176), ("mlp_r", 176), ("mlp_n", 152)):
        for idx, (o, i) in ((0, (64, din)), (2, (64, 64))):
            b = 1.0 / math.sqrt(i)
            sd[f"{name}.{idx}.weight"] = (torch.rand((o, i), generator=g) * 2 - 1) * b
            sd[f"{name}.{idx}.bias"] = (torch.rand((o,), generator=g) * 2 - 1) * b
    return sd
