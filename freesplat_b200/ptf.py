"""Pixel-wise Triplet Fusion on the B200 kernels.

`fuse_gaussians` keeps the signature and return value of the reference method
EncoderFreeSplat.fuse_gaussians (/root/reference/src/model/encoder/encoder_freesplat.py:431-522):

    fuse_gaussians(self, gaussians, coords, densities, weight_emb, depths, extrinsics, intrinsics,
                   image_shape, depth_thres=0.1) -> (feats [1,N,F], coords [1,N,3], extrinsics [1,N,4,4], depths [1,N])

so it can be bound onto the reference encoder unchanged:
    EncoderFreeSplat.fuse_gaussians = freesplat_b200.ptf.fuse_gaussians
(`self` only needs `.gru`, the reference's GRU module, networks.py:188-214 -- its Linear layers stay
ordinary nn.Linear / cuBLAS GEMMs as SURVEY §7 prescribes; everything index-related and the state
compaction / weighted merge run in libfreesplat_b200.so through fs_ptf_match / fs_ptf_merge).

The fold over views is inherently sequential (order-dependent GRU and running sums); one host read
of the step counters per view sizes the GRU batch.  CPU tensors raise: there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr


class FsPtfArgs(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("F", C.c_int32), ("n_upper", C.c_int32), ("depth_thres", C.c_float),
        ("feats", C.c_void_p), ("coords", C.c_void_p), ("dens", C.c_void_p), ("wemb", C.c_void_p), ("ext", C.c_void_p),
        ("depth", C.c_void_p), ("counts_in", C.c_void_p),
        ("v_feats", C.c_void_p), ("v_coords", C.c_void_p), ("v_dens", C.c_void_p), ("v_wemb", C.c_void_p),
        ("v_depth", C.c_void_p), ("v_ext", C.c_void_p), ("E_inv", C.c_void_p), ("K_px", C.c_void_p),
        ("zbuf", C.c_void_p), ("pix", C.c_void_p), ("zeta", C.c_void_p), ("match", C.c_void_p), ("append", C.c_void_p),
        ("block_counts", C.c_void_p), ("pair_j", C.c_void_p), ("pair_p", C.c_void_p), ("counts_out", C.c_void_p),
        ("gru_out", C.c_void_p),
        ("o_feats", C.c_void_p), ("o_coords", C.c_void_p), ("o_dens", C.c_void_p), ("o_wemb", C.c_void_p),
        ("o_ext", C.c_void_p), ("o_depth", C.c_void_p),
    ]


def positional_encoding(positions: torch.Tensor, freqs: int) -> torch.Tensor:
    """encoder_freesplat.py:62-77 (ori=False)."""
    freq_bands = (2 ** torch.arange(freqs).float()).to(positions.device)
    pts = (positions[..., None] * freq_bands).reshape(positions.shape[:-1] + (freqs * positions.shape[-1],))
    return torch.stack([torch.sin(pts), torch.cos(pts)], dim=-1).reshape(pts.shape[:-1] + (pts.shape[-1] * 2,))


class _State:
    """Ping-pong SoA buffers of the global Gaussian state (capacity = V*HW)."""

    def __init__(self, cap, F, dev):
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        self.feats, self.coords, self.dens, self.wemb, self.ext, self.depth = e(cap, F), e(cap, 3), e(cap), e(cap), e(cap, 16), e(cap)


def fuse_views(gru, feats, coords, dens, wemb, depths, extrinsics, intrinsics, image_shape, depth_thres=0.1,
               E_inv=None, return_debug=False):
    """Flat form: feats [V,HW,F], coords [V,HW,3], dens/wemb [V,HW], depths [V,HW], extrinsics [V,4,4] (c2w),
    intrinsics [V,3,3] (normalised).  Returns (feats [N,F], coords [N,3], ext [N,4,4], depth [N]) (+ debug)."""
    L = _lib.lib()
    if not feats.is_cuda:
        raise _lib.FreeSplatB200Error("fuse_gaussians needs CUDA tensors (no CPU fallback exists)")
    dev = feats.device
    f32 = lambda t: t.float().contiguous()
    feats, coords, dens, wemb, depths, extrinsics, intrinsics = map(f32, (feats, coords, dens, wemb, depths, extrinsics, intrinsics))
    V, HW, F = feats.shape
    h, w = image_shape
    assert HW == h * w
    cap = V * HW
    stream = torch.cuda.current_stream(dev).cuda_stream
    K_px = intrinsics.clone()
    K_px[:, :1, :] *= w                       # encoder_freesplat.py:445-447
    K_px[:, 1:2, :] *= h
    if E_inv is None:
        E_inv = torch.linalg.inv(extrinsics)  # extrinsic.inverse() (:454)
    E_inv = f32(E_inv)
    ext16 = extrinsics.reshape(V, 16)
    cur, nxt = _State(cap, F, dev), _State(cap, F, dev)
    cur.feats[:HW] = feats[0]; cur.coords[:HW] = coords[0]; cur.dens[:HW] = dens[0]; cur.wemb[:HW] = wemb[0]
    cur.ext[:HW] = ext16[0]; cur.depth[:HW] = depths[0]
    i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
    counts = torch.zeros((V + 1, 8), dtype=torch.int32, device=dev)
    counts[0, 0] = HW
    counts[0, 4] = HW
    zbuf, pix, zeta = i32(HW), i32(cap), torch.empty(cap, dtype=torch.float32, device=dev)
    match, append = torch.empty(cap, dtype=torch.uint8, device=dev), torch.empty(HW, dtype=torch.uint8, device=dev)
    nb = (cap + 1023) // 1024 + 1
    block_counts, pair_j, pair_p = i32(3 * nb), i32(cap), i32(cap)
    n_upper = HW
    N = HW
    debug = []
    with torch.cuda.device(dev):
        for i in range(1, V):
            cin = counts[i - 1, 4:5]                      # N of the current state, on the device
            a = FsPtfArgs(
                H=h, W=w, F=F, n_upper=n_upper, depth_thres=depth_thres,
                feats=ptr(cur.feats), coords=ptr(cur.coords), dens=ptr(cur.dens), wemb=ptr(cur.wemb), ext=ptr(cur.ext),
                depth=ptr(cur.depth), counts_in=cin.data_ptr(),
                v_feats=ptr(feats[i]), v_coords=ptr(coords[i]), v_dens=ptr(dens[i]), v_wemb=ptr(wemb[i]),
                v_depth=ptr(depths[i]), v_ext=ptr(ext16[i]), E_inv=ptr(E_inv[i]), K_px=ptr(K_px[i]),
                zbuf=ptr(zbuf), pix=ptr(pix), zeta=ptr(zeta), match=ptr(match), append=ptr(append),
                block_counts=ptr(block_counts), pair_j=ptr(pair_j), pair_p=ptr(pair_p), counts_out=counts[i].data_ptr(),
                gru_out=None, o_feats=ptr(nxt.feats), o_coords=ptr(nxt.coords), o_dens=ptr(nxt.dens), o_wemb=ptr(nxt.wemb),
                o_ext=ptr(nxt.ext), o_depth=ptr(nxt.depth))
            check(L.fs_ptf_match(C.byref(a), C.c_void_p(stream)), "fs_ptf_match")
            c = counts[i].tolist()                         # the one host read of this step
            M, N_out = c[2], c[4]
            gru_out = None
            if M > 0:
                pj, pp = pair_j[:M].long(), pair_p[:M].long()
                hidden = cur.feats[pj]                     # global latent   (networks.py:201 `hidden_feat`)
                inp = feats[i][pp]                         # view-i latent   (`input_feat`)
                e_in = positional_encoding(torch.stack([cur.dens[pj], wemb[i][pp]], -1), 6)
                e_h = positional_encoding(torch.stack([dens[i][pp], cur.wemb[pj]], -1), 6)
                gru_out = gru(inp[None, :, None, :], hidden[None, :, None, :], e_in[None, :, None, :],
                              e_h[None, :, None, :])[0, :, 0, :].float().contiguous()
            a.gru_out = ptr(gru_out)
            check(L.fs_ptf_merge(C.byref(a), C.c_void_p(stream)), "fs_ptf_merge")
            if return_debug:
                debug.append(dict(pix=pix[:N].clone(), zeta=zeta[:N].clone(), match=match[:N].clone(),
                                  append=append.clone(), zbuf=zbuf.clone(), counts=c))
            cur, nxt = nxt, cur
            N = N_out
            n_upper = N
    out = (cur.feats[:N], cur.coords[:N], cur.ext[:N].reshape(N, 4, 4), cur.depth[:N])
    if return_debug:
        return out, debug, (cur.dens[:N], cur.wemb[:N])
    return out


def fuse_gaussians(self, gaussians, coords, densities, weight_emb, depths, extrinsics, intrinsics, image_shape,
                   depth_thres=0.1):
    """Reference signature (shapes as passed at encoder_freesplat.py:357-368)."""
    g = gaussians[0][0]                         # [V,HW,F]
    x = coords[0][0, :, :, 0, 0, :]             # [V,HW,3]
    V = g.shape[0]
    d = densities[0].reshape(V, -1)
    we = weight_emb[0].reshape(V, -1)
    dep = depths.reshape(V, -1)                 # "v c h w -> v (c h w)"
    F_, X_, E_, Z_ = fuse_views(self.gru, g, x, d, we, dep, extrinsics[0], intrinsics[0], image_shape, depth_thres)
    return F_[None], X_[None], E_[None], Z_[None]
