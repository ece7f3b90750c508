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
import os

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
        ("o_ext", C.c_void_p), ("o_depth", C.c_void_p), ("map_old", C.c_void_p), ("map_px", C.c_void_p),
    ]


class FsPtfMergeBwdArgs(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("F", C.c_int32), ("N", C.c_int32), ("n_keep", C.c_int32), ("n_match", C.c_int32)] + \
        [(n, C.c_void_p) for n in ("coords", "dens", "ext", "depth", "v_coords", "v_dens", "v_depth", "v_ext", "match", "pix", "map_old",
                                   "map_px", "g_feats", "g_coords", "g_dens", "g_wemb", "g_ext", "g_depth", "d_feats", "d_coords",
                                   "d_dens", "d_wemb", "d_ext", "d_depth", "dv_feats", "dv_coords", "dv_dens", "dv_wemb", "dv_depth",
                                   "d_gru")]


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


def _gru_has_reference_structure(gru) -> bool:
    """networks.py:188-200: mlp_z / mlp_r / mlp_n = Sequential(Linear, ReLU, Linear)."""
    try:
        return all(isinstance(getattr(gru, n)[0], torch.nn.Linear) and isinstance(getattr(gru, n)[2], torch.nn.Linear)
                   and len(getattr(gru, n)) == 3 for n in ("mlp_z", "mlp_r", "mlp_n"))
    except Exception:
        return False


class FsPtfGruArgs(C.Structure):
    _fields_ = [("M", C.c_int32), ("flags", C.c_int32), ("pair_j", C.c_void_p), ("pair_p", C.c_void_p),
                ("feats", C.c_void_p), ("dens", C.c_void_p), ("wemb", C.c_void_p),
                ("v_feats", C.c_void_p), ("v_dens", C.c_void_p), ("v_wemb", C.c_void_p),
                ("W_r0", C.c_void_p), ("W_z0", C.c_void_p), ("W_r2", C.c_void_p), ("W_z2", C.c_void_p), ("W_n0", C.c_void_p),
                ("W_n2", C.c_void_p), ("biases", C.c_void_p), ("wscratch", C.c_void_p), ("out", C.c_void_p), ("M_dev", C.c_void_p),
                ("save", C.c_void_p), ("save_a1", C.c_void_p)]


class FsGruBwdDataArgs(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("mode", C.c_int32), ("lda", C.c_int32), ("ldc", C.c_int32), ("ldm", C.c_int32),
                ("A", C.c_void_p), ("W", C.c_void_p), ("mask", C.c_void_p), ("C", C.c_void_p),
                ("h", C.c_void_p), ("r_lin", C.c_void_p), ("dr_lin", C.c_void_p), ("ldh", C.c_int32), ("reserved", C.c_int32)]


class FsGruBwdWeightsArgs(C.Structure):
    _fields_ = [("M", C.c_int32), ("ldy0", C.c_int32), ("ldy1", C.c_int32), ("ldx0", C.c_int32), ("nx0", C.c_int32), ("ldx1", C.c_int32),
                ("nx1", C.c_int32), ("ldg", C.c_int32), ("Y0", C.c_void_p), ("Y1", C.c_void_p), ("X0", C.c_void_p), ("X1", C.c_void_p),
                ("G", C.c_void_p)]


# "tc": the whole GRU on the tensor cores (fs_ptf_gru, 3xTF32) ; "cublas": glue kernels + nn.Linear GEMMs
GRU_MODE = os.environ.get("FREESPLAT_B200_PTF_GRU", "tc")
# 1: the inference fold never reads a counter back until the end (grids sized by upper bounds); 0: one host read per step
SYNC_FREE = os.environ.get("FREESPLAT_B200_PTF_SYNC_FREE", "1") == "1"


def _gru_tc_ok(gru, F) -> bool:
    try:
        return (F == 64 and tuple(gru.mlp_r[0].weight.shape) == (64, 176) and tuple(gru.mlp_z[0].weight.shape) == (64, 176)
                and tuple(gru.mlp_n[0].weight.shape) == (64, 152) and all(tuple(getattr(gru, n)[2].weight.shape) == (64, 64)
                                                                         for n in ("mlp_r", "mlp_z", "mlp_n")))
    except Exception:
        return False


class _GruTc:
    """GRU.forward (networks.py:201-214) for the matched pairs in ONE kernel on the tensor cores (fs_ptf_gru).  The weights
    are split / tiled into `wscratch` by the first call and reused by the following fold steps; with `M_dev` the pair count
    stays on the device (M is then only the upper bound that sizes the grid)."""

    def __init__(self, gru, dev):
        L = _lib.lib()
        w = lambda t: t.detach().float().contiguous()
        self.Ws = [w(gru.mlp_r[0].weight), w(gru.mlp_z[0].weight), w(gru.mlp_r[2].weight), w(gru.mlp_z[2].weight),
                   w(gru.mlp_n[0].weight), w(gru.mlp_n[2].weight)]
        self.biases = torch.cat([w(gru.mlp_r[0].bias), w(gru.mlp_z[0].bias), w(gru.mlp_r[2].bias), w(gru.mlp_z[2].bias),
                                 w(gru.mlp_n[0].bias), w(gru.mlp_n[2].bias)]).contiguous()
        L.fs_ptf_gru_wscratch_bytes.restype = C.c_int64
        self.scratch = torch.empty(int(L.fs_ptf_gru_wscratch_bytes()), dtype=torch.uint8, device=dev)
        self.prepared = False
        self.dev = dev

    def __call__(self, M, pair_j, pair_p, state, view_feats, view_dens, view_wemb, stream, out=None, M_dev=None, save=None, save_a1=None):
        L = _lib.lib()
        Ws = self.Ws
        if out is None:
            out = torch.empty((M, 64), dtype=torch.float32, device=self.dev)
        a = FsPtfGruArgs(M=M, flags=int(self.prepared), pair_j=ptr(pair_j), pair_p=ptr(pair_p), feats=ptr(state[0]), dens=ptr(state[2]),
                         wemb=ptr(state[3]), v_feats=ptr(view_feats), v_dens=ptr(view_dens), v_wemb=ptr(view_wemb),
                         W_r0=ptr(Ws[0]), W_z0=ptr(Ws[1]), W_r2=ptr(Ws[2]), W_z2=ptr(Ws[3]), W_n0=ptr(Ws[4]), W_n2=ptr(Ws[5]),
                         biases=ptr(self.biases), wscratch=ptr(self.scratch), out=ptr(out), M_dev=ptr(M_dev), save=ptr(save), save_a1=ptr(save_a1))
        check(L.fs_ptf_gru(C.byref(a), C.c_void_p(stream)), "fs_ptf_gru")
        self.prepared = True
        return out


def _gru_fused_tc(gru, M, pair_j, pair_p, state, view_feats, view_dens, view_wemb, stream):
    return _GruTc(gru, state[0].device)(M, pair_j, pair_p, state, view_feats, view_dens, view_wemb, stream)


def _gru_fused(gru, M, F, pair_j, pair_p, state, view_feats, view_dens, view_wemb, stream):
    """GRU.forward (networks.py:201-214) for the matched pairs, inference path: three glue kernels of ours around the
    six nn.Linear GEMMs (cuBLAS), instead of ~30 element-wise torch launches."""
    L = _lib.lib()
    dev = state[0].device
    A1 = torch.empty((M, 2 * F + 48), dtype=torch.float32, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())
    check(L.fs_ptf_gru_inputs(C.c_int32(M), C.c_int32(F), vp(pair_j), vp(pair_p), vp(state[0]), vp(state[2]), vp(state[3]),
                              vp(view_feats), vp(view_dens), vp(view_wemb), vp(A1), C.c_void_p(stream)), "fs_ptf_gru_inputs")
    r_lin = gru.mlp_r(A1).contiguous()
    z_lin = gru.mlp_z(A1).contiguous()
    U = torch.empty((M, 2 * F + 24), dtype=torch.float32, device=dev)
    check(L.fs_ptf_gru_update(C.c_int32(M), C.c_int32(F), vp(A1), vp(r_lin), vp(U), C.c_void_p(stream)), "fs_ptf_gru_update")
    q_lin = gru.mlp_n(U).contiguous()
    out = torch.empty((M, F), dtype=torch.float32, device=dev)
    check(L.fs_ptf_gru_output(C.c_int32(M), C.c_int32(F), vp(A1), vp(z_lin), vp(q_lin), vp(out), C.c_void_p(stream)), "fs_ptf_gru_output")
    return out


# "tc": the 12 matrix products of the GRU backward on the tensor cores (fs_ptf_gru_bwd_data / _weights); "cublas": the
# round-2a path (recompute + 18 torch fp32 GEMMs) kept for A/B runs
GRU_BWD = os.environ.get("FREESPLAT_B200_PTF_GRU_BWD", "tc")


def _bwd_data(L, st, A, W, N, out, mode=0, mask=None, h=None, r_lin=None, dr_lin=None):
    """out[M,N] (op)= A[M,64] @ W[64,N] on the tensor cores (fs_ptf_gru_bwd_data); mode 3: out = dA1, see the header."""
    a = FsGruBwdDataArgs(M=A.shape[0], N=N, mode=mode, lda=A.stride(0), ldc=out.stride(0), ldm=0 if mask is None else mask.stride(0),
                         A=ptr(A), W=ptr(W), mask=ptr(mask), C=ptr(out), h=ptr(h), r_lin=ptr(r_lin), dr_lin=ptr(dr_lin),
                         ldh=0 if h is None else h.stride(0))
    check(L.fs_ptf_gru_bwd_data(C.byref(a), st), "fs_ptf_gru_bwd_data")
    return out


def _bwd_weights(L, st, Y0, Y1, X0, X1, G):
    """G[128,ldg] = [Y0 | Y1]^T @ [X0 | X1 | 1] over the pairs (fs_ptf_gru_bwd_weights)."""
    a = FsGruBwdWeightsArgs(M=Y0.shape[0], ldy0=Y0.stride(0), ldy1=0 if Y1 is None else Y1.stride(0), ldx0=X0.stride(0), nx0=X0.shape[1],
                            ldx1=0 if X1 is None else X1.stride(0), nx1=0 if X1 is None else X1.shape[1], ldg=G.shape[1],
                            Y0=ptr(Y0), Y1=ptr(Y1), X0=ptr(X0), X1=ptr(X1), G=ptr(G))
    check(L.fs_ptf_gru_bwd_weights(C.byref(a), st), "fs_ptf_gru_bwd_weights")
    return G


class _GruTrain(torch.autograd.Function):
    """GRU.forward (networks.py:201-214) of the matched pairs for the TRAINING fold, all of it in libfreesplat_b200.so.

    forward : the fused tensor-core kernel (fs_ptf_gru, 3xTF32), which also leaves the six intermediate activations
              [Hr | Hz | r_lin | z_lin | Hn | q_lin] (1.5 KB per pair) for the backward.
    backward: chain rule by hand.  Element-wise work (gathers, positional encodings and their derivatives, gates, scatter of
              the input gradients incl. atomics for tied pixels) = the glue kernels fs_ptf_gru_{inputs,update} and
              fs_ptf_gru_{output,update,inputs}_backward; the six data-gradient products = fs_ptf_gru_bwd_data (ReLU masks and
              the accumulation into dA1 in its epilogue); the six weight gradients and six bias gradients = three
              fs_ptf_gru_bwd_weights launches (two layers stacked on the 128 MMA rows, a column of ones for the biases).
              No torch matmul / cuBLAS, no recomputation of the layers.
    FREESPLAT_B200_PTF_GRU_BWD=cublas selects the earlier recompute + 18 fp32 GEMM path (A/B timing, cross-check in the tests)."""

    @staticmethod
    def forward(ctx, feats, dens, wemb, v_feats, v_dens, v_wemb, pair_j, pair_p, M, gru_tc, *params):
        dev = feats.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        state = (feats.contiguous(), None, dens.contiguous(), wemb.contiguous())
        ctx.tc = GRU_BWD == "tc"
        acts = torch.empty((6, max(M, 1), 64), dtype=torch.float32, device=dev) if ctx.tc else None
        a1 = torch.empty((max(M, 1), 2 * 64 + 48), dtype=torch.float32, device=dev) if ctx.tc else None
        with torch.cuda.device(dev):
            out = gru_tc(M, pair_j, pair_p, state, v_feats.contiguous(), v_dens.contiguous(), v_wemb.contiguous(), stream, save=acts,
                         save_a1=a1)
        ctx.M = M
        ctx.sizes = (feats.shape[0], v_feats.shape[0])
        ctx.acts, ctx.a1 = acts, a1
        ctx.save_for_backward(state[0], state[2], state[3], v_feats, v_dens, v_wemb, pair_j[:M].clone(), pair_p[:M].clone(), *params)
        return out

    @staticmethod
    def backward(ctx, g):
        if not ctx.tc:
            return _GruTrain._backward_cublas(ctx, g)
        L = _lib.lib()
        feats, dens, wemb, v_feats, v_dens, v_wemb, pj, pp = ctx.saved_tensors[:8]
        (Wr0, br0, Wr2, br2, Wz0, bz0, Wz2, bz2, Wn0, bn0, Wn2, bn2) = ctx.saved_tensors[8:]
        M, F = ctx.M, feats.shape[1]
        N, HW = ctx.sizes
        dev = feats.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        vp = lambda t: C.c_void_p(t.data_ptr())
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        w = lambda t: t.detach().float().contiguous()
        Hr, Hz, r_lin, z_lin, Hn, q_lin = ctx.acts.unbind(0)
        g = g.contiguous()
        K1, K3 = 2 * F + 48, 2 * F + 24
        with torch.cuda.device(dev), torch.no_grad():
            A1, U = ctx.a1[:M], e(M, K3)
            check(L.fs_ptf_gru_update(C.c_int32(M), C.c_int32(F), vp(A1), vp(r_lin), vp(U), st), "fs_ptf_gru_update")
            dHn, dHr, dHz = e(M, F), e(M, F), e(M, F)
            dz_c, dq_c, dr_c, dA1 = e(M, F), e(M, F), e(M, F), e(M, K1)
            check(L.fs_ptf_gru_output_backward(C.c_int32(M), C.c_int32(F), vp(A1), vp(z_lin), vp(q_lin), vp(g), vp(dz_c), vp(dq_c),
                                               vp(dA1), st), "fs_ptf_gru_output_backward")
            _bwd_data(L, st, dq_c, w(Wn2), F, dHn, mode=1, mask=Hn)
            # dU = dHn @ Wn0 never reaches memory: the update-gate chain rule runs in the product's epilogue (mode 3)
            _bwd_data(L, st, dHn, w(Wn0), K3, dA1, mode=3, h=A1, r_lin=r_lin, dr_lin=dr_c)
            _bwd_data(L, st, dz_c, w(Wz2), F, dHz, mode=1, mask=Hz)
            _bwd_data(L, st, dr_c, w(Wr2), F, dHr, mode=1, mask=Hr)
            _bwd_data(L, st, dHz, w(Wz0), K1, dA1, mode=2)
            _bwd_data(L, st, dHr, w(Wr0), K1, dA1, mode=2)
            Gn = _bwd_weights(L, st, dq_c, dHn, Hn, U, e(128, 224))          # [Hn (64) | U (152) | 1]
            G2 = _bwd_weights(L, st, dr_c, dz_c, Hr, Hz, e(128, 144))        # [Hr (64) | Hz (64) | 1]
            G0 = _bwd_weights(L, st, dHr, dHz, A1, None, e(128, 192))        # [A1 (176) | 1]
            gWn2, gbn2, gWn0, gbn0 = Gn[:F, :F], Gn[:F, F + K3], Gn[F:, F:F + K3], Gn[F:, F + K3]
            gWr2, gbr2, gWz2, gbz2 = G2[:F, :F], G2[:F, 2 * F], G2[F:, F:2 * F], G2[F:, 2 * F]
            gWr0, gbr0, gWz0, gbz0 = G0[:F, :K1], G0[:F, K1], G0[F:, :K1], G0[F:, K1]
            d_feats, d_dens, d_wemb = z(N, F), z(N), z(N)
            dv_feats, dv_dens, dv_wemb = z(HW, F), z(HW), z(HW)
            check(L.fs_ptf_gru_inputs_backward(C.c_int32(M), C.c_int32(F), vp(pj), vp(pp), vp(dens), vp(wemb), vp(v_dens), vp(v_wemb), vp(dA1),
                                               vp(d_feats), vp(d_dens), vp(d_wemb), vp(dv_feats), vp(dv_dens), vp(dv_wemb), st),
                  "fs_ptf_gru_inputs_backward")
        return (d_feats, d_dens, d_wemb, dv_feats, dv_dens, dv_wemb, None, None, None, None,
                gWr0, gbr0, gWr2, gbr2, gWz0, gbz0, gWz2, gbz2, gWn0, gbn0, gWn2, gbn2)

    @staticmethod
    def _backward_cublas(ctx, g):
        L = _lib.lib()
        feats, dens, wemb, v_feats, v_dens, v_wemb, pj, pp = ctx.saved_tensors[:8]
        (Wr0, br0, Wr2, br2, Wz0, bz0, Wz2, bz2, Wn0, bn0, Wn2, bn2) = ctx.saved_tensors[8:]
        M, F = ctx.M, feats.shape[1]
        N, HW = ctx.sizes
        dev = feats.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        vp = lambda t: C.c_void_p(t.data_ptr())
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        g = g.contiguous()
        with torch.cuda.device(dev), torch.no_grad():
            # ---- recompute the forward activations ----
            A1 = e(M, 2 * F + 48)
            check(L.fs_ptf_gru_inputs(C.c_int32(M), C.c_int32(F), vp(pj), vp(pp), vp(feats), vp(dens), vp(wemb), vp(v_feats), vp(v_dens),
                                      vp(v_wemb), vp(A1), st), "fs_ptf_gru_inputs")
            Hr = torch.addmm(br0, A1, Wr0.t()).relu_(); Hz = torch.addmm(bz0, A1, Wz0.t()).relu_()
            r_lin = torch.addmm(br2, Hr, Wr2.t()); z_lin = torch.addmm(bz2, Hz, Wz2.t())
            U = e(M, 2 * F + 24)
            check(L.fs_ptf_gru_update(C.c_int32(M), C.c_int32(F), vp(A1), vp(r_lin), vp(U), st), "fs_ptf_gru_update")
            Hn = torch.addmm(bn0, U, Wn0.t()).relu_()
            q_lin = torch.addmm(bn2, Hn, Wn2.t())
            # ---- backward ----
            dz_lin, dq_lin, dA1 = e(M, F), e(M, F), e(M, 2 * F + 48)
            check(L.fs_ptf_gru_output_backward(C.c_int32(M), C.c_int32(F), vp(A1), vp(z_lin), vp(q_lin), vp(g), vp(dz_lin), vp(dq_lin),
                                               vp(dA1), st), "fs_ptf_gru_output_backward")
            gWn2 = dq_lin.t() @ Hn; gbn2 = dq_lin.sum(0)
            dHn = (dq_lin @ Wn2).mul_(Hn > 0)
            gWn0 = dHn.t() @ U; gbn0 = dHn.sum(0)
            dU = dHn @ Wn0
            dr_lin = e(M, F)
            check(L.fs_ptf_gru_update_backward(C.c_int32(M), C.c_int32(F), vp(A1), vp(r_lin), vp(dU), vp(dr_lin), vp(dA1), st),
                  "fs_ptf_gru_update_backward")
            gWz2 = dz_lin.t() @ Hz; gbz2 = dz_lin.sum(0)
            dHz = (dz_lin @ Wz2).mul_(Hz > 0)
            gWz0 = dHz.t() @ A1; gbz0 = dHz.sum(0)
            dA1.addmm_(dHz, Wz0)
            gWr2 = dr_lin.t() @ Hr; gbr2 = dr_lin.sum(0)
            dHr = (dr_lin @ Wr2).mul_(Hr > 0)
            gWr0 = dHr.t() @ A1; gbr0 = dHr.sum(0)
            dA1.addmm_(dHr, Wr0)
            d_feats, d_dens, d_wemb = z(N, F), z(N), z(N)
            dv_feats, dv_dens, dv_wemb = z(HW, F), z(HW), z(HW)
            check(L.fs_ptf_gru_inputs_backward(C.c_int32(M), C.c_int32(F), vp(pj), vp(pp), vp(dens), vp(wemb), vp(v_dens), vp(v_wemb), vp(dA1),
                                               vp(d_feats), vp(d_dens), vp(d_wemb), vp(dv_feats), vp(dv_dens), vp(dv_wemb), st),
                  "fs_ptf_gru_inputs_backward")
        return (d_feats, d_dens, d_wemb, dv_feats, dv_dens, dv_wemb, None, None, None, None,
                gWr0, gbr0, gWr2, gbr2, gWz0, gbz0, gWz2, gbz2, gWn0, gbn0, gWn2, gbn2)


def _gru_train(gru, gru_tc, M, pair_j, pair_p, state, v_feats, v_dens, v_wemb):
    p = (gru.mlp_r[0].weight, gru.mlp_r[0].bias, gru.mlp_r[2].weight, gru.mlp_r[2].bias, gru.mlp_z[0].weight, gru.mlp_z[0].bias,
         gru.mlp_z[2].weight, gru.mlp_z[2].bias, gru.mlp_n[0].weight, gru.mlp_n[0].bias, gru.mlp_n[2].weight, gru.mlp_n[2].bias)
    return _GruTrain.apply(state[0], state[2], state[3], v_feats, v_dens, v_wemb, pair_j, pair_p, M, gru_tc, *p)


def _ptf_args(h, w, F, n_upper, depth_thres, state, cin, view, scratch, counts_out, out=None, gru_out=None, maps=(None, None)):
    feats, coords, dens, wemb, ext, depth = state
    v_feats, v_coords, v_dens, v_wemb, v_depth, v_ext, E_inv, K_px = view
    zbuf, pix, zeta, match, append, block_counts, pair_j, pair_p = scratch
    o = out if out is not None else (None,) * 6
    return FsPtfArgs(
        H=h, W=w, F=F, n_upper=n_upper, depth_thres=depth_thres,
        feats=ptr(feats), coords=ptr(coords), dens=ptr(dens), wemb=ptr(wemb), ext=ptr(ext), depth=ptr(depth),
        counts_in=ptr(cin), v_feats=ptr(v_feats), v_coords=ptr(v_coords), v_dens=ptr(v_dens), v_wemb=ptr(v_wemb),
        v_depth=ptr(v_depth), v_ext=ptr(v_ext), E_inv=ptr(E_inv), K_px=ptr(K_px),
        zbuf=ptr(zbuf), pix=ptr(pix), zeta=ptr(zeta), match=ptr(match), append=ptr(append), block_counts=ptr(block_counts),
        pair_j=ptr(pair_j), pair_p=ptr(pair_p), counts_out=ptr(counts_out), gru_out=ptr(gru_out),
        o_feats=ptr(o[0]), o_coords=ptr(o[1]), o_dens=ptr(o[2]), o_wemb=ptr(o[3]), o_ext=ptr(o[4]), o_depth=ptr(o[5]),
        map_old=ptr(maps[0]), map_px=ptr(maps[1]))


class _PtfMerge(torch.autograd.Function):
    """Differentiable wrapper of fs_ptf_merge (training path).  forward = the compaction / merge kernel, which also leaves
    the index maps (where every old row / appended pixel went); backward = fs_ptf_merge_backward: one kernel routes the
    gradients back (kept -> old state, appended -> view i, fused -> both by their density weights plus the derivative of the
    weighted means w.r.t. the densities; the latent rows of fused Gaussians go to the GRU output)."""

    @staticmethod
    def forward(ctx, feats, coords, dens, wemb, ext, depth, v_feats, v_coords, v_dens, v_wemb, v_depth, gru_out, meta):
        L = _lib.lib()
        (h, w, depth_thres, cin, v_ext, E_inv, K_px, scratch, counts_row, c) = meta
        N, nk, M, na, N_out = c[0], c[1], c[2], c[3], c[4]
        dev = feats.device
        F = feats.shape[1]
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        out = (e(N_out, F), e(N_out, 3), e(N_out), e(N_out), e(N_out, 16), e(N_out))
        state = tuple(t.contiguous() for t in (feats, coords, dens, wemb, ext, depth))
        view = (v_feats.contiguous(), v_coords.contiguous(), v_dens.contiguous(), v_wemb.contiguous(), v_depth.contiguous(),
                v_ext, E_inv, K_px)
        zbuf, pix, zeta, match, append, block_counts, pair_j, pair_p = scratch
        map_old = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
        map_px = torch.empty(h * w, dtype=torch.int32, device=dev)
        a = _ptf_args(h, w, F, N, depth_thres, state, cin, view, scratch, counts_row, out,
                      None if gru_out is None else gru_out.contiguous(), maps=(map_old, map_px))
        with torch.cuda.device(dev):
            check(L.fs_ptf_merge(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_ptf_merge")
        ctx.sizes = (N, nk, M, na, h, w, F)
        # the scratch buffers are reused by the next fold step: keep private copies of the two that backward reads
        ctx.save_for_backward(state[1], state[2], state[4], state[5], view[1], view[2], view[4], v_ext,
                              match[:max(N, 1)].clone(), pix[:max(N, 1)].clone(), map_old, map_px)
        ctx.has_gru = gru_out is not None
        return out

    @staticmethod
    def backward(ctx, gF, gX, gD, gW, gE, gZ):
        L = _lib.lib()
        (coords, dens, ext, depth, v_coords, v_dens, v_depth, v_ext, match, pix, map_old, map_px) = ctx.saved_tensors
        N, nk, M, na, h, w, F = ctx.sizes
        dev = coords.device
        HW = h * w
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        c = lambda g: None if g is None else g.contiguous()
        gF, gX, gD, gW, gE, gZ = map(c, (gF, gX, gD, gW, gE, gZ))
        d_state = (e(N, F), e(N, 3), e(N), e(N), e(N, 16), e(N))
        d_view = (e(HW, F), e(HW, 3), e(HW), e(HW), e(HW))
        d_gru = e(M, F) if M > 0 else None
        a = FsPtfMergeBwdArgs(H=h, W=w, F=F, N=N, n_keep=nk, n_match=M, coords=ptr(coords), dens=ptr(dens), ext=ptr(ext),
                              depth=ptr(depth), v_coords=ptr(v_coords), v_dens=ptr(v_dens), v_depth=ptr(v_depth), v_ext=ptr(v_ext),
                              match=ptr(match), pix=ptr(pix), map_old=ptr(map_old), map_px=ptr(map_px),
                              g_feats=ptr(gF), g_coords=ptr(gX), g_dens=ptr(gD), g_wemb=ptr(gW), g_ext=ptr(gE), g_depth=ptr(gZ),
                              d_feats=ptr(d_state[0]), d_coords=ptr(d_state[1]), d_dens=ptr(d_state[2]), d_wemb=ptr(d_state[3]),
                              d_ext=ptr(d_state[4]), d_depth=ptr(d_state[5]), dv_feats=ptr(d_view[0]), dv_coords=ptr(d_view[1]),
                              dv_dens=ptr(d_view[2]), dv_wemb=ptr(d_view[3]), dv_depth=ptr(d_view[4]), d_gru=ptr(d_gru))
        with torch.cuda.device(dev):
            check(L.fs_ptf_merge_backward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_ptf_merge_backward")
        return (*d_state, *d_view, d_gru if ctx.has_gru else None, None)


# 1: the inference fold keeps the state in an append-only pool (rows never move, an order index carries the reference's output
#    order, one gather at the end); 0: the compacting fold (fs_ptf_merge rewrites the whole state every step)
#    Measured (B200, 640x480): 10 views 4.36-4.51 ms (pool) vs 4.49-4.61 ms (compacting), 3 views 1.15 vs 1.08 ms: the final gather
#    and the order passes cost what the smaller per-step traffic saves on short folds -> "auto" uses the pool from 6 views on
#    (it also halves the state memory: no ping-pong buffers).
POOL = os.environ.get("FREESPLAT_B200_PTF_POOL", "auto")
POOL = {"1": True, "0": False}.get(POOL, "auto")


def _fold_pool(L, gru_tc, feats, coords, dens, wemb, depths, ext16, E_inv, K_px, h, w, F, V, HW, cap, depth_thres, counts, scratch, stream,
               view_ready, ts, dev):
    """Sync-free inference fold on the append-only pool (csrc/ptf.cu, "Append-only pool").

    The fold is a few small kernels per view; what the host spends per step decides whether the GPU waits (measured: 6.2 ms of
    Python for 4.1 ms of kernels when every step re-built its argument structs from tensor views).  The two structs are therefore
    built ONCE and only the fields that change are patched per step with plain pointer arithmetic."""
    zbuf, pix, zeta, match, append, block_counts, pair_j, pair_p = scratch
    e = lambda *s_: torch.empty(s_, dtype=torch.float32, device=dev)
    pool = (e(cap, F), e(cap, 3), e(cap), e(cap), e(cap, 16), e(cap))
    pool[0][:HW] = feats[0]; pool[1][:HW] = coords[0]; pool[2][:HW] = dens[0]; pool[3][:HW] = wemb[0]
    pool[4][:HW] = ext16[0]; pool[5][:HW] = depths[0]
    if V == 1:
        return (pool[0][:HW], pool[1][:HW], pool[4][:HW].reshape(HW, 4, 4), pool[5][:HW])
    match_all = torch.empty((V - 1, cap), dtype=torch.uint8, device=dev)       # step i's flags by pool row: the order passes read them later
    phys = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(2)]
    blk = torch.empty((cap + 511) // 512 + 1, dtype=torch.int32, device=dev)
    gru_buf = e(min(cap, (V - 1) * HW), F)
    side = torch.cuda.Stream(dev)
    ev = torch.cuda.Event()

    def base_stride(t):                       # per-view pointer = base + i * stride (the views may be strided slices of a block)
        return t[0].data_ptr(), (t[1].data_ptr() - t[0].data_ptr())
    pf, pc, pd, pw, pz = (base_stride(t) for t in (feats, coords, dens, wemb, depths))
    p_ext, p_einv, p_kpx, p_counts, p_match = ext16.data_ptr(), E_inv.data_ptr(), K_px.data_ptr(), counts.data_ptr(), match_all.data_ptr()
    sc = (zbuf, pix, zeta, match_all[0], append, block_counts, pair_j, pair_p)
    view0 = (feats[1], coords[1], dens[1], wemb[1], depths[1], ext16[1], E_inv[1], K_px[1])
    a = _ptf_args(h, w, F, HW, depth_thres, pool, counts[0, 4:5], view0, sc, counts[1])
    a.gru_out = ptr(gru_buf)
    Ws = gru_tc.Ws
    g = FsPtfGruArgs(M=HW, flags=int(gru_tc.prepared), pair_j=ptr(pair_j), pair_p=ptr(pair_p), feats=ptr(pool[0]), dens=ptr(pool[2]),
                     wemb=ptr(pool[3]), v_feats=pf[0], v_dens=pd[0], v_wemb=pw[0], W_r0=ptr(Ws[0]), W_z0=ptr(Ws[1]), W_r2=ptr(Ws[2]),
                     W_z2=ptr(Ws[3]), W_n0=ptr(Ws[4]), W_n2=ptr(Ws[5]), biases=ptr(gru_tc.biases), wscratch=ptr(gru_tc.scratch),
                     out=ptr(gru_buf), M_dev=None)
    st, sst = C.c_void_p(stream), C.c_void_p(side.cuda_stream)
    ra, rg = C.byref(a), C.byref(g)
    phys_ptr = (phys[0].data_ptr(), phys[1].data_ptr())
    blk_ptr = C.c_void_p(blk.data_ptr())
    with torch.cuda.device(dev):
        for i in range(1, V):
            if view_ready is not None:
                ts.wait_event(view_ready[i])
            n_up = min(cap, i * HW)
            crow = p_counts + 32 * i                                   # counts[i] (8 x int32)
            a.n_upper = n_up
            a.counts_in = p_counts + 32 * (i - 1) + 16                 # counts[i-1][4] = N
            a.counts_out = crow
            a.match = p_match + cap * (i - 1)
            a.v_feats = pf[0] + i * pf[1]; a.v_coords = pc[0] + i * pc[1]; a.v_dens = pd[0] + i * pd[1]; a.v_wemb = pw[0] + i * pw[1]
            a.v_depth = pz[0] + i * pz[1]; a.v_ext = p_ext + 64 * i; a.E_inv = p_einv + 64 * i; a.K_px = p_kpx + 36 * i
            check(L.fs_ptf_match(ra, st), "fs_ptf_match")
            ev.record(ts)
            g.M = n_up; g.flags = int(gru_tc.prepared)
            g.v_feats = a.v_feats; g.v_dens = a.v_dens; g.v_wemb = a.v_wemb; g.M_dev = crow + 8     # counts[i][2] = pairs
            check(L.fs_ptf_gru(rg, st), "fs_ptf_gru")
            gru_tc.prepared = True
            check(L.fs_ptf_pool_update(ra, st), "fs_ptf_pool_update")
            # the order index only needs this step's flags and counters: side stream, under the next steps' kernels
            side.wait_event(ev)
            check(L.fs_ptf_pool_order(C.c_int32(min(cap, (i + 1) * HW)), C.c_void_p(crow), C.c_void_p(phys_ptr[(i - 1) & 1] if i > 1 else None),
                                      C.c_void_p(a.match), blk_ptr, C.c_void_p(phys_ptr[i & 1]), sst), "fs_ptf_pool_order")
        ts.wait_stream(side)
        out = (e(cap, F), e(cap, 3), e(cap), e(cap), e(cap, 16), e(cap))
        check(L.fs_ptf_pool_gather(C.c_int32(cap), C.c_void_p(p_counts + 32 * (V - 1) + 16), C.c_void_p(phys_ptr[(V - 1) & 1]), C.c_int32(F),
                                   *[C.c_void_p(ptr(t)) for t in pool], *[C.c_void_p(ptr(t)) for t in out], st), "fs_ptf_pool_gather")
    for t in (match_all, blk, *phys, *pool, gru_buf, counts):
        t.record_stream(side)
    N = int(counts[V - 1, 4])
    return (out[0][:N], out[1][:N], out[4][:N].reshape(N, 4, 4), out[5][:N])


def fuse_views(gru, feats, coords, dens, wemb, depths, extrinsics, intrinsics, image_shape, depth_thres=0.1,
               E_inv=None, return_debug=False, timings=None, view_ready=None):
    """Flat form: feats [V,HW,F], coords [V,HW,3], dens/wemb [V,HW], depths [V,HW], extrinsics [V,4,4] (c2w),
    intrinsics [V,3,3] (normalised).  Returns (feats [N,F], coords [N,3], ext [N,4,4], depth [N]) (+ debug).
    With autograd enabled and differentiable inputs the result is differentiable w.r.t. feats, coords, dens,
    wemb, depths and the GRU parameters (index decisions are piecewise constant, as in the reference).
    view_ready: optional list of V CUDA events; fold step i (and the initial copy of view 0) waits for view_ready[i] on the
    current stream before touching view i -- the cross-view exchange of the candidates (parallel.ViewExchange) then overlaps
    the fold of the views that have already arrived."""
    L = _lib.lib()
    if not feats.is_cuda:
        raise _lib.FreeSplatB200Error("fuse_gaussians needs CUDA tensors (no CPU fallback exists)")
    dev = feats.device
    f32 = lambda t: t.float().contiguous()

    def per_view(t):
        # the kernels read ONE view at a time: a [V, ...] tensor only needs contiguous slices t[i] (parallel.ViewExchange hands
        # over strided views of its packed per-view blocks; a global .contiguous() would re-copy the whole candidate set)
        t = t.float()
        return t if all(t[i].is_contiguous() for i in range(t.shape[0])) else t.contiguous()
    feats, coords, dens, wemb, depths = map(per_view, (feats, coords, dens, wemb, depths))
    extrinsics, intrinsics = f32(extrinsics), f32(intrinsics)
    V, HW, F = feats.shape
    h, w = image_shape
    assert HW == h * w
    cap = V * HW
    stream = torch.cuda.current_stream(dev).cuda_stream
    need_grad = torch.is_grad_enabled() and (any(t.requires_grad for t in (feats, coords, dens, wemb, depths))
                                             or any(p.requires_grad for p in gru.parameters()))
    fused_gru = (not need_grad) and (not torch.is_grad_enabled() or not any(p.requires_grad for p in gru.parameters())) \
        and _gru_has_reference_structure(gru) and F == gru.mlp_z[2].out_features
    ext16 = extrinsics.detach().reshape(V, 16).contiguous()
    # per-view constants in ONE launch: pixel-space intrinsics (encoder_freesplat.py:445-447) and extrinsic.inverse() (:454)
    # in the canonical arithmetic of oracle/ptf.py::canonical_inverse (fp64 cofactors, one rounding): the projected
    # coordinates feed rounding / z-buffer decisions, so the public path is bit-reproducible.  `E_inv` overrides it (the
    # golden tests pass the inverse the reference's LAPACK build produced).
    K_px = torch.empty((V, 3, 3), dtype=torch.float32, device=dev)
    E_can = torch.empty((V, 4, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(L.fs_ptf_view_setup(C.c_int32(V), C.c_int32(h), C.c_int32(w), C.c_void_p(ptr(ext16)),
                                  C.c_void_p(ptr(intrinsics.detach())), C.c_void_p(ptr(E_can)), C.c_void_p(ptr(K_px)),
                                  C.c_void_p(stream)), "fs_ptf_view_setup")
    E_inv = E_can if E_inv is None else f32(E_inv)
    i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
    counts = torch.zeros((V + 1, 8), dtype=torch.int32, device=dev)
    counts[0, 0] = HW
    counts[0, 4] = HW
    zbuf, pix, zeta = i32(HW), i32(cap), torch.empty(cap, dtype=torch.float32, device=dev)
    match, append = torch.empty(cap, dtype=torch.uint8, device=dev), torch.empty(HW, dtype=torch.uint8, device=dev)
    nb = (cap + 511) // 512 + 1              # blocks of the flag scans (csrc/ptf.cu: kPtfItems = 512)
    block_counts, pair_j, pair_p = i32(3 * nb), i32(cap), i32(cap)
    scratch = (zbuf, pix, zeta, match, append, block_counts, pair_j, pair_p)
    ts = torch.cuda.current_stream(dev)
    if view_ready is not None:
        ts.wait_event(view_ready[0])
    use_tc = fused_gru and GRU_MODE == "tc" and _gru_tc_ok(gru, F)
    train_tc = need_grad and GRU_MODE == "tc" and _gru_has_reference_structure(gru) and _gru_tc_ok(gru, F)
    gru_tc = _GruTc(gru, dev) if (use_tc or train_tc) else None
    # inference with the tensor-core GRU: the whole fold is enqueued without reading a counter back (grids are sized by
    # upper bounds, the kernels take N / M from the device counters); ONE host read at the end returns the final size
    sync_free = use_tc and (not need_grad) and timings is None and not return_debug and SYNC_FREE
    if sync_free and (POOL is True or (POOL == "auto" and V >= 6)):
        return _fold_pool(L, gru_tc, feats, coords, dens, wemb, depths, ext16, E_inv, K_px, h, w, F, V, HW, cap, depth_thres, counts,
                          scratch, stream, view_ready, ts, dev)
    # one unbind per input: the per-view rows are used by several consumers (GRU, merge) -- with feats[i] every use would get its
    # own SelectBackward, i.e. a zero-filled [V,HW,F] gradient and a full-size add per use (1.7 ms of the 3-view training fold)
    fv, xv, dv, wv, zv = (t.unbind(0) for t in (feats, coords, dens, wemb, depths))
    if need_grad:
        state = (fv[0], xv[0], dv[0], wv[0], ext16[0][None].expand(HW, 16).contiguous(), zv[0])
        nxt = None
    else:
        cur, nxt = _State(cap, F, dev), _State(cap, F, dev)
        cur.feats[:HW] = feats[0]; cur.coords[:HW] = coords[0]; cur.dens[:HW] = dens[0]; cur.wemb[:HW] = wemb[0]
        cur.ext[:HW] = ext16[0]; cur.depth[:HW] = depths[0]
        state = (cur.feats, cur.coords, cur.dens, cur.wemb, cur.ext, cur.depth)
    N = HW
    debug = []
    if sync_free:
        # matched pairs of step i: M <= N_i <= i * HW (with exact z-buffer ties several globals match one pixel, so HW is
        # NOT a bound); the buffer and the GRU grid take the same upper bound as the state, CTAs beyond M_dev exit at once
        gru_buf = torch.empty((min(cap, max(V - 1, 1) * HW), F), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            for i in range(1, V):
                if view_ready is not None:
                    ts.wait_event(view_ready[i])
                cin = counts[i - 1, 4:5]
                view = (feats[i], coords[i], dens[i], wemb[i], depths[i], ext16[i], E_inv[i], K_px[i])
                out_bufs = (nxt.feats, nxt.coords, nxt.dens, nxt.wemb, nxt.ext, nxt.depth)
                a = _ptf_args(h, w, F, min(cap, i * HW), depth_thres, state, cin, view, scratch, counts[i], out_bufs)
                check(L.fs_ptf_match(C.byref(a), C.c_void_p(stream)), "fs_ptf_match")
                gru_tc(min(cap, i * HW), pair_j, pair_p, state, feats[i], dens[i], wemb[i], stream, out=gru_buf, M_dev=counts[i, 2:3])
                a.gru_out = ptr(gru_buf)
                check(L.fs_ptf_merge(C.byref(a), C.c_void_p(stream)), "fs_ptf_merge")
                cur, nxt = nxt, cur
                state = (cur.feats, cur.coords, cur.dens, cur.wemb, cur.ext, cur.depth)
        N = int(counts[V - 1, 4]) if V > 1 else HW
        return (state[0][:N], state[1][:N], state[4][:N].reshape(N, 4, 4), state[5][:N])
    with torch.cuda.device(dev):
        for i in range(1, V):
            if view_ready is not None:
                ts.wait_event(view_ready[i])
            cin = counts[i - 1, 4:5]                      # N of the current state, on the device
            view = (fv[i], xv[i], dv[i], wv[i], zv[i], ext16[i], E_inv[i], K_px[i])
            det = tuple(t.detach() for t in state)
            vdet = tuple(t.detach() for t in view)
            out_bufs = None if need_grad else (nxt.feats, nxt.coords, nxt.dens, nxt.wemb, nxt.ext, nxt.depth)
            a = _ptf_args(h, w, F, N, depth_thres, det, cin, vdet, scratch, counts[i], out_bufs)
            if timings is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
            check(L.fs_ptf_match(C.byref(a), C.c_void_p(stream)), "fs_ptf_match")
            c = counts[i].tolist()                         # the one host read of this step
            if timings is not None:
                ev[1].record()
            M, N_out = c[2], c[4]
            gru_out = None
            if M > 0 and train_tc:
                # training: tensor-core forward, hand-derived backward (glue kernels + GEMMs), see _GruTrain
                gru_out = _gru_train(gru, gru_tc, M, pair_j, pair_p, state, fv[i], dv[i], wv[i])
            elif M > 0 and use_tc:
                gru_out = gru_tc(M, pair_j, pair_p, det, fv[i], dv[i], wv[i], stream)
            elif M > 0 and fused_gru:
                gru_out = _gru_fused(gru, M, F, pair_j, pair_p, det, fv[i], dv[i], wv[i], stream)
            elif M > 0:
                pj, pp = pair_j[:M].long(), pair_p[:M].long()
                hidden = state[0][pj]                      # global latent   (networks.py:201 `hidden_feat`)
                inp = fv[i][pp]                         # view-i latent   (`input_feat`)
                e_in = positional_encoding(torch.stack([state[2][pj], wv[i][pp]], -1), 6)
                e_h = positional_encoding(torch.stack([dv[i][pp], state[3][pj]], -1), 6)
                gru_out = gru(inp[None, :, None, :], hidden[None, :, None, :], e_in[None, :, None, :],
                              e_h[None, :, None, :])[0, :, 0, :].float().contiguous()
            if timings is not None:
                ev[2].record()
            if return_debug:
                debug.append(dict(pix=pix[:N].clone(), zeta=zeta[:N].clone(), match=match[:N].clone(),
                                  append=append.clone(), zbuf=zbuf.clone(), counts=c))
            if need_grad:
                meta = (h, w, depth_thres, cin, ext16[i], E_inv[i], K_px[i], scratch, counts[i], c)
                state = _PtfMerge.apply(*(t_ if t_.shape[0] == N else t_[:N] for t_ in state),
                                        fv[i], xv[i], dv[i], wv[i], zv[i], gru_out, meta)
            else:
                a.gru_out = ptr(gru_out)
                check(L.fs_ptf_merge(C.byref(a), C.c_void_p(stream)), "fs_ptf_merge")
                cur, nxt = nxt, cur
                state = (cur.feats, cur.coords, cur.dens, cur.wemb, cur.ext, cur.depth)
            if timings is not None:
                ev[3].record(); torch.cuda.synchronize()
                timings.append(dict(step=i, N_in=c[0], matched=M, N_out=N_out, match_ms=ev[0].elapsed_time(ev[1]),
                                    gru_ms=ev[1].elapsed_time(ev[2]), merge_ms=ev[2].elapsed_time(ev[3])))
            N = N_out
    cut = lambda t_: t_ if t_.shape[0] == N else t_[:N]           # no SliceBackward (zeros + copy) when nothing is cut
    out = (cut(state[0]), cut(state[1]), cut(state[4]).reshape(N, 4, 4), cut(state[5]))
    if return_debug:
        return out, debug, (cut(state[2]), cut(state[3]))
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
