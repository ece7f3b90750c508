"""CPU: how reproducible is the GRU gradient of the PTF training fold in fp32 at all?

The reference's own op sequence (positional encodings, three Linear-ReLU-Linear MLPs, gates: networks.py:188-214,
encoder_freesplat.py:485-491) is evaluated in fp32 and in fp64 on the same 60 000 matched pairs.  A ReLU pre-activation within
rounding distance of 0 flips between the two evaluations for a handful of (pair, unit) combinations and moves that pair's
gradient by O(1): single elements differ by > 1e-3 of the tensor's maximum and the first-layer weight sums by > 1e-4, while all
other elements agree to 1e-6.  This is why tests/test_ptf_gpu.py counts such elements instead of demanding 1e-4 everywhere: no
fp32 implementation with another summation order (ours: tensor-core forward, hand-derived backward) can do better against
the reference's fp32 result than the reference does against its own fp64 evaluation."""
import torch

from freesplat_b200 import synth


def test_reference_fp32_gradient_deviates_from_fp64_at_relu_kinks():
    torch.manual_seed(0)
    F, M, N, HW = 64, 60000, 80000, 76800
    mk = lambda din: torch.nn.Sequential(torch.nn.Linear(din, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64))

    class G(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mlp_z, self.mlp_r, self.mlp_n = mk(176), mk(176), mk(152)

    def pe(a, b):
        pos = torch.stack([a, b], -1)
        pts = (pos[..., None] * (2 ** torch.arange(6)).to(a.dtype)).reshape(pos.shape[:-1] + (12,))
        return torch.stack([torch.sin(pts), torch.cos(pts)], -1).reshape(pts.shape[:-1] + (24,))
    base = dict(feats=torch.randn(N, F), dens=torch.rand(N) * 10 + 0.5, wemb=torch.rand(N) * 5, vf=torch.randn(HW, F),
                vd=torch.rand(HW) * 10 + 0.5, vw=torch.rand(HW) * 5)
    pj, pp, g = torch.randperm(N)[:M], torch.randperm(HW)[:M], torch.randn(M, F)
    out = {}
    for dt in (torch.float32, torch.float64):
        gru = G(); gru.load_state_dict(synth.gru_state(5)); gru = gru.to(dt)
        t = {k: v.to(dt).clone().requires_grad_(True) for k, v in base.items()}
        hidden, inp = t["feats"][pj], t["vf"][pp]
        x1 = torch.cat((inp, pe(t["dens"][pj], t["vw"][pp])), -1)
        cat = torch.cat((hidden, pe(t["vd"][pp], t["wemb"][pj]), x1), -1)
        r, z = torch.sigmoid(gru.mlp_r(cat)), torch.sigmoid(gru.mlp_z(cat))
        q = torch.tanh(gru.mlp_n(torch.cat((r * hidden, x1), -1)))
        ((1 - z) * hidden + z * q).backward(g.to(dt))
        out[dt] = {**{k: v.grad.double() for k, v in t.items()}, **{n: p.grad.double() for n, p in gru.named_parameters()}}
    worst_elem, worst_param, frac = 0.0, 0.0, 0.0
    for k, b in out[torch.float64].items():
        err = (out[torch.float32][k] - b).abs()
        sc = float(b.abs().max())
        if k.startswith("mlp_"):
            worst_param = max(worst_param, float(err.max()) / sc)
        else:
            worst_elem = max(worst_elem, float(err.max()) / sc)
            frac = max(frac, float((err > 1e-4 * b.abs() + 1e-5 * sc).double().mean()))
    assert worst_elem > 1e-3 and worst_param > 1e-4          # the reference's own fp32 result is not reproducible to 1e-4 ...
    assert frac < 1e-3 and worst_elem < 0.2                   # ... but only on a handful of flipped pairs
