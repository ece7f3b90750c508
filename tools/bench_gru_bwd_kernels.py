"""CUDA-event time of every launch shape of the GRU backward products (fs_ptf_gru_bwd_data / _weights) at M pairs (default 180 000
= one fold step of BASELINE config 3), each with the fp32 torch matmul it replaces.   python tools/bench_gru_bwd_kernels.py [M]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from freesplat_b200 import _lib, ptf
dev = torch.device("cuda", 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 180000
L = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
r = lambda *s: torch.randn(*s, device=dev)
res = {"M": M}
A, W64, W152, W176, mask, A1, rl = r(M, 64), r(64, 64), r(64, 152), r(64, 176), r(M, 64), r(M, 176), r(M, 64)
o64, o152, o176, dr = r(M, 64), r(M, 152), r(M, 176), r(M, 64)
with torch.cuda.device(dev):
    t = lambda f: round(bench.gpu_ms(f, n=10, warm=3) * 1e3, 1)
    res["data_N64_mask_us"] = t(lambda: ptf._bwd_data(L, st, A, W64, 64, o64, mode=1, mask=mask))
    res["data_N152_store_us"] = t(lambda: ptf._bwd_data(L, st, A, W152, 152, o152, mode=0))
    res["data_N152_gate_us"] = t(lambda: ptf._bwd_data(L, st, A, W152, 152, o176, mode=3, h=A1, r_lin=rl, dr_lin=dr))
    res["data_N176_accumulate_us"] = t(lambda: ptf._bwd_data(L, st, A, W176, 176, o176, mode=2))
    res["torch_fp32_N176_addmm_us"] = t(lambda: o176.addmm_(A, W176))
    Y0, Y1, X64a, X64b, U = r(M, 64), r(M, 64), r(M, 64), r(M, 64), r(M, 152)
    res["weights_224_us"] = t(lambda: ptf._bwd_weights(L, st, Y0, Y1, X64a, U, torch.empty(128, 224, device=dev)))
    res["weights_144_us"] = t(lambda: ptf._bwd_weights(L, st, Y0, Y1, X64a, X64b, torch.empty(128, 144, device=dev)))
    res["weights_192_us"] = t(lambda: ptf._bwd_weights(L, st, Y0, Y1, A1, None, torch.empty(128, 192, device=dev)))
    res["torch_fp32_176_gemm_us"] = t(lambda: Y0.t() @ A1)
print(json.dumps(res))
