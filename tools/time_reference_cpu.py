"""Times the REFERENCE's own PyTorch code for the cost volume and PTF on this container's host cores, on a
full, unscaled unit of work each: one reference view of the cost volume (K = 2, 48 x 120 x 160, D = 128 planes:
AVGFeatureVolumeManager.build_cost_volume, cost_volume.py:429-619) and the 3-view 640 x 480 PTF fold
(EncoderFreeSplat.fuse_gaussians, encoder_freesplat.py:431-522).  /root/reference does not exist on the GPU box, so the
result is committed (profiles/r2_reference_cpu_timings.json) and quoted by bench.py next to the GPU numbers.
    python tools/time_reference_cpu.py > profiles/r2_reference_cpu_timings.json"""
import json
import os
import sys
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freesplat_b200 import synth  # noqa: E402
from tests.golden import ref_loader  # noqa: E402


def main():
    n = len(os.sched_getaffinity(0))
    torch.set_num_threads(n)
    out = {"host_threads": n, "torch": torch.__version__, "where": "build container (no GPU)"}
    cvmod = ref_loader.load_cost_volume_module()
    V, K, Hf, Wf, D = 3, 2, 120, 160, 128
    inp = synth.cost_volume_inputs(0, V, K, 48, Hf, Wf)
    one = {k: (v[:1] if v.shape[0] == V else v) for k, v in inp.items()}
    m = cvmod.AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, mlp_channels=[49, 32, 32, 1], matching_dim_size=48)
    with torch.no_grad():
        t0 = time.perf_counter()
        m.build_cost_volume(**one)
        dt = time.perf_counter() - t0
    out["cost_volume_fwd_one_view_K2_D128_120x160_s"] = dt
    cur = one["cur_feats"].clone().requires_grad_(True); src = one["src_feats"].clone().requires_grad_(True)
    t0 = time.perf_counter()
    vol, _, _ = m.build_cost_volume(**{**one, "cur_feats": cur, "src_feats": src})
    vol.sum().backward()
    out["cost_volume_fwd_bwd_one_view_s"] = time.perf_counter() - t0
    fuse, pe, GRU = ref_loader.load_fuse_gaussians()
    pin = synth.ptf_inputs(0, 3, 480, 640)
    gru = GRU(); gru.load_state_dict(synth.gru_state(0))
    with torch.no_grad():
        t0 = time.perf_counter()
        r = fuse(SimpleNamespace(gru=gru), pin["gaussians"], pin["coords"], pin["densities"], pin["weight_emb"], pin["depths"],
                 pin["extrinsics"], pin["intrinsics"], pin["image_shape"])
        out["ptf_3views_640x480_s"] = time.perf_counter() - t0
    out["ptf_3views_N_out"] = int(r[0].shape[1])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
