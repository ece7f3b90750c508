"""Host logic of the direct-binning workspace (freesplat_b200.rasterizer._alloc_state / _fwd_args / the fallback bookkeeping):
no GPU needed -- the buffers are allocated on the CPU and no kernel is launched."""
import torch

from freesplat_b200 import rasterizer


def _state(V=3, H=480, W=640, P=1000, **kw):
    return rasterizer._alloc_state("cpu", P, V, H, W, 9, 2, 1.0, 4096, 1, 9, **kw)


def test_bins_follow_the_tile_grid_and_the_switches(monkeypatch):
    monkeypatch.setattr(rasterizer, "BIN_CAP", 4096)
    st = _state()
    nt = 3 * 40 * 30
    assert st.bin_cap == 4096 and st.bins.numel() == nt * 4096 and st.bins.dtype == torch.int64
    # [counters | cursors | flag, ready, 2 spare]: the two words behind the cursors are the bin-overflow flag and the scan-ready word
    assert st.tile_buf.numel() == 2 * nt + 4 and st.status.numel() == 4 and st.status.data_ptr() != st.tile_buf[2 * nt:].data_ptr()
    a = rasterizer._fwd_args(st, torch.zeros(1000, 3), torch.zeros(1000), torch.zeros(3, 48), torch.zeros(1000, 9, 3), None, None, None,
                             torch.zeros(1000, 9))
    assert a.bin_cap == 4096 and a.bins == st.bins.data_ptr() and a.tile_cursor == a.tile_count + 4 * nt
    for name, val in (("BIN_CAP", 0), ("BIN_CAP", 8192), ("FUSED_SCAN", True), ("RENDER_PACKED", True)):
        with monkeypatch.context() as m:
            m.setattr(rasterizer, name, val)
            off = _state()
            assert off.bins is None and off.bin_cap == 0, name
            assert rasterizer._fwd_args(off, torch.zeros(1000, 3), torch.zeros(1000), torch.zeros(3, 48), torch.zeros(1000, 9, 3), None,
                                        None, None, torch.zeros(1000, 9)).bins in (None, 0)
    assert _state(bins_ok=False).bins is None
    # more than 1 GiB of bins (here: 64 views of 1920 x 1080): the call keeps the scatter pass instead
    assert _state(V=64, H=1080, W=1920).bins is None


def test_a_reported_fallback_switches_the_bins_off_for_that_shape(monkeypatch):
    monkeypatch.setattr(rasterizer, "BIN_CAP", 2048)
    rasterizer.reset_capacity_hints()
    key = ("cpu", 1000, 3, 480, 640)
    rasterizer._scratch_cache[(key, 4096, 9, 0)] = object()
    rasterizer._scratch_cache[(("cpu", 5, 1, 16, 16), 4096, 9, 0)] = object()
    rasterizer._note_bin_fallback(key)
    assert key in rasterizer._bins_off
    assert [k[0] for k in rasterizer._scratch_cache] == [("cpu", 5, 1, 16, 16)]      # cached workspaces of that shape are dropped
    rasterizer._scratch_cache.clear()
    rasterizer.reset_capacity_hints()
    assert not rasterizer._bins_off
