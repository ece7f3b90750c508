"""Evaluation image dump: the oracle against the reference's own prep_image (golden), the PNG writer
(host logic) on the CPU; the fs_image_u8 kernel against the golden on the GPU, bit for bit."""
import os
import struct
import zlib

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "image_u8.npz")
CASES = ("rgb", "batch", "gray", "rgba", "edges")


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_prep_image(name):
    from oracle import image as oimg
    z = np.load(GOLD)
    assert np.array_equal(oimg.prep_image(z["in_" + name]), z["out_" + name])


def test_png_writer_roundtrip():
    from freesplat_b200.image_io import _png_bytes
    a = (np.arange(5 * 7 * 3) % 251).astype(np.uint8).reshape(5, 7, 3)
    b = _png_bytes(a)
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    # walk the chunks: IHDR fields, CRCs, and the inflated scanlines (filter byte 0 + raw row)
    pos, idat, seen = 8, b"", []
    while pos < len(b):
        n, tag = struct.unpack(">I4s", b[pos:pos + 8])
        data = b[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        seen.append(tag)
        if tag == b"IHDR":
            assert struct.unpack(">IIBBBBB", data) == (7, 5, 8, 2, 0, 0, 0)
        if tag == b"IDAT":
            idat += data
        pos += 12 + n
    assert seen == [b"IHDR", b"IDAT", b"IEND"]
    raw = zlib.decompress(idat)
    rows = [raw[y * (1 + 21) + 1:(y + 1) * (1 + 21)] for y in range(5)]
    assert all(raw[y * 22] == 0 for y in range(5)) and b"".join(rows) == a.tobytes()
    try:
        import io
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(b))), a)
    except ImportError:
        pass


def test_cpu_tensor_raises():
    from freesplat_b200 import _lib
    from freesplat_b200.image_io import prep_image
    with pytest.raises(_lib.FreeSplatB200Error):
        prep_image(torch.zeros(3, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_prep_image(name, tmp_path):
    from freesplat_b200.image_io import prep_image, save_image
    z = np.load(GOLD)
    img = torch.from_numpy(z["in_" + name]).to("cuda:0")
    got = prep_image(img)
    assert got.dtype == np.uint8 and np.array_equal(got, z["out_" + name])
    save_image(img, tmp_path / "sub" / f"{name}.png")
    try:
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(tmp_path / "sub" / f"{name}.png")), z["out_" + name])
    except ImportError:
        assert (tmp_path / "sub" / f"{name}.png").stat().st_size > 60
