"""Evaluation image dump on the B200 kernels.

`prep_image` / `save_image` keep the signatures of the reference's /root/reference/src/misc/image_io.py:36-67 (called for every
context / target / depth image of a test scene at src/model/model_wrapper.py:382-416): float images in [0,1] ->
uint8 HWC (batch side by side, single channel repeated), written as PNG.  The quantisation runs in one kernel (fs_image_u8), so the
device -> host copy moves one byte per sample instead of four; the PNG container is written by PIL when it is installed and by a
minimal zlib writer otherwise.  CPU tensors raise: there is no fallback."""
from __future__ import annotations

import ctypes as C
import struct
import zlib
from pathlib import Path
from typing import Union

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def prep_image(image: torch.Tensor) -> np.ndarray:
    """[H,W] | [C,H,W] | [B,C,H,W] float -> uint8 [H, B*W, 3 or 4] (numpy, host)."""
    if not image.is_cuda:
        raise _lib.FreeSplatB200Error("prep_image needs a CUDA tensor (no CPU fallback exists)")
    if image.ndim == 2:
        image = image[None, None]
    elif image.ndim == 3:
        image = image[None]
    B, Ch, H, W = image.shape
    assert Ch in (1, 3, 4)
    img = image.detach().float().contiguous()
    out = torch.empty((H, B * W, 3 if Ch == 1 else Ch), dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        check(_lib.lib().fs_image_u8(C.c_int32(B), C.c_int32(Ch), C.c_int32(H), C.c_int32(W), C.c_void_p(ptr(img)), C.c_void_p(ptr(out)),
                                     C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)), "fs_image_u8")
    return out.cpu().numpy()


def _png_bytes(a: np.ndarray) -> bytes:
    """Minimal PNG encoder (8-bit RGB / RGBA, filter 0) for hosts without PIL."""
    h, w, c = a.shape
    raw = b"".join(b"\x00" + a[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 6, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def save_image(image: torch.Tensor, path: Union[Path, str]) -> None:
    """Save an image. Assumed to be in range 0-1 (image_io.py:56-67)."""
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    a = prep_image(image)
    try:
        from PIL import Image
        Image.fromarray(a).save(path)
    except ImportError:
        path.write_bytes(_png_bytes(a))
