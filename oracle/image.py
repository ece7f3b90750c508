"""TEST INFRASTRUCTURE -- CPU restatement (numpy fp32) of the reference's `prep_image`
(/root/reference/src/misc/image_io.py:36-53): [H,W] | [C,H,W] | [B,C,H,W] float -> uint8 [H, B*W, C'], batch side by side, single
channel repeated to 3, value = uint8(clip(x, 0, 1) * 255) by truncation.  Pinned against outputs of the reference function itself:
tests/golden/image_u8.npz (tests/golden/make_image_golden.py).  Only tests/ and __graft_entry__.smoke() may import this module."""
import numpy as np


def prep_image(image: np.ndarray) -> np.ndarray:
    a = np.asarray(image, np.float32)
    if a.ndim == 2:
        a = a[None, None]
    elif a.ndim == 3:
        a = a[None]
    B, C, H, W = a.shape
    a = a.transpose(1, 2, 0, 3).reshape(C, H, B * W)              # "b c h w -> c h (b w)"
    if C == 1:
        a = np.repeat(a, 3, axis=0)
    q = (np.clip(a, np.float32(0), np.float32(1)) * np.float32(255)).astype(np.float32)
    return np.ascontiguousarray(q.astype(np.uint8).transpose(1, 2, 0))
