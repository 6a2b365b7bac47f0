"""Pixel stages of the hot path (oracle; test infrastructure): greyscale conversion, constant
pad, half-pixel bilinear resize, threshold.  numpy float32, IEEE operation order as in the
reference (no FMA contraction)."""
from __future__ import annotations

import numpy as np

from . import BLACK_VALUE

F = np.float32
ITU_WEIGHTS = np.array([0.299, 0.587, 0.114], dtype=np.float32)  # preprocess.rs:171


class ImageSourceError(ValueError):
    """preprocess.rs:37-46"""

    UNSUPPORTED_CHANNEL_COUNT = "channel count is not 1, 3 or 4"
    INVALID_DATA_LENGTH = "data length is not a multiple of `width * height`"


def image_source_from_bytes(data: bytes | np.ndarray, width: int, height: int) -> np.ndarray:
    """`ImageSource::from_bytes` (preprocess.rs:81-102): returns an HWC u8 view."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1)
    channel_len = int(width) * int(height)
    if channel_len == 0:
        raise ImageSourceError(ImageSourceError.UNSUPPORTED_CHANNEL_COUNT)
    if buf.size % channel_len != 0:
        raise ImageSourceError(ImageSourceError.INVALID_DATA_LENGTH)
    channels = buf.size // channel_len
    if channels not in (1, 3, 4):
        raise ImageSourceError(ImageSourceError.UNSUPPORTED_CHANNEL_COUNT)
    return buf.reshape(height, width, channels)


def check_image_source(arr: np.ndarray, order: str) -> None:
    """`ImageSource::from_tensor` (preprocess.rs:105-123)."""
    if arr.ndim != 3:
        raise ImageSourceError("expected a 3-D tensor")
    channels = arr.shape[2] if order == "hwc" else arr.shape[0]
    if channels not in (1, 3, 4):
        raise ImageSourceError(ImageSourceError.UNSUPPORTED_CHANNEL_COUNT)


def prepare_image(arr: np.ndarray, order: str = "hwc") -> np.ndarray:
    """`prepare_image` / `convert_pixels` (preprocess.rs:149-248).

    u8 or f32 input, HWC or CHW, 1/3/4 channels -> f32 [1, H, W] =
    ((BLACK_VALUE + c0*w0) + c1*w1) + c2*w2, alpha ignored; for u8 the weights are
    ITU/255 computed in f32 (preprocess.rs:182)."""
    check_image_source(arr, order)
    if order == "chw":
        arr = np.moveaxis(arr, 0, 2)
    h, w, c = arr.shape
    if arr.dtype == np.uint8:
        weights = (ITU_WEIGHTS / F(255.0)).astype(np.float32) if c != 1 else np.array([F(1.0) / F(255.0)], np.float32)
    elif arr.dtype == np.float32:
        weights = ITU_WEIGHTS if c != 1 else np.array([1.0], np.float32)
    else:
        raise TypeError("pixels must be u8 or f32")
    out = np.full((h, w), F(BLACK_VALUE), dtype=np.float32)
    for ci in range(len(weights)):
        out = out + arr[:, :, ci].astype(np.float32) * weights[ci]
    return out[None, :, :]


def pad_bottom_right(img: np.ndarray, pad_bottom: int, pad_right: int, value: float = BLACK_VALUE) -> np.ndarray:
    """ONNX `Pad` constant mode on the last two axes (call detection.rs:155-164)."""
    pads = [(0, 0)] * (img.ndim - 2) + [(0, pad_bottom), (0, pad_right)]
    return np.pad(img, pads, mode="constant", constant_values=F(value))


def _axis_coords(n_in: int, n_out: int):
    scale = F(n_in) / F(n_out)  # inverse scale, f32
    d = np.arange(n_out, dtype=np.float32)
    src = scale * (d + F(0.5)) - F(0.5)
    src = np.clip(src, F(0.0), F(n_in - 1))
    i0 = src.astype(np.int64)  # trunc; src >= 0
    i1 = np.minimum(i0 + 1, n_in - 1)
    w = (src - i0.astype(np.float32)).astype(np.float32)
    return i0, i1, w


def resize_bilinear(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """rten `resize_image` = ONNX Resize(mode=linear, half_pixel), no antialias
    (calls detection.rs:168,194; recognition.rs:121).  Pinned by lib.rs:437-445.

    src = (dst + 0.5) * (in/out) - 0.5 clamped to [0, in-1]; lerp along x first, then y:
        top = (1-wx)*tl + wx*tr ; bot = (1-wx)*bl + wx*br ; out = (1-wy)*top + wy*bot
    Works on the last two axes."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    in_h, in_w = img.shape[-2:]
    if (in_h, in_w) == (out_h, out_w):
        return img.copy()
    y0, y1, wy = _axis_coords(in_h, out_h)
    x0, x1, wx = _axis_coords(in_w, out_w)
    one = F(1.0)
    r0 = img[..., y0, :]
    r1 = img[..., y1, :]
    top = (one - wx) * r0[..., x0] + wx * r0[..., x1]
    bot = (one - wx) * r1[..., x0] + wx * r1[..., x1]
    wy_ = wy[:, None]
    return ((one - wy_) * top + wy_ * bot).astype(np.float32)


def threshold_mask(prob: np.ndarray, threshold: float = 0.2) -> np.ndarray:
    """detection.rs:110 (strict `>`; threshold compared in f32)."""
    return prob > F(threshold)
