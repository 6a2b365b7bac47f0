"""Host-side mirror of the reference's public API (ocrs/src/lib.rs) over the C ABI.

Same names, argument meaning and error behaviour as the Rust crate so that the parity tests read
like the reference's own tests:

    engine = OcrEngine(OcrEngineParams(detection_model=..., recognition_model=...))
    inp    = engine.prepare_input(ImageSource.from_bytes(data, (w, h)))
    words  = engine.detect_words(inp)
    lines  = engine.find_text_lines(inp, words)
    texts  = engine.recognize_text(inp, lines)

Everything numerical happens inside libocrs_b200.so on the GPU; this module only marshals."""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import EngineParamsC, OcrsError, PageC, PoolParamsC, RectC, RotatedRectC, TextResultC, check, lib

# lib.rs:34
DEFAULT_ALPHABET = " 0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~EABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"


class DimOrder(enum.IntEnum):
    """preprocess.rs:50-57"""
    Hwc = 0
    Chw = 1


class DecodeMethod(enum.IntEnum):
    """recognition.rs:199-205"""
    Greedy = 0
    BeamSearch = 1


class ImageSourceError(ValueError):
    """preprocess.rs:37-46"""


class ImageSource:
    """preprocess.rs:61-124: a validated view of caller pixels (u8 or f32; HWC or CHW)."""

    def __init__(self, data: np.ndarray, order: DimOrder):
        self.data = data
        self.order = order

    @staticmethod
    def from_bytes(data, dimensions: Tuple[int, int]) -> "ImageSource":
        """`ImageSource::from_bytes(bytes, (width, height))` (preprocess.rs:81-102)."""
        width, height = dimensions
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1)
        channel_len = int(width) * int(height)
        if channel_len == 0:
            raise ImageSourceError("channel count is not 1, 3 or 4")
        if buf.size % channel_len != 0:
            raise ImageSourceError("data length is not a multiple of `width * height`")
        return ImageSource.from_tensor(buf.reshape(height, width, buf.size // channel_len), DimOrder.Hwc)

    @staticmethod
    def from_tensor(data: np.ndarray, order: DimOrder) -> "ImageSource":
        """`ImageSource::from_tensor` (preprocess.rs:105-123)."""
        if data.ndim != 3 or data.dtype not in (np.uint8, np.float32):
            raise ImageSourceError("expected a 3-D u8 or f32 tensor")
        channels = data.shape[2] if order == DimOrder.Hwc else data.shape[0]
        if channels not in (1, 3, 4):
            raise ImageSourceError("channel count is not 1, 3 or 4")
        return ImageSource(np.ascontiguousarray(data), DimOrder(order))


@dataclass
class RotatedRect:
    """rten_imageproc::RotatedRect as plain floats (centre, unit up axis, width, height)."""
    cx: float
    cy: float
    ux: float
    uy: float
    w: float
    h: float

    def raw(self):
        return (self.cx, self.cy, self.ux, self.uy, self.w, self.h)

    def rounded_vertices(self) -> List[List[int]]:
        """`rounded_vertex_coords` (ocrs-cli/src/output.rs:24-27): the four corners as [x, y]."""
        xy = (C.c_int32 * 8)()
        rr = RotatedRectC(*self.raw())
        check(lib.ocrs_b200_rotated_rect_vertices(C.byref(rr), xy))
        return [[xy[2 * i], xy[2 * i + 1]] for i in range(4)]


@dataclass
class Rect:
    top: int
    left: int
    bottom: int
    right: int

    def tlbr(self):
        return (self.top, self.left, self.bottom, self.right)


@dataclass
class TextChar:
    """text_items.rs:47-53"""
    char: str
    rect: Rect


def _rects_c(chars: Sequence["TextChar"]):
    arr = (RectC * max(len(chars), 1))()
    for i, c in enumerate(chars):
        arr[i] = RectC(c.rect.top, c.rect.left, c.rect.bottom, c.rect.right)
    return arr


class TextItem:
    """`trait TextItem` (text_items.rs:8-31): a non-empty run of recognised characters."""
    chars: List[TextChar]

    def __str__(self) -> str:
        return "".join(c.char for c in self.chars)

    def bounding_rect(self) -> Rect:
        """text_items.rs:13-15"""
        return Rect(min(c.rect.top for c in self.chars), min(c.rect.left for c in self.chars),
                    max(c.rect.bottom for c in self.chars), max(c.rect.right for c in self.chars))

    def rotated_rect(self) -> RotatedRect:
        """text_items.rs:18-30: min-area rect of all character corners, oriented upright."""
        out = RotatedRectC()
        check(lib.ocrs_b200_text_item_rotated_rect(_rects_c(self.chars), len(self.chars), C.byref(out)))
        return RotatedRect(out.cx, out.cy, out.ux, out.uy, out.w, out.h)


@dataclass
class TextWord(TextItem):
    """text_items.rs:92-101"""
    chars: List[TextChar]

    def __post_init__(self):
        assert self.chars, "Text words must not be empty"

    __str__ = TextItem.__str__


@dataclass
class TextLine(TextItem):
    """text_items.rs:59-82"""
    chars: List[TextChar]

    def __post_init__(self):
        assert self.chars, "Text lines must not be empty"

    __str__ = TextItem.__str__

    def words(self) -> List[TextWord]:
        """text_items.rs:76-82: split on ' ', empty pieces dropped."""
        out, cur = [], []
        for c in self.chars:
            if c.char == " ":
                if cur:
                    out.append(TextWord(cur))
                cur = []
            else:
                cur.append(c)
        if cur:
            out.append(TextWord(cur))
        return out


def _to_text_result(lines: Sequence[Optional[TextLine]]):
    """Flattens `[Option<TextLine>]` into the C ABI's ocrs_b200_text_result (arrays kept alive by the tuple)."""
    n = len(lines)
    total = sum(len(l.chars) for l in lines if l is not None)
    present = (C.c_uint8 * max(n, 1))()
    offs = (C.c_int64 * (n + 1))()
    chars = (C.c_uint32 * max(total, 1))()
    rects = (RectC * max(total, 1))()
    k = 0
    for i, l in enumerate(lines):
        offs[i] = k
        present[i] = 0 if l is None else 1
        if l is None:
            continue
        for c in l.chars:
            chars[k] = ord(c.char)
            rects[k] = RectC(c.rect.top, c.rect.left, c.rect.bottom, c.rect.right)
            k += 1
    offs[n] = k
    res = TextResultC(n, C.cast(present, C.POINTER(C.c_uint8)), C.cast(offs, C.POINTER(C.c_int64)),
                      C.cast(chars, C.POINTER(C.c_uint32)), C.cast(rects, C.POINTER(RectC)))
    return res, (present, offs, chars, rects)


def _take_string(ptr: C.c_void_p) -> str:
    try:
        return C.string_at(ptr).decode("utf-8")
    finally:
        lib.ocrs_b200_free(ptr)


def format_text_output(text_lines: Sequence[Optional[TextLine]]) -> str:
    """`format_text_output` (ocrs-cli/src/output.rs:87-94): recognised lines joined with '\n'."""
    res, keep = _to_text_result(text_lines)
    out = C.c_void_p()
    check(lib.ocrs_b200_format_text_output(C.byref(res), C.byref(out)))
    return _take_string(out)


def format_json_output(input_path: str, input_hw: Sequence[int], text_lines: Sequence[Optional[TextLine]]) -> str:
    """`format_json_output` (ocrs-cli/src/output.rs:97-100): HierText-style JSON document with one
    paragraph holding every line, words split on spaces, vertices = rounded rotated-rect corners."""
    res, keep = _to_text_result(text_lines)
    out = C.c_void_p()
    check(lib.ocrs_b200_format_json_output(C.byref(res), input_path.encode("utf-8"), int(input_hw[0]), int(input_hw[1]),
                                           C.byref(out)))
    return _take_string(out)


@dataclass
class OcrEngineParams:
    """lib.rs:37-71.  Models are `.onnx` file images (bytes) or paths."""
    detection_model: Optional[object] = None
    recognition_model: Optional[object] = None
    debug: bool = False
    decode_method: DecodeMethod = DecodeMethod.Greedy
    beam_width: int = 100
    alphabet: Optional[str] = None
    allowed_chars: Optional[str] = None
    device: int = 0


def _model_bytes(m) -> Optional[bytes]:
    if m is None:
        return None
    if isinstance(m, (bytes, bytearray, memoryview)):
        return bytes(m)
    with open(m, "rb") as fp:
        return fp.read()


def _rects_to_c(rects: Sequence[RotatedRect]):
    arr = (RotatedRectC * max(len(rects), 1))()
    for i, r in enumerate(rects):
        arr[i] = RotatedRectC(*r.raw())
    return arr


def _rects_from_c(ptr, n: int) -> List[RotatedRect]:
    return [RotatedRect(ptr[i].cx, ptr[i].cy, ptr[i].ux, ptr[i].uy, ptr[i].w, ptr[i].h) for i in range(n)]


def _text_result(res_ptr) -> List[Optional[TextLine]]:
    r = res_ptr.contents
    out: List[Optional[TextLine]] = []
    for i in range(r.n_lines):
        if not r.line_present[i]:
            out.append(None)
            continue
        chars = []
        for k in range(r.char_offsets[i], r.char_offsets[i + 1]):
            rc = r.char_rects[k]
            chars.append(TextChar(chr(r.chars[k]), Rect(rc.top, rc.left, rc.bottom, rc.right)))
        out.append(TextLine(chars))
    return out


class OcrInput:
    """lib.rs:125-128: the prepared greyscale page, resident in GPU memory."""

    def __init__(self, engine: "OcrEngine", handle):
        self._engine = engine
        self._h = handle

    @property
    def shape(self) -> Tuple[int, int, int]:
        h, w = C.c_int(), C.c_int()
        check(lib.ocrs_b200_input_shape(self._h, C.byref(h), C.byref(w)))
        return (1, h.value, w.value)

    def image(self) -> np.ndarray:
        """Host copy of the CHW f32 tensor (`OcrInput::image`)."""
        _, h, w = self.shape
        out = np.empty((1, h, w), dtype=np.float32)
        check(lib.ocrs_b200_input_read(self._engine._h, self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib.ocrs_b200_input_destroy(self._h)
            self._h = None


class Model:
    """The inner seam (`trait Model`, model.rs:6-17): input_shape() + run()."""

    def __init__(self, model, device: int = 0):
        data = _model_bytes(model)
        self._h = C.c_void_p()
        self._buf = data
        check(lib.ocrs_b200_model_load(data, len(data), device, C.byref(self._h)))

    def input_shape(self) -> List[object]:
        dims = (C.c_int64 * 8)()
        nd = C.c_int()
        check(lib.ocrs_b200_model_input_shape(self._h, dims, C.byref(nd)))
        return [int(dims[i]) if dims[i] >= 0 else "sym" for i in range(nd.value)]

    def run(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        shape = (C.c_int64 * 8)(*x.shape)
        out = C.POINTER(C.c_float)()
        oshape = (C.c_int64 * 8)()
        ond = C.c_int()
        check(lib.ocrs_b200_model_run(self._h, x.ctypes.data_as(C.c_void_p), shape, x.ndim, C.byref(out), oshape,
                                      C.byref(ond)))
        shp = tuple(int(oshape[i]) for i in range(ond.value))
        n = int(np.prod(shp)) if shp else 1
        res = np.ctypeslib.as_array(out, shape=(n,)).copy().reshape(shp)
        lib.ocrs_b200_free(out)
        return res

    def last_flops(self) -> float:
        return float(lib.ocrs_b200_model_last_flops(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.ocrs_b200_model_destroy(self._h)
            self._h = None


def _fill_engine_params(p: EngineParamsC, params: "OcrEngineParams"):
    """Fills the C struct; returns the objects that must stay alive while the struct is in use."""
    det = _model_bytes(params.detection_model)
    rec = _model_bytes(params.recognition_model)
    p.detection_model = C.cast(C.c_char_p(det), C.c_void_p) if det else None
    p.detection_model_len = len(det) if det else 0
    p.recognition_model = C.cast(C.c_char_p(rec), C.c_void_p) if rec else None
    p.recognition_model_len = len(rec) if rec else 0
    p.debug = int(params.debug)
    p.decode_method = int(params.decode_method)
    p.beam_width = int(params.beam_width)
    p.alphabet_utf8 = params.alphabet.encode("utf-8") if params.alphabet is not None else None
    p.allowed_chars_utf8 = params.allowed_chars.encode("utf-8") if params.allowed_chars is not None else None
    p.device = int(params.device)
    return det, rec


class OcrEngine:
    """lib.rs:111-301"""

    def __init__(self, params: Optional[OcrEngineParams], _handle=None, _alphabet=None):
        if _handle is not None:  # a pool's worker engine (OcrPool.engine): same handle type, shared ownership
            self._h = _handle
            self.alphabet = _alphabet if _alphabet is not None else DEFAULT_ALPHABET
            return
        p = EngineParamsC()
        self._det, self._rec = _fill_engine_params(p, params)
        self._h = C.c_void_p()
        check(lib.ocrs_b200_engine_create(C.byref(p), C.byref(self._h)))
        self.alphabet = params.alphabet if params.alphabet is not None else DEFAULT_ALPHABET

    def __del__(self):
        if getattr(self, "_h", None):
            lib.ocrs_b200_engine_destroy(self._h)
            self._h = None

    # -- lib.rs:183
    def prepare_input(self, image: ImageSource) -> OcrInput:
        a = image.data
        if image.order == DimOrder.Hwc:
            h, w, c = a.shape
        else:
            c, h, w = a.shape
        dtype = 0 if a.dtype == np.uint8 else 1
        out = C.c_void_p()
        check(lib.ocrs_b200_engine_prepare_input(self._h, a.ctypes.data_as(C.c_void_p), dtype, int(image.order), h, w, c,
                                                 C.byref(out)))
        return OcrInput(self, out)

    def prepare_input_device(self, device_ptr: int, dtype: int, order: DimOrder, h: int, w: int, c: int) -> OcrInput:
        """Pixels already in this GPU's memory (used to time the HBM-resident path)."""
        out = C.c_void_p()
        check(lib.ocrs_b200_engine_prepare_input_device(self._h, C.c_void_p(device_ptr), dtype, int(order), h, w, c,
                                                        C.byref(out)))
        return OcrInput(self, out)

    # -- lib.rs:193
    def detect_words(self, inp: OcrInput) -> List[RotatedRect]:
        rects = C.POINTER(RotatedRectC)()
        n = C.c_size_t()
        check(lib.ocrs_b200_engine_detect_words(self._h, inp._h, C.byref(rects), C.byref(n)))
        out = _rects_from_c(rects, n.value)
        lib.ocrs_b200_free(rects)
        return out

    def detect_words_batch(self, inputs: Sequence[OcrInput]) -> List[List[RotatedRect]]:
        hs = (C.c_void_p * max(len(inputs), 1))(*[i._h for i in inputs])
        rects = C.POINTER(RotatedRectC)()
        offs = C.POINTER(C.c_size_t)()
        check(lib.ocrs_b200_engine_detect_words_batch(self._h, hs, len(inputs), C.byref(rects), C.byref(offs)))
        out = []
        for i in range(len(inputs)):
            s, e = offs[i], offs[i + 1]
            out.append([RotatedRect(rects[k].cx, rects[k].cy, rects[k].ux, rects[k].uy, rects[k].w, rects[k].h)
                        for k in range(s, e)])
        lib.ocrs_b200_free(rects)
        lib.ocrs_b200_free(offs)
        return out

    # -- lib.rs:207
    def detect_text_pixels(self, inp: OcrInput) -> np.ndarray:
        _, h, w = inp.shape
        out = np.empty((h, w), dtype=np.float32)
        check(lib.ocrs_b200_engine_detect_text_pixels(self._h, inp._h, out.ctypes.data_as(C.c_void_p)))
        return out

    # -- lib.rs:222
    def find_text_lines(self, inp: Optional[OcrInput], words: Sequence[RotatedRect]) -> List[List[RotatedRect]]:
        return find_text_lines(words)

    # -- lib.rs:237
    def recognize_text(self, inp: OcrInput, lines: Sequence[Sequence[RotatedRect]]) -> List[Optional[TextLine]]:
        flat = [w for line in lines for w in line]
        arr = _rects_to_c(flat)
        offs = (C.c_size_t * (len(lines) + 1))()
        k = 0
        for i, line in enumerate(lines):
            offs[i] = k
            k += len(line)
        offs[len(lines)] = k
        res = C.POINTER(TextResultC)()
        check(lib.ocrs_b200_engine_recognize_text(self._h, inp._h, arr, offs, len(lines), C.byref(res)))
        out = _text_result(res)
        lib.ocrs_b200_text_result_free(res)
        return out

    # -- lib.rs:268
    def prepare_recognition_input(self, inp: OcrInput, line: Sequence[RotatedRect]) -> np.ndarray:
        arr = _rects_to_c(line)
        out = C.POINTER(C.c_float)()
        h, w = C.c_int(), C.c_int()
        check(lib.ocrs_b200_engine_prepare_recognition_input(self._h, inp._h, arr, len(line), C.byref(out), C.byref(h),
                                                             C.byref(w)))
        res = np.ctypeslib.as_array(out, shape=(h.value * w.value,)).copy().reshape(h.value, w.value)
        lib.ocrs_b200_free(out)
        return res

    # -- ocrs-cli debug outputs, batched (main.rs:423-443)
    def detect_text_pixels_batch(self, inputs: Sequence[OcrInput], maps: bool = True, masks: bool = True):
        """`--text-map` / `--text-mask` for a batch of pages in one detection pass: returns (maps, masks), lists of
        f32 [H, W] probability maps and u8 [H, W] masks (x > detection_threshold); a list is None when not requested."""
        n = len(inputs)
        hs = (C.c_void_p * max(n, 1))(*[i._h for i in inputs])
        mp = (C.c_void_p * max(n, 1))() if maps else None
        mk = (C.c_void_p * max(n, 1))() if masks else None
        check(lib.ocrs_b200_engine_detect_text_pixels_batch(self._h, hs, n, mp, mk))
        out_maps, out_masks = ([] if maps else None), ([] if masks else None)
        for i, inp in enumerate(inputs):
            _, h, w = inp.shape
            if maps:
                out_maps.append(np.ctypeslib.as_array(C.cast(mp[i], C.POINTER(C.c_float)), shape=(h * w,)).copy().reshape(h, w))
                lib.ocrs_b200_free(mp[i])
            if masks:
                out_masks.append(np.ctypeslib.as_array(C.cast(mk[i], C.POINTER(C.c_uint8)), shape=(h * w,)).copy().reshape(h, w))
                lib.ocrs_b200_free(mk[i])
        return out_maps, out_masks

    def prepare_recognition_inputs(self, inp: OcrInput, lines: Sequence[Sequence[RotatedRect]]) -> List[np.ndarray]:
        """`--text-line-images`: the [height, w'] recognition input of every line of a page, one crop launch."""
        flat = [w for line in lines for w in line]
        arr = _rects_to_c(flat)
        offs = (C.c_size_t * (len(lines) + 1))()
        k = 0
        for i, line in enumerate(lines):
            offs[i] = k
            k += len(line)
        offs[len(lines)] = k
        images = C.POINTER(C.c_float)()
        height = C.c_int()
        widths = C.POINTER(C.c_int)()
        offsets = C.POINTER(C.c_size_t)()
        check(lib.ocrs_b200_engine_prepare_recognition_inputs(self._h, inp._h, arr, offs, len(lines), C.byref(images),
                                                              C.byref(height), C.byref(widths), C.byref(offsets)))
        out = []
        for i in range(len(lines)):
            n = height.value * widths[i]
            a = np.ctypeslib.as_array(C.cast(C.addressof(images.contents) + 4 * offsets[i], C.POINTER(C.c_float)), shape=(n,))
            out.append(a.copy().reshape(height.value, widths[i]))
        lib.ocrs_b200_free(images)
        lib.ocrs_b200_free(widths)
        lib.ocrs_b200_free(offsets)
        return out

    # -- lib.rs:282
    def detection_threshold(self) -> float:
        return float(lib.ocrs_b200_engine_detection_threshold(self._h))

    # -- lib.rs:290
    def get_text(self, inp: OcrInput) -> str:
        s = C.c_char_p()
        check(lib.ocrs_b200_engine_get_text(self._h, inp._h, C.byref(s)))
        text = s.value.decode("utf-8")
        lib.ocrs_b200_free(s)
        return text

    # -- batched pipeline (configs 2-5)
    def ocr_batch(self, inputs: Sequence[OcrInput]) -> List[List[Optional[TextLine]]]:
        hs = (C.c_void_p * max(len(inputs), 1))(*[i._h for i in inputs])
        res = (C.POINTER(TextResultC) * max(len(inputs), 1))()
        check(lib.ocrs_b200_engine_ocr_batch(self._h, hs, len(inputs), res))
        out = []
        for i in range(len(inputs)):
            out.append(_text_result(res[i]))
            lib.ocrs_b200_text_result_free(res[i])
        return out

    def ocr_batch_text(self, inputs: Sequence[OcrInput]) -> List[str]:
        """`get_text` (lib.rs:290-300) for a batch of pages."""
        hs = (C.c_void_p * max(len(inputs), 1))(*[i._h for i in inputs])
        res = (C.c_void_p * max(len(inputs), 1))()
        check(lib.ocrs_b200_engine_ocr_batch_text(self._h, hs, len(inputs), res))
        out = []
        for i in range(len(inputs)):
            out.append(C.string_at(res[i]).decode("utf-8"))
            lib.ocrs_b200_free(res[i])
        return out

    def stats(self, reset: bool = False) -> dict:
        buf = (C.c_double * 8)()
        check(lib.ocrs_b200_engine_stats(self._h, buf, int(reset)))
        return {"det_flops": buf[0], "rec_flops": buf[1], "words": int(buf[2]), "lines": int(buf[3]),
                "timesteps": int(buf[4]), "rec_batches": int(buf[5])}


    # -- measurement hooks
    def set_profiling(self, on: bool) -> None:
        check(lib.ocrs_b200_engine_set_profiling(self._h, int(on)))

    def profile(self, reset: bool = True) -> dict:
        import json
        s = C.c_char_p()
        check(lib.ocrs_b200_engine_profile_json(self._h, C.byref(s), int(reset)))
        out = json.loads(s.value.decode())
        lib.ocrs_b200_free(s)
        return out

    def timer_start(self) -> None:
        check(lib.ocrs_b200_engine_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(lib.ocrs_b200_engine_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def transfer_bytes(self) -> Tuple[int, int]:
        buf = (C.c_int64 * 2)()
        check(lib.ocrs_b200_engine_transfer_bytes(self._h, buf))
        return int(buf[0]), int(buf[1])


class OcrPool:
    """Engine pool inside the library (include/ocrs_b200.h, "engine pool"): per device `in_flight` worker
    threads with one engine each; batches are submitted asynchronously and collected by ticket.

        pool = OcrPool(OcrEngineParams(detection_model=..., recognition_model=...), devices=[0], in_flight=2)
        t = pool.submit([ImageSource.from_tensor(page, DimOrder.Hwc) for page in pages])
        texts = pool.wait_text(t)
    """

    def __init__(self, params: OcrEngineParams, devices: Optional[Sequence[int]] = None, in_flight: int = 2,
                 pin_numa: bool = True, layout_threads: int = 4):
        p = PoolParamsC()
        keep = _fill_engine_params(p.engine, params)
        ids = None
        if devices is not None:
            ids = (C.c_int32 * max(len(devices), 1))(*[int(d) for d in devices])
            p.device_ids = C.cast(ids, C.POINTER(C.c_int32))
            p.n_devices = len(devices)
        p.in_flight = int(in_flight)
        p.pin_numa = 1 if pin_numa else -1
        p.layout_threads = int(layout_threads)
        self._h = C.c_void_p()
        check(lib.ocrs_b200_pool_create(C.byref(p), C.byref(self._h)))
        del keep, ids
        self._alphabet = params.alphabet
        self._pending = {}  # ticket -> (n_pages, objects kept alive)

    def __del__(self):
        if getattr(self, "_h", None):
            lib.ocrs_b200_pool_destroy(self._h)
            self._h = None

    @property
    def shape(self) -> Tuple[int, int]:
        nd, nf = C.c_int(), C.c_int()
        check(lib.ocrs_b200_pool_shape(self._h, C.byref(nd), C.byref(nf)))
        return nd.value, nf.value

    def describe(self) -> str:
        out = C.c_void_p()
        check(lib.ocrs_b200_pool_describe(self._h, C.byref(out)))
        return _take_string(out)

    def engine(self, dev_index: int = 0, slot: int = 0) -> OcrEngine:
        h = C.c_void_p()
        check(lib.ocrs_b200_pool_engine(self._h, dev_index, slot, C.byref(h)))
        return OcrEngine(None, _handle=h, _alphabet=self._alphabet)

    def submit(self, pages: Sequence[ImageSource]) -> int:
        """Host pages; the arrays are kept alive until the ticket has been waited for."""
        arr = (PageC * max(len(pages), 1))()
        for i, img in enumerate(pages):
            a = img.data
            if img.order == DimOrder.Hwc:
                h, w, c = a.shape
            else:
                c, h, w = a.shape
            arr[i] = PageC(a.ctypes.data, 0 if a.dtype == np.uint8 else 1, int(img.order), h, w, c, 0)
        t = C.c_uint64()
        check(lib.ocrs_b200_pool_submit(self._h, arr, len(pages), C.byref(t)))
        self._pending[t.value] = (len(pages), [p.data for p in pages])
        return t.value

    def submit_device(self, ptrs: Sequence[int], dtype: int, order: DimOrder, h: int, w: int, c: int) -> int:
        """Pages already resident in the memory of one of the pool's GPUs (device pointers)."""
        arr = (PageC * max(len(ptrs), 1))()
        for i, ptr in enumerate(ptrs):
            arr[i] = PageC(int(ptr), int(dtype), int(order), h, w, c, 1)
        t = C.c_uint64()
        check(lib.ocrs_b200_pool_submit(self._h, arr, len(ptrs), C.byref(t)))
        self._pending[t.value] = (len(ptrs), None)
        return t.value

    def done(self, ticket: int) -> bool:
        d = C.c_int()
        check(lib.ocrs_b200_pool_done(self._h, C.c_uint64(ticket), C.byref(d)))
        return bool(d.value)

    def wait_text(self, ticket: int) -> List[str]:
        n, _keep = self._pending.pop(ticket)
        res = (C.c_void_p * max(n, 1))()
        check(lib.ocrs_b200_pool_wait_text(self._h, C.c_uint64(ticket), res, n))
        out = []
        for i in range(n):
            out.append(C.string_at(res[i]).decode("utf-8"))
            lib.ocrs_b200_free(res[i])
        return out

    def wait(self, ticket: int) -> List[List[Optional[TextLine]]]:
        n, _keep = self._pending.pop(ticket)
        res = (C.POINTER(TextResultC) * max(n, 1))()
        check(lib.ocrs_b200_pool_wait(self._h, C.c_uint64(ticket), res, n))
        out = []
        for i in range(n):
            out.append(_text_result(res[i]))
            lib.ocrs_b200_text_result_free(res[i])
        return out


def kernel_launch_count() -> int:
    return int(lib.ocrs_b200_kernel_launch_count())


def inspect_model(model) -> dict:
    """Parses an `.onnx` file (path or bytes) on the host -- no GPU involved -- and returns its inputs,
    outputs, operator histogram and the operators the executor does not implement.  Raises OcrsError
    (`OCRS_B200_ERR_MODEL_LOAD`) for malformed files and `.rten` containers."""
    import json
    data = _model_bytes(model)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    out = C.c_void_p()
    check(lib.ocrs_b200_model_inspect(C.cast(buf, C.c_void_p), len(data), C.byref(out)))
    return json.loads(_take_string(out))


def find_text_lines(words: Sequence[RotatedRect]) -> List[List[RotatedRect]]:
    """layout_analysis.rs:158 through the C ABI (host code; needs no GPU)."""
    arr = _rects_to_c(words)
    out_words = C.POINTER(RotatedRectC)()
    offs = C.POINTER(C.c_size_t)()
    n_lines = C.c_size_t()
    check(lib.ocrs_b200_find_text_lines(arr, len(words), C.byref(out_words), C.byref(offs), C.byref(n_lines)))
    lines = [[RotatedRect(out_words[k].cx, out_words[k].cy, out_words[k].ux, out_words[k].uy, out_words[k].w,
                          out_words[k].h) for k in range(offs[i], offs[i + 1])] for i in range(n_lines.value)]
    lib.ocrs_b200_free(out_words)
    lib.ocrs_b200_free(offs)
    return lines


def device_count() -> int:
    return int(lib.ocrs_b200_device_count())
