"""Oracle `OcrEngine`: the reference's public surface on the CPU (test infrastructure).

Restates ocrs/src/lib.rs (OcrEngine, OcrEngineParams), detection.rs (TextDetector) and the
driver half of recognition.rs (TextRecognizer::recognize_text_lines).  A `model` is anything
with `input_shape() -> [int | str, ...]` and `run(np.ndarray) -> np.ndarray`, mirroring the
`Model` trait (model.rs:6-17) so the reference's fake models (lib.rs:339-422) plug in.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

import numpy as np

from . import BLACK_VALUE
from .contours import find_contours_external
from .geometry import F, PointF, Rect, RotatedRect, bounding_rect_of_rotated, min_area_rect, polygon_bounding_rect, simplify_polygon
from .imageops import pad_bottom_right, prepare_image, resize_bilinear, threshold_mask
from .layout import find_text_lines
from .recognition import (
    DecodeStep, LineRecResult, TextChar, TextRecLine, ctc_decode_beam, ctc_decode_greedy, filter_excluded_char_labels,
    line_polygon, next_multiple_of, prepare_text_line, prepare_text_line_batch, resized_line_width,
    text_lines_from_recognition_results,
)

# lib.rs:34
DEFAULT_ALPHABET = " 0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~EABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"


class ModelRunError(RuntimeError):
    """errors.rs:6-12"""


class _Stage:
    """Optional per-stage wall-clock accounting of the CPU path (bench.py's `cpu_baseline.stage_s`):
    `timers` is a dict name -> seconds, or None."""

    def __init__(self, timers, name):
        self.timers, self.name = timers, name

    def __enter__(self):
        if self.timers is not None:
            import time
            self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        if self.timers is not None:
            import time
            self.timers[self.name] = self.timers.get(self.name, 0.0) + time.perf_counter() - self.t0
        return False


def find_connected_component_rects(mask: np.ndarray, expand_dist: float, min_area: float) -> List[RotatedRect]:
    """detection.rs:41-62"""
    out = []
    for poly in find_contours_external(mask):
        pts = [PointF(F(x), F(y)) for x, y in poly.tolist()]
        simplified = simplify_polygon(pts, 2.0)
        rect = min_area_rect(simplified)
        if rect is None:
            continue
        rect.resize(rect.width() + F(2.0) * F(expand_dist), rect.height() + F(2.0) * F(expand_dist))
        if rect.area() >= F(min_area):
            out.append(rect)
    return out


class TextDetector:
    """detection.rs:66-201"""

    def __init__(self, model, min_area: float = 100.0, text_threshold: float = 0.2):
        self.model = model
        self.min_area = min_area  # detection.rs:31
        self.text_threshold = text_threshold  # detection.rs:34
        self.input_shape = model.input_shape()

    def threshold(self) -> float:
        return self.text_threshold

    def detect_text_pixels(self, image: np.ndarray, timers=None) -> np.ndarray:
        """detection.rs:131-200.  image f32 [1,H,W] -> prob map f32 [H,W]."""
        _, img_h, img_w = image.shape
        shp = self.input_shape
        if len(shp) != 4 or not isinstance(shp[2], int) or not isinstance(shp[3], int):
            raise ValueError("failed to get model dims")
        in_h, in_w = shp[2], shp[3]
        x = image[None]
        pad_bottom = max(in_h - img_h, 0)
        pad_right = max(in_w - img_w, 0)
        if pad_bottom > 0 or pad_right > 0:
            x = pad_bottom_right(x, pad_bottom, pad_right, BLACK_VALUE)
        if x.shape[2] != in_h or x.shape[3] != in_w:
            x = resize_bilinear(x, in_h, in_w)
        with _Stage(timers, "detection_net"):
            mask = self.model.run(np.ascontiguousarray(x, dtype=np.float32))
        mask = mask[:, :, : in_h - pad_bottom, : in_w - pad_right]
        mask = resize_bilinear(mask, img_h, img_w)
        return mask.reshape(img_h, img_w)

    def detect_words(self, image: np.ndarray, timers=None) -> List[RotatedRect]:
        """detection.rs:104-122"""
        with _Stage(timers, "detect_pad_resize_net"):
            prob = self.detect_text_pixels(image, timers)
        with _Stage(timers, "threshold_contours_rects"):
            binary = threshold_mask(prob, self.text_threshold)
            return find_connected_component_rects(binary, 3.0, self.min_area)


class TextRecognizer:
    """recognition.rs:315-562"""

    def __init__(self, model):
        self.model = model
        self.input_shape = model.input_shape()

    def input_height(self) -> int:
        d = self.input_shape[2]
        return d if isinstance(d, int) else 50  # recognition.rs:332-337

    def run(self, inp: np.ndarray) -> np.ndarray:
        """recognition.rs:341-360: [B,1,H,W] -> [B, T, C]."""
        try:
            out = self.model.run(inp)
        except Exception as e:  # noqa: BLE001
            raise ModelRunError(f"model run failed: {e}") from e
        if out.ndim != 3:
            raise ModelRunError(f"expected recognition output to have 3 dims but it has {out.ndim}")
        return np.transpose(out, (1, 0, 2))

    def line_plan(self, lines: Sequence[Sequence[RotatedRect]]):
        """recognition.rs:429-459: bucket lines by padded width, chunks of <= 20.
        Groups are emitted in ascending width (the reference iterates a HashMap; order is
        irrelevant because results are re-sorted by line index, recognition.rs:535)."""
        h = self.input_height()
        groups = {}
        for idx, words in enumerate(lines):
            br = bounding_rect_of_rotated(words)
            assert br is not None, "line has no words"
            line_rect = br.integral_bounding_rect()
            rw = resized_line_width(line_rect.width(), line_rect.height(), h)
            gw = next_multiple_of(rw, 50)
            groups.setdefault(gw, []).append(TextRecLine(idx, line_polygon(words), rw))
        out = []
        for gw in sorted(groups):
            ls = groups[gw]
            for i in range(0, len(ls), 20):
                out.append((gw, ls[i:i + 20]))
        return out

    def prepare_input(self, image: np.ndarray, line: Sequence[RotatedRect]) -> np.ndarray:
        """recognition.rs:366-393"""
        br = bounding_rect_of_rotated(line)
        assert br is not None, "line has no words"
        line_rect = br.integral_bounding_rect()
        h = self.input_height()
        rw = resized_line_width(line_rect.width(), line_rect.height(), h)
        return prepare_text_line(image, line_polygon(line), rw, h)

    def recognize_text_lines(self, image: np.ndarray, lines, alphabet: str, excluded_char_labels=None,
                             decode_method: str = "greedy", beam_width: int = 100, collect=None, timers=None):
        """recognition.rs:404-540"""
        h = self.input_height()
        alphabet_len = len(alphabet)
        results: List[LineRecResult] = []
        with _Stage(timers, "line_polygons"):
            plan = self.line_plan(lines)
        for gw, group in plan:
            with _Stage(timers, "line_crops"):
                rec_input = prepare_text_line_batch(image, group, h, gw)
            with _Stage(timers, "recognition_net"):
                rec_output = self.run(rec_input)
            if alphabet_len + 1 != rec_output.shape[2]:
                raise ModelRunError(
                    f"output column count ({rec_output.shape[2]}) does not match alphabet size ({alphabet_len + 1})")
            ctc_input_len = rec_output.shape[1]
            with _Stage(timers, "ctc_decode"):
                for gi, line in enumerate(group):
                    seq = filter_excluded_char_labels(rec_output[gi], excluded_char_labels)
                    if decode_method == "greedy":
                        steps, _ = ctc_decode_greedy(seq)
                    else:
                        steps, _ = ctc_decode_beam(seq, beam_width)
                    results.append(LineRecResult(line, gw, ctc_input_len, steps))
                    if collect is not None:
                        collect.append((line.index, gw, rec_input[gi], rec_output[gi]))
        results.sort(key=lambda r: r.line.index)
        with _Stage(timers, "text_assembly"):
            return text_lines_from_recognition_results(results, alphabet)


@dataclass
class OcrEngineParams:
    """lib.rs:37-71"""
    detection_model: Any = None
    recognition_model: Any = None
    debug: bool = False
    decode_method: str = "greedy"  # or "beam"
    beam_width: int = 100
    alphabet: Optional[str] = None
    allowed_chars: Optional[str] = None


class OcrEngine:
    """lib.rs:111-301"""

    def __init__(self, params: OcrEngineParams):
        self.detector = TextDetector(params.detection_model) if params.detection_model is not None else None
        self.recognizer = TextRecognizer(params.recognition_model) if params.recognition_model is not None else None
        self.alphabet = params.alphabet if params.alphabet is not None else DEFAULT_ALPHABET
        self.excluded_char_labels = None
        if params.allowed_chars is not None:  # lib.rs:153-170
            self.excluded_char_labels = [i + 1 for i, ch in enumerate(self.alphabet) if ch not in params.allowed_chars]
        self.decode_method = params.decode_method
        self.beam_width = params.beam_width

    def prepare_input(self, pixels: np.ndarray, order: str = "hwc") -> np.ndarray:
        """lib.rs:183-187 -> f32 [1,H,W]"""
        return prepare_image(pixels, order)

    def detect_words(self, image: np.ndarray, timers=None) -> List[RotatedRect]:
        if self.detector is None:
            raise RuntimeError("Detection model not loaded")  # lib.rs:197
        return self.detector.detect_words(image, timers)

    def detect_text_pixels(self, image: np.ndarray) -> np.ndarray:
        if self.detector is None:
            raise RuntimeError("Detection model not loaded")
        return self.detector.detect_text_pixels(image)

    def find_text_lines(self, image, words: Sequence[RotatedRect]) -> List[List[RotatedRect]]:
        return find_text_lines(words)

    def recognize_text(self, image: np.ndarray, lines, collect=None, timers=None):
        if self.recognizer is None:
            raise RuntimeError("Recognition model not loaded")  # lib.rs:254
        return self.recognizer.recognize_text_lines(
            image, lines, self.alphabet, self.excluded_char_labels, self.decode_method, self.beam_width, collect, timers)

    def prepare_recognition_input(self, image: np.ndarray, line) -> np.ndarray:
        if self.recognizer is None:
            raise RuntimeError("Recognition model not loaded")
        return self.recognizer.prepare_input(image, line)

    def detection_threshold(self) -> float:
        return self.detector.threshold() if self.detector is not None else 0.2

    def get_text(self, image: np.ndarray, timers=None) -> str:
        """lib.rs:290-300.  `timers` (dict or None) accumulates seconds per stage."""
        words = self.detect_words(image, timers)
        with _Stage(timers, "find_text_lines"):
            lines = self.find_text_lines(image, words)
        texts = self.recognize_text(image, lines, timers=timers)
        return "\n".join(line_text(t) for t in texts if t is not None)


def line_text(chars: Sequence[TextChar]) -> str:
    """text_items.rs:33-44 (Display)"""
    return "".join(c.char for c in chars)


def line_words(chars: Sequence[TextChar]) -> List[List[TextChar]]:
    """`TextLine::words` (text_items.rs:76-82): split on ' ' and drop empty runs."""
    out, cur = [], []
    for c in chars:
        if c.char == " ":
            if cur:
                out.append(cur)
            cur = []
        else:
            cur.append(c)
    if cur:
        out.append(cur)
    return out


def item_bounding_rect(chars: Sequence[TextChar]) -> Rect:
    """text_items.rs:13-15"""
    out = None
    for c in chars:
        out = c.rect if out is None else out.union(c.rect)
    assert out is not None, "expected valid rect"
    return out


def item_rotated_rect(chars: Sequence[TextChar]) -> RotatedRect:
    """text_items.rs:18-30"""
    from .geometry import Vec2
    pts = [p.to_f32() for c in chars for p in c.rect.corners()]
    rect = min_area_rect(pts)
    assert rect is not None, "expected valid rect"
    return rect.orient_towards(Vec2.from_yx(-1.0, 0.0))
