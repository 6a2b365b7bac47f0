"""Shared helpers for the `-m gpu` parity tests (everything goes through the C ABI)."""
import os

import numpy as np

import ocrs_b200 as ob
from oracle.geometry import RotatedRect as ORect
from tools import models as M

_TMP = {}


def model_paths():
    """Seeded synthetic (or committed trained) detection / recognition models."""
    return M.ensure_models()


def fake_paths(tmpdir, det_hw=(200, 100)):
    key = (str(tmpdir), det_hw)
    if key not in _TMP:
        d = os.path.join(str(tmpdir), f"fake_det_{det_hw[0]}x{det_hw[1]}.onnx")
        r = os.path.join(str(tmpdir), "fake_rec.onnx")
        M.export_fake_detection(d, det_hw)
        M.export_fake_recognition(r)
        _TMP[key] = (d, r)
    return _TMP[key]


def to_oracle_rects(rects):
    return [ORect.from_raw(*r.raw()) for r in rects]


def raw32(rects):
    return [tuple(np.float32(v) for v in r.raw()) for r in rects]


def lines_raw(lines):
    return [raw32(l) for l in lines]


def text_of(lines):
    return [None if t is None else str(t) for t in lines]


def oracle_text_of(lines):
    from oracle.engine import line_text
    return [None if t is None else line_text(t) for t in lines]


def char_boxes(lines):
    return [None if t is None else [c.rect.tlbr() for c in t.chars] for t in lines]


def oracle_char_boxes(lines):
    return [None if t is None else [c.rect.tlbr() for c in t] for t in lines]
