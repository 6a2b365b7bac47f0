"""Text items and CLI output formats on the host library (no GPU needed), against the reference's
own known-answer tests: ocrs/src/text_items.rs:120-157 and ocrs-cli/src/output.rs:218-249 with
ocrs-cli/test-data/format-json-expected.json (re-typed below: the fixture is 60 lines)."""
import json

import pytest

import ocrs_b200 as ob
from oracle.engine import item_rotated_rect
from oracle.geometry import Rect as ORect


def gen_text_chars(text, width):
    """output.rs:199-207 / text_items.rs:124-132: Rect::from_tlhw(0, i*width, 25, width)."""
    return [ob.TextChar(ch, ob.Rect(0, i * width, 25, (i + 1) * width)) for i, ch in enumerate(text)]


EXPECTED_JSON = {  # ocrs-cli/test-data/format-json-expected.json
    "url": "image.jpeg", "image_width": 256, "image_height": 256,
    "paragraphs": [{"lines": [
        {"text": "line one", "vertices": [[80, 25], [0, 25], [0, 0], [80, 0]], "words": [
            {"text": "line", "vertices": [[40, 25], [0, 25], [0, 0], [40, 0]]},
            {"text": "one", "vertices": [[80, 25], [50, 25], [50, 0], [80, 0]]}]},
        {"text": "line two", "vertices": [[80, 25], [0, 25], [0, 0], [80, 0]], "words": [
            {"text": "line", "vertices": [[40, 25], [0, 25], [0, 0], [40, 0]]},
            {"text": "two", "vertices": [[80, 25], [50, 25], [50, 0], [80, 0]]}]},
    ]}],
}


def _lines():
    return [ob.TextLine(gen_text_chars("line one", 10)), None, ob.TextLine(gen_text_chars("line two", 10))]


def test_format_json_output():
    """output.rs:218-236"""
    text = ob.format_json_output("image.jpeg", [256, 256], _lines())
    assert json.loads(text) == EXPECTED_JSON
    # serde_json::to_string_pretty layout: sorted keys, two-space indentation, one element per line
    assert text.startswith('{\n  "image_height": 256,\n  "image_width": 256,\n  "paragraphs": [\n    {\n      "lines": [\n')
    assert text.endswith('  ],\n  "url": "image.jpeg"\n}')
    assert text == json.dumps(EXPECTED_JSON, indent=2, sort_keys=True)


def test_format_text_output():
    """output.rs:238-249"""
    assert ob.format_text_output(_lines()).split("\n") == ["line one", "line two"]
    assert ob.format_text_output([None]) == ""


def test_item_display_and_words():
    """text_items.rs:134-147"""
    line = ob.TextLine(gen_text_chars("foo bar baz", 10))
    assert str(line) == "foo bar baz"
    assert [str(w) for w in line.words()] == ["foo", "bar", "baz"]
    assert [str(w) for w in ob.TextLine(gen_text_chars("  a  b ", 4)).words()] == ["a", "b"]


def test_item_rotated_rect():
    """text_items.rs:138-157, same word and assertions."""
    word = ob.TextWord(gen_text_chars("foo", 10))
    assert word.bounding_rect().tlbr() == (0, 0, 25, 30)
    rr = word.rotated_rect()
    assert (rr.ux, rr.uy) == (0.0, -1.0)                                    # up_axis == Vec2::from_yx(-1., 0.)
    assert rr.rounded_vertices() == [[30, 25], [0, 25], [0, 0], [30, 0]]    # corners (y,x): (25,30),(25,0),(0,0),(0,30)


def test_item_rects():
    """A longer axis-aligned item."""
    line = ob.TextLine(gen_text_chars("foo bar baz", 10))
    assert line.bounding_rect().tlbr() == (0, 0, 25, 110)
    rr = line.rotated_rect()
    assert (rr.cx, rr.cy, rr.ux, rr.uy, rr.w, rr.h) == (55.0, 12.5, 0.0, -1.0, 110.0, 25.0)
    assert rr.rounded_vertices() == [[110, 25], [0, 25], [0, 0], [110, 0]]


@pytest.mark.parametrize("seed", range(5))
def test_rotated_rect_matches_oracle_on_ragged_boxes(seed):
    import numpy as np
    from oracle.recognition import TextChar as OChar
    rng = np.random.default_rng(seed)
    x, chars, ochars = 5, [], []
    for i in range(int(rng.integers(1, 12))):
        w, t, h = int(rng.integers(4, 15)), int(rng.integers(0, 9)), int(rng.integers(8, 30))
        chars.append(ob.TextChar("x", ob.Rect(t, x, t + h, x + w)))
        ochars.append(OChar("x", ORect(t, x, t + h, x + w)))
        x += w + int(rng.integers(0, 4))
    got, exp = ob.TextLine(chars).rotated_rect(), item_rotated_rect(ochars)
    assert tuple(np.float32(v) for v in got.raw()) == tuple(np.float32(v) for v in exp.raw())


def test_json_escaping_and_unicode():
    line = ob.TextLine([ob.TextChar(c, ob.Rect(0, 10 * i, 10, 10 * i + 10)) for i, c in enumerate('a"\\é€\t')])
    doc = json.loads(ob.format_json_output('dir/"x".png', [10, 60], [line]))
    assert doc["paragraphs"][0]["lines"][0]["text"] == 'a"\\é€\t'
    assert doc["url"] == 'dir/"x".png'
    empty = json.loads(ob.format_json_output("p", [1, 1], [None, None]))
    assert empty["paragraphs"] == [{"lines": []}]
