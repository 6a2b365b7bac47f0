"""The oracle pipeline with the committed model fixtures reads a synthetic page (CPU only).  This is
what makes the GPU-vs-oracle string comparisons meaningful: the fixtures are real (small) OCR models
trained on the page generator (tools/train_models.py), not random weights."""
import pytest

from oracle.engine import OcrEngine, OcrEngineParams
from oracle.onnx_eval import OnnxModel
from tools.models import ensure_models
from tools.synth import make_page


@pytest.mark.slow
def test_oracle_reads_a_page():
    det, rec = ensure_models()
    eng = OcrEngine(OcrEngineParams(detection_model=OnnxModel(det), recognition_model=OnnxModel(rec)))
    page, texts = make_page(7)
    got = eng.get_text(eng.prepare_input(page, "hwc")).split("\n")
    exact = sum(1 for g in got if g in texts)
    assert len(got) == len(texts)
    assert exact >= 0.9 * len(texts), (exact, len(texts))
