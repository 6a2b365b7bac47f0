"""Engine pool (ocrs_b200_pool_*): asynchronous batches through worker engines give exactly what the
single-engine calls give; tickets, errors and device-resident pages."""
import numpy as np
import pytest

import ocrs_b200 as ob
from tools.models import ensure_models
from tools.synth import make_page

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    return ensure_models()


@pytest.fixture(scope="module")
def pages():
    return [np.ascontiguousarray(make_page(900 + i, 384, 512, n_rows=6)[0]) for i in range(5)]


def _params(models, **kw):
    return ob.OcrEngineParams(detection_model=models[0], recognition_model=models[1], **kw)


def test_pool_matches_engine(models, pages):
    eng = ob.OcrEngine(_params(models))
    want = eng.ocr_batch([eng.prepare_input(ob.ImageSource.from_tensor(p, ob.DimOrder.Hwc)) for p in pages])
    want_text = ["\n".join(str(l) for l in pg if l is not None) for pg in want]
    pool = ob.OcrPool(_params(models), devices=[0], in_flight=2)
    assert pool.shape == (1, 2)
    src = [ob.ImageSource.from_tensor(p, ob.DimOrder.Hwc) for p in pages]
    tickets = [pool.submit(src) for _ in range(4)] + [pool.submit(src[:2]), pool.submit([])]
    got = [pool.wait(t) for t in tickets[:2]] + [pool.wait_text(t) for t in tickets[2:]]
    for g in got[:2]:
        assert [[(str(l), [c.rect.tlbr() for c in l.chars]) if l is not None else None for l in pg] for pg in g] == \
               [[(str(l), [c.rect.tlbr() for c in l.chars]) if l is not None else None for l in pg] for pg in want]
    assert got[2] == want_text and got[3] == want_text
    assert got[4] == want_text[:2] and got[5] == []
    assert "worker 0.0" in pool.describe() and "worker 0.1" in pool.describe()
    # worker engines are visible for the statistics hooks
    assert pool.engine(0, 1).stats()["lines"] + pool.engine(0, 0).stats()["lines"] > 0


def test_pool_device_resident_pages(models, pages):
    torch = pytest.importorskip("torch")
    pool = ob.OcrPool(_params(models), devices=[0], in_flight=1)
    host = pool.wait_text(pool.submit([ob.ImageSource.from_tensor(p, ob.DimOrder.Hwc) for p in pages]))
    dev = [torch.from_numpy(p).cuda(0) for p in pages]
    t = pool.submit_device([d.data_ptr() for d in dev], 0, ob.DimOrder.Hwc, pages[0].shape[0], pages[0].shape[1], 3)
    assert pool.wait_text(t) == host


def test_pool_errors(models, pages):
    pool = ob.OcrPool(_params(models), devices=[0], in_flight=1)
    bad = ob.ImageSource(np.zeros((8, 8, 2), np.uint8), ob.DimOrder.Hwc)  # bypasses from_tensor's check
    with pytest.raises(ob.OcrsError) as ei:
        pool.submit([bad])
    assert ei.value.code == -2  # UnsupportedChannelCount (preprocess.rs:41), reported at submit
    t = pool.submit([ob.ImageSource.from_tensor(pages[0], ob.DimOrder.Hwc)])
    assert len(pool.wait_text(t)) == 1
    pool._pending[t] = (1, None)
    with pytest.raises(ob.OcrsError, match="ticket"):
        pool.wait_text(t)  # a ticket is consumed by its wait
    with pytest.raises(ob.OcrsError):
        ob.OcrPool(_params(models), devices=[0, 0])
    with pytest.raises(ob.OcrsError):
        ob.OcrPool(_params(models), devices=[99])
    # a failing batch (recognition model missing) surfaces at wait, and the pool keeps working
    p2 = ob.OcrPool(ob.OcrEngineParams(detection_model=models[0]), devices=[0], in_flight=1)
    t = p2.submit([ob.ImageSource.from_tensor(pages[0], ob.DimOrder.Hwc)])
    with pytest.raises(ob.OcrsError) as ei:
        p2.wait_text(t)
    assert ei.value.code == -4  # "Recognition model not loaded" (lib.rs:254)
