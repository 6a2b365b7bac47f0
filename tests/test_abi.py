"""C-ABI checks that need no GPU: the library loads, exports every symbol include/ocrs_b200.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import ocrs_b200 as ob
from ocrs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ocrs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ocrs_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 25
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/ocrs_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) <= set(names)


def test_version_and_error_channel():
    assert b"sm_100a" in _lib.lib.ocrs_b200_version()
    rc = _lib.lib.ocrs_b200_model_input_shape(None, None, None)
    assert rc == _lib.ERR_INVALID_ARG
    assert b"null" in _lib.lib.ocrs_b200_last_error()


def test_image_source_validation_matches_reference():
    # preprocess.rs:274-360 through the Python mirror (host logic)
    ob.ImageSource.from_bytes(np.arange(100, dtype=np.uint8), (10, 10))
    with pytest.raises(ob.ImageSourceError, match="multiple"):
        ob.ImageSource.from_bytes(np.arange(50, dtype=np.uint8), (10, 10))
    with pytest.raises(ob.ImageSourceError, match="channel count"):
        ob.ImageSource.from_bytes(np.arange(128, dtype=np.uint8), (8, 8))
    with pytest.raises(ob.ImageSourceError, match="channel count"):
        ob.ImageSource.from_bytes(np.zeros(0, dtype=np.uint8), (0, 10))
    ob.ImageSource.from_tensor(np.zeros((1, 5, 5), np.uint8), ob.DimOrder.Chw)
    with pytest.raises(ob.ImageSourceError):
        ob.ImageSource.from_tensor(np.zeros((1, 5, 5), np.uint8), ob.DimOrder.Hwc)
    with pytest.raises(ob.ImageSourceError):
        ob.ImageSource.from_tensor(np.zeros((0, 5, 5), np.uint8), ob.DimOrder.Chw)


@pytest.mark.skipif(ob.device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(ob.OcrsError) as ei:
        ob.OcrEngine(ob.OcrEngineParams())
    assert ei.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(ob.OcrsError) as ei:
        ob.Model(b"\x08\x08")
    assert ei.value.code == _lib.ERR_NO_DEVICE


def test_pool_create_without_device_fails_cleanly():
    """The pool is a compute entry point: no GPU -> OCRS_B200_ERR_NO_DEVICE, never a fallback."""
    import ocrs_b200 as ob
    if ob.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ob.OcrsError) as ei:
        ob.OcrPool(ob.OcrEngineParams(), devices=[0])
    assert ei.value.code == -9
