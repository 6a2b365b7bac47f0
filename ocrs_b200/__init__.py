"""ocrs_b200: B200-native OCR hot path behind the public surface of robertknight/ocrs.

The package holds the CUDA/C++ library sources (`csrc/`), its C ABI binding (`_lib.py`) and a
host-side mirror of the reference API (`engine.py`).  Importing it loads
`libocrs_b200.so`; there is no CPU fallback."""
from .engine import (  # noqa: F401
    DEFAULT_ALPHABET, DecodeMethod, DimOrder, ImageSource, ImageSourceError, Model, OcrEngine, OcrEngineParams,
    OcrInput, OcrPool, Rect, RotatedRect, TextChar, TextItem, TextLine, TextWord, device_count, find_text_lines,
    format_json_output, format_text_output, inspect_model, kernel_launch_count,
)
from ._lib import OcrsError, LIB_PATH  # noqa: F401
