"""ctypes binding of libocrs_b200.so (C ABI in include/ocrs_b200.h).

The library is built in-tree by `make` / `__graft_entry__.build()`.  There is deliberately no
fallback: if the shared object is missing, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libocrs_b200.so")


class OcrsError(RuntimeError):
    """A negative status from the C ABI.  `.code` holds the ocrs_b200_status value."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


# status codes (include/ocrs_b200.h)
OK = 0
ERR_INVALID_ARG = -1
ERR_UNSUPPORTED_CHANNEL_COUNT = -2
ERR_INVALID_DATA_LENGTH = -3
ERR_MODEL_NOT_LOADED = -4
ERR_MODEL_LOAD = -5
ERR_RUN_FAILED = -6
ERR_WRONG_OUTPUT = -7
ERR_CUDA = -8
ERR_NO_DEVICE = -9
ERR_INTERNAL = -10


class RotatedRectC(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("ux", C.c_float), ("uy", C.c_float),
                ("w", C.c_float), ("h", C.c_float)]


class RectC(C.Structure):
    _fields_ = [("top", C.c_int32), ("left", C.c_int32), ("bottom", C.c_int32), ("right", C.c_int32)]


class TextResultC(C.Structure):
    _fields_ = [("n_lines", C.c_int32), ("line_present", C.POINTER(C.c_uint8)),
                ("char_offsets", C.POINTER(C.c_int64)), ("chars", C.POINTER(C.c_uint32)),
                ("char_rects", C.POINTER(RectC))]


class EngineParamsC(C.Structure):
    _fields_ = [("detection_model", C.c_void_p), ("detection_model_len", C.c_size_t),
                ("recognition_model", C.c_void_p), ("recognition_model_len", C.c_size_t),
                ("debug", C.c_int32), ("decode_method", C.c_int32), ("beam_width", C.c_uint32),
                ("alphabet_utf8", C.c_char_p), ("allowed_chars_utf8", C.c_char_p), ("device", C.c_int32)]


class PageC(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("dtype", C.c_int32), ("order", C.c_int32), ("height", C.c_int32),
                ("width", C.c_int32), ("channels", C.c_int32), ("on_device", C.c_int32)]


class PoolParamsC(C.Structure):
    _fields_ = [("engine", EngineParamsC), ("device_ids", C.POINTER(C.c_int32)), ("n_devices", C.c_int32),
                ("in_flight", C.c_int32), ("pin_numa", C.c_int32), ("layout_threads", C.c_int32)]


# every symbol include/ocrs_b200.h declares: (restype, argtypes)
P = C.POINTER
SIGNATURES = {
    "ocrs_b200_last_error": (C.c_char_p, []),
    "ocrs_b200_device_count": (C.c_int, []),
    "ocrs_b200_free": (None, [C.c_void_p]),
    "ocrs_b200_version": (C.c_char_p, []),
    "ocrs_b200_model_load": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, P(C.c_void_p)]),
    "ocrs_b200_model_load_file": (C.c_int, [C.c_char_p, C.c_int, P(C.c_void_p)]),
    "ocrs_b200_model_input_shape": (C.c_int, [C.c_void_p, P(C.c_int64), P(C.c_int)]),
    "ocrs_b200_model_run": (C.c_int, [C.c_void_p, C.c_void_p, P(C.c_int64), C.c_int, P(P(C.c_float)),
                                      P(C.c_int64), P(C.c_int)]),
    "ocrs_b200_model_inspect": (C.c_int, [C.c_void_p, C.c_size_t, P(C.c_void_p)]),
    "ocrs_b200_model_last_flops": (C.c_double, [C.c_void_p]),
    "ocrs_b200_model_destroy": (None, [C.c_void_p]),
    "ocrs_b200_engine_create": (C.c_int, [P(EngineParamsC), P(C.c_void_p)]),
    "ocrs_b200_engine_destroy": (None, [C.c_void_p]),
    "ocrs_b200_engine_prepare_input_bytes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                                       P(C.c_void_p)]),
    "ocrs_b200_engine_prepare_input": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 P(C.c_void_p)]),
    "ocrs_b200_engine_prepare_input_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                        C.c_int, P(C.c_void_p)]),
    "ocrs_b200_input_shape": (C.c_int, [C.c_void_p, P(C.c_int), P(C.c_int)]),
    "ocrs_b200_input_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ocrs_b200_input_destroy": (None, [C.c_void_p]),
    "ocrs_b200_engine_detect_text_pixels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ocrs_b200_engine_detect_words": (C.c_int, [C.c_void_p, C.c_void_p, P(P(RotatedRectC)), P(C.c_size_t)]),
    "ocrs_b200_engine_find_text_lines": (C.c_int, [C.c_void_p, C.c_void_p, P(RotatedRectC), C.c_size_t,
                                                   P(P(RotatedRectC)), P(P(C.c_size_t)), P(C.c_size_t)]),
    "ocrs_b200_find_text_lines": (C.c_int, [P(RotatedRectC), C.c_size_t, P(P(RotatedRectC)), P(P(C.c_size_t)),
                                            P(C.c_size_t)]),
    "ocrs_b200_engine_recognize_text": (C.c_int, [C.c_void_p, C.c_void_p, P(RotatedRectC), P(C.c_size_t), C.c_size_t,
                                                  P(P(TextResultC))]),
    "ocrs_b200_text_result_free": (None, [P(TextResultC)]),
    "ocrs_b200_text_item_rotated_rect": (C.c_int, [P(RectC), C.c_size_t, P(RotatedRectC)]),
    "ocrs_b200_rotated_rect_vertices": (C.c_int, [P(RotatedRectC), P(C.c_int32)]),
    "ocrs_b200_format_text_output": (C.c_int, [P(TextResultC), P(C.c_void_p)]),
    "ocrs_b200_format_json_output": (C.c_int, [P(TextResultC), C.c_char_p, C.c_int, C.c_int, P(C.c_void_p)]),
    "ocrs_b200_engine_prepare_recognition_input": (C.c_int, [C.c_void_p, C.c_void_p, P(RotatedRectC), C.c_size_t,
                                                             P(P(C.c_float)), P(C.c_int), P(C.c_int)]),
    "ocrs_b200_engine_detection_threshold": (C.c_float, [C.c_void_p]),
    "ocrs_b200_engine_get_text": (C.c_int, [C.c_void_p, C.c_void_p, P(C.c_char_p)]),
    "ocrs_b200_engine_ocr_batch": (C.c_int, [C.c_void_p, P(C.c_void_p), C.c_size_t, P(P(TextResultC))]),
    "ocrs_b200_engine_ocr_batch_text": (C.c_int, [C.c_void_p, P(C.c_void_p), C.c_size_t, P(C.c_void_p)]),
    "ocrs_b200_engine_detect_words_batch": (C.c_int, [C.c_void_p, P(C.c_void_p), C.c_size_t, P(P(RotatedRectC)),
                                                      P(P(C.c_size_t))]),
    "ocrs_b200_engine_stats": (C.c_int, [C.c_void_p, P(C.c_double), C.c_int]),
    "ocrs_b200_engine_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "ocrs_b200_engine_profile_json": (C.c_int, [C.c_void_p, P(C.c_char_p), C.c_int]),
    "ocrs_b200_engine_timer_start": (C.c_int, [C.c_void_p]),
    "ocrs_b200_engine_timer_stop": (C.c_int, [C.c_void_p, P(C.c_float)]),
    "ocrs_b200_engine_transfer_bytes": (C.c_int, [C.c_void_p, P(C.c_int64)]),
    "ocrs_b200_kernel_launch_count": (C.c_int64, []),
    "ocrs_b200_engine_detect_text_pixels_batch": (C.c_int, [C.c_void_p, P(C.c_void_p), C.c_size_t, P(C.c_void_p),
                                                            P(C.c_void_p)]),
    "ocrs_b200_engine_prepare_recognition_inputs": (C.c_int, [C.c_void_p, C.c_void_p, P(RotatedRectC), P(C.c_size_t),
                                                              C.c_size_t, P(P(C.c_float)), P(C.c_int), P(P(C.c_int)),
                                                              P(P(C.c_size_t))]),
    "ocrs_b200_pool_create": (C.c_int, [P(PoolParamsC), P(C.c_void_p)]),
    "ocrs_b200_pool_destroy": (None, [C.c_void_p]),
    "ocrs_b200_pool_submit": (C.c_int, [C.c_void_p, P(PageC), C.c_size_t, P(C.c_uint64)]),
    "ocrs_b200_pool_wait": (C.c_int, [C.c_void_p, C.c_uint64, P(P(TextResultC)), C.c_size_t]),
    "ocrs_b200_pool_wait_text": (C.c_int, [C.c_void_p, C.c_uint64, P(C.c_void_p), C.c_size_t]),
    "ocrs_b200_pool_done": (C.c_int, [C.c_void_p, C.c_uint64, P(C.c_int)]),
    "ocrs_b200_pool_shape": (C.c_int, [C.c_void_p, P(C.c_int), P(C.c_int)]),
    "ocrs_b200_pool_engine": (C.c_int, [C.c_void_p, C.c_int, C.c_int, P(C.c_void_p)]),
    "ocrs_b200_pool_describe": (C.c_int, [C.c_void_p, P(C.c_void_p)]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "ocrs_b200 has no Python/CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.ocrs_b200_last_error()
        raise OcrsError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")
