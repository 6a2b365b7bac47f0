"""The C ABI from plain C: tests/c/abi_smoke.c is compiled with gcc against include/ocrs_b200.h and
linked with libocrs_b200.so only -- the situation of the reference's Rust binding."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ocrs_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cabi") / "abi_smoke")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", out, "-L" + LIBDIR, "-locrs_b200",
           "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return out


def test_host_entry_points_from_c(exe):
    r = subprocess.run([exe, "host"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert lines[0].startswith("version ocrs_b200")
    assert lines[1] == "lines 2: [30 90 150] [30 90 150]"      # two rows, words ordered left to right
    assert lines[2] == "text <hi yo>"
    assert lines[4] == "vertices 50,25 0,25 0,0 50,0"
    assert lines[5].startswith("empty item -> -1 (")            # OCRS_B200_ERR_INVALID_ARG + message, no abort


@pytest.mark.gpu
def test_full_pipeline_from_c(exe, tmp_path):
    """engine_create -> prepare_input_bytes -> get_text from C equals the Python mirror's result."""
    import ocrs_b200 as ob
    from tests.gpu_util import model_paths
    from tools.synth import make_page
    det, rec = model_paths()
    page, _ = make_page(11, 384, 640, n_rows=5)
    raw = tmp_path / "page.rgb"
    raw.write_bytes(np.ascontiguousarray(page).tobytes())
    r = subprocess.run([exe, "gpu", det, rec, str(raw), "640", "384"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    eng = ob.OcrEngine(ob.OcrEngineParams(detection_model=det, recognition_model=rec))
    want = eng.get_text(eng.prepare_input(ob.ImageSource.from_bytes(page.tobytes(), (640, 384))))
    assert r.stdout.rstrip("\n") == want
    assert len(want.split("\n")) >= 3
