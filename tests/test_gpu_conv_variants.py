"""The conv stack has one kernel per layer shape plus fallbacks behind environment switches (read once per
process).  Every variant must stay inside the same 1e-3 log-prob tolerance against the torch-CPU interpreter, so
each one runs the recognition net in its own process and the results are compared here."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle.onnx_eval import OnnxModel
from tests.gpu_util import model_paths

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOGIT_TOL = 1e-3

SNIPPET = """
import sys, numpy as np
import ocrs_b200 as ob
rec = ob.Model(sys.argv[1])
x = np.load(sys.argv[2])
np.save(sys.argv[3], rec.run(x))
"""

VARIANTS = {
    "default (halo pair + weights-stationary)": {},
    "32->64 layer: resident weights": {"OCRS_B200_CONV_WS": "0"},
    "32->64 layer: streaming kernel": {"OCRS_B200_CONV_WS": "0", "OCRS_B200_CONV_RES": "0"},
    "128-channel layers: CTA pair": {"OCRS_B200_CONV_MODE": "1"},
    "128-channel layers: single CTA": {"OCRS_B200_CONV_MODE": "0"},
}


def test_conv_kernel_variants_agree_with_the_cpu_interpreter():
    _, rec_path = model_paths()
    x = np.random.default_rng(7).uniform(-0.5, 0.5, (3, 1, 64, 520)).astype(np.float32)
    exp = OnnxModel(rec_path).run(x)
    with tempfile.TemporaryDirectory() as tmp:
        xin = os.path.join(tmp, "x.npy")
        np.save(xin, x)
        outs = {}
        for name, env in VARIANTS.items():
            out = os.path.join(tmp, "y%d.npy" % len(outs))
            e = dict(os.environ, **env)
            e["PYTHONPATH"] = ROOT + os.pathsep + e.get("PYTHONPATH", "")
            r = subprocess.run([sys.executable, "-c", SNIPPET, rec_path, xin, out], env=e, cwd=ROOT, capture_output=True, text=True,
                               timeout=300)
            assert r.returncode == 0, (name, r.stderr[-2000:])
            outs[name] = np.load(out)
    for name, got in outs.items():
        assert got.shape == exp.shape, name
        assert np.abs(got - exp).max() < LOGIT_TOL, (name, float(np.abs(got - exp).max()))
    # the variants differ only in summation order and in whether partial sums are promoted
    base = outs["default (halo pair + weights-stationary)"]
    for name, got in outs.items():
        assert np.abs(got - base).max() < LOGIT_TOL, name
