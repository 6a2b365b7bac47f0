"""Where does the recognition log-prob error come from?  Runs the CRNN on real-looking line
images under each kill switch (one subprocess per setting: the switches are read at load) and
prints error statistics against the oracle (torch-CPU fp32).
usage: python tools/diag_rec_error.py            (driver)
       python tools/diag_rec_error.py --child    (one setting)"""
import json
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import ocrs_b200 as ob
    from oracle.onnx_eval import OnnxModel
    from tools.models import ensure_models
    from tools.synth import make_line_batch
    _, rec = ensure_models()
    x = make_line_batch(300, 64)
    exp = OnnxModel(rec).run(x)
    got = ob.Model(rec).run(x)
    err = np.abs(got - exp)
    i = np.unravel_index(err.argmax(), err.shape)
    top = exp.argmax(-1)
    top_err = np.take_along_axis(err, top[..., None], -1)
    out = {
        "max_abs": float(err.max()), "at": [int(v) for v in i], "exp_at": float(exp[i]), "got_at": float(got[i]),
        "mean_abs": float(err.mean()), "p999": float(np.quantile(err, 0.999)),
        "max_abs_top1": float(top_err.max()),
        "max_abs_where_logp>-10": float(err[exp > -10].max()),
        "max_rel": float((err / np.maximum(np.abs(exp), 1.0)).max()),
        "argmax_mismatch": int((got.argmax(-1) != top).sum()),
    }
    print(json.dumps(out))


def main():
    settings = [{}, {"OCRS_B200_DISABLE_TC_GRU": "1"}, {"OCRS_B200_DISABLE_SEQ_HEAD": "1"},
                {"OCRS_B200_DISABLE_TC": "1"}, {"OCRS_B200_DISABLE_TC": "1", "OCRS_B200_DISABLE_TC_GRU": "1"}]
    for s in settings:
        env = dict(os.environ, **s)
        r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        print(s or "default", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:])


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
