"""Micro-benchmark of the recognition network through the C ABI engine profile hooks.
usage: python tools/microbench_rec.py [N] [W]"""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrs_b200 as ob
from tools.models import ensure_models

det, rec = ensure_models()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2400
eng = ob.OcrEngine(ob.OcrEngineParams(recognition_model=rec))
# a page tall enough to hold N lines of height 20 and width such that resized width == W
Hl = 20
Wl = int(W * Hl / 64)
page = (np.random.default_rng(0).random((N * (Hl + 4) + 8, Wl + 8, 1)) * 255).astype(np.uint8)
inp = eng.prepare_input(ob.ImageSource.from_tensor(page, ob.DimOrder.Hwc))
lines = [[ob.RotatedRect(4 + Wl / 2, 4 + i * (Hl + 4) + Hl / 2, 0.0, 1.0, float(Wl), float(Hl))] for i in range(N)]
for _ in range(2):
    eng.recognize_text(inp, lines)
eng.set_profiling(True)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    eng.recognize_text(inp, lines)
dt = (time.perf_counter() - t0) / reps
prof = eng.profile()
print(f"N={N} W={W} wall {dt*1e3:.2f} ms")
for k, v in sorted(((k, v) for k, v in prof.items() if not k.startswith("host/")), key=lambda kv: -kv[1]["ms"])[:12]:
    print(f"  {k:24s} {v['ms']/reps:9.3f} ms  calls {v['calls']//reps:4d}  {v['flops']/max(v['ms'],1e-9)/1e9:9.1f} TFLOP/s")
