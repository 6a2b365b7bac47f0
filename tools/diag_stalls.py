"""Looks for sporadic stalls: times every recognize/ocr call of a steady loop, one or two engines
in flight, and prints the slowest calls with the engine's host-section timers."""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrs_b200 as ob
from tools.models import ensure_models
from tools.synth import make_page

det, rec = ensure_models()
n_eng = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
pages = [np.ascontiguousarray(make_page(200 + i)[0]) for i in range(8)]
engines = [ob.OcrEngine(ob.OcrEngineParams(detection_model=det, recognition_model=rec)) for _ in range(n_eng)]
inputs = [[e.prepare_input(ob.ImageSource.from_tensor(p, ob.DimOrder.Hwc)) for p in pages] for e in engines]
times = [[] for _ in engines]


def loop(k):
    e = engines[k]
    for it in range(iters):
        t0 = time.perf_counter()
        e.ocr_batch_text(inputs[k])
        dt = (time.perf_counter() - t0) * 1e3
        prof = e.profile(reset=True)
        times[k].append((dt, it, {n: round(v["ms"], 1) for n, v in prof.items() if n.startswith("host/")}))


ts = [threading.Thread(target=loop, args=(k,)) for k in range(n_eng)]
t0 = time.perf_counter()
[t.start() for t in ts]
[t.join() for t in ts]
wall = time.perf_counter() - t0
print(f"{n_eng} engines x {iters} batches of 8 pages: {n_eng * iters * 8 / wall:.1f} pages/s")
for k in range(n_eng):
    ds = sorted(times[k], key=lambda v: -v[0])
    med = np.median([d[0] for d in ds])
    print(f"engine {k}: median {med:.1f} ms; slowest:")
    for d in ds[:4]:
        print("   ", round(d[0], 1), "ms at iter", d[1], d[2])
