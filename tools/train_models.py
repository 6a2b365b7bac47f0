"""Trains the two stand-in networks on the synthetic page generator (CPU, torch) and exports them
to models/*.onnx.  Fixture tooling, run once offline; the resulting files are committed so that the
GPU box never needs to train.  With trained weights the pipeline reads real text, which makes the
end-to-end parity tests (identical strings) meaningful and robust.

    python tools/train_models.py det --minutes 15
    python tools/train_models.py rec --minutes 45
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import cv2
import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.engine import DEFAULT_ALPHABET  # noqa: E402
from oracle.imageops import pad_bottom_right, prepare_image, resize_bilinear  # noqa: E402
from tools.models import (  # noqa: E402
    DET_INPUT_HW, DetectionNet, RecognitionNet, default_model_dir, export_detection, export_recognition,
)
from tools.synth import make_page  # noqa: E402


def log(msg):
    print(time.strftime("%H:%M:%S"), msg, flush=True)


# ---------------------------------------------------------------------------------------------
def det_sample(seed):
    """Page -> (net input [800,600] f32, target mask [800,600] f32) exactly as detection.rs:155-171
    prepares it (pad bottom/right with -0.5, bilinear resize)."""
    rng = np.random.default_rng(seed)
    h, w = (768, 1024) if rng.random() < 0.7 else (int(rng.integers(400, 900)), int(rng.integers(500, 1100)))
    page, _, boxes = make_page(seed, h, w, with_boxes=True)
    grey = prepare_image(page, "hwc")
    in_h, in_w = DET_INPUT_HW
    pb, pr = max(in_h - h, 0), max(in_w - w, 0)
    x = resize_bilinear(pad_bottom_right(grey[None], pb, pr), in_h, in_w)[0, 0]
    sy, sx = in_h / (h + pb), in_w / (w + pr)
    target = np.zeros((in_h, in_w), np.float32)
    for row in boxes:
        for (t, l, b, r) in row:
            t2, b2 = (t + 3) * sy, (b - 3) * sy
            l2, r2 = (l + 4) * sx, (r - 4) * sx
            if b2 - t2 < 2 or r2 - l2 < 2:
                continue
            target[int(round(t2)):int(round(b2)), int(round(l2)):int(round(r2))] = 1.0
    return x, target


def train_det(minutes: float, out_path: str, threads: int):
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = DetectionNet().train()
    opt = torch.optim.Adam(net.parameters(), lr=3e-3)
    ch, cw = 384, 288
    t0 = time.time()
    step = 0
    seed = 10_000
    pool = []
    bce = nn.BCELoss()
    while time.time() - t0 < minutes * 60:
        if len(pool) < 8 or step % 4 == 0:
            pool.append(det_sample(seed))
            seed += 1
            pool = pool[-24:]
        xs, ys = [], []
        rng = np.random.default_rng(step)
        for _ in range(6):
            x, y = pool[int(rng.integers(0, len(pool)))]
            oy, ox = int(rng.integers(0, x.shape[0] - ch + 1)), int(rng.integers(0, x.shape[1] - cw + 1))
            xs.append(x[oy:oy + ch, ox:ox + cw])
            ys.append(y[oy:oy + ch, ox:ox + cw])
        xb = torch.from_numpy(np.stack(xs))[:, None]
        yb = torch.from_numpy(np.stack(ys))[:, None]
        for g in opt.param_groups:
            g["lr"] = 3e-3 * (0.5 * (1 + np.cos(np.pi * min(1.0, (time.time() - t0) / (minutes * 60))))) + 1e-4
        pred = net(xb).clamp(1e-6, 1 - 1e-6)
        w = 1.0 + 2.0 * yb
        loss = (nn.functional.binary_cross_entropy(pred, yb, reduction="none") * w).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 25 == 0:
            with torch.no_grad():
                iou = ((pred > 0.5) & (yb > 0.5)).sum().item() / max(((pred > 0.5) | (yb > 0.5)).sum().item(), 1)
            log(f"det step {step} loss {loss.item():.4f} iou {iou:.3f}")
        step += 1
    net.eval()
    export_detection(net, out_path)
    torch.save(net.state_dict(), out_path + ".pt")
    log(f"det: wrote {out_path} after {step} steps")


# ---------------------------------------------------------------------------------------------
def rec_lines(seed, max_width=1300):
    """Rows of a synthetic page -> [(line image f32 [64, w'], label ids)] cropped like
    recognition.rs:91-126 (axis-aligned rows; word boxes expanded by 3 px as detection does)."""
    rng = np.random.default_rng(seed)
    two_col = rng.random() < 0.8
    page, texts, boxes = make_page(seed, 768, 1024, two_col=two_col, with_boxes=True)
    grey = prepare_image(page, "hwc")[0]
    out = []
    for text, row in zip(texts, boxes):
        t = min(b[0] for b in row) - 3
        l = min(b[1] for b in row) - 3
        b_ = max(b[2] for b in row) + 3
        r = max(b[3] for b in row) + 3
        t, l = max(t, 0), max(l, 0)
        b_, r = min(b_, grey.shape[0]), min(r, grey.shape[1])
        crop = grey[t:b_, l:r]
        w2 = int(min(max(64 * crop.shape[1] / crop.shape[0], 10), 2400))
        if w2 > max_width:
            continue
        img = cv2.resize(crop, (w2, 64), interpolation=cv2.INTER_LINEAR)
        ids = [DEFAULT_ALPHABET.index(c) + 1 for c in text]
        out.append((img.astype(np.float32), ids))
    return out


_PIPE = {}


def rec_lines_pipeline(seed, max_width=1300):
    """Like rec_lines, but the crops come from the oracle pipeline itself (trained detection net ->
    word rects -> find_text_lines -> line polygon crop), so the recogniser sees exactly what the
    engine feeds it.  Lines are labelled with the ground-truth row they cover word for word."""
    from oracle.engine import OcrEngine, OcrEngineParams
    from oracle.onnx_eval import OnnxModel
    if "eng" not in _PIPE:
        det = os.path.join(default_model_dir(), "text-detection.onnx")
        _PIPE["eng"] = OcrEngine(OcrEngineParams(detection_model=OnnxModel(det)))
        from oracle.engine import TextRecognizer

        class _Shape:
            def input_shape(self):
                return ["batch", 1, 64, "seq"]
        _PIPE["rec"] = TextRecognizer(_Shape())
    eng, recog = _PIPE["eng"], _PIPE["rec"]
    rng = np.random.default_rng(seed)
    page, texts, boxes = make_page(seed, 768, 1024, two_col=bool(rng.random() < 0.8), with_boxes=True)
    img = eng.prepare_input(page, "hwc")
    lines = eng.find_text_lines(img, eng.detect_words(img))
    out = []
    for line in lines:
        centers = [(float(w.cy), float(w.cx)) for w in line]
        label = None
        for text, row in zip(texts, boxes):
            if len(row) != len(line):
                continue
            if all(b[0] - 2 <= cy <= b[2] + 2 and b[1] - 2 <= cx <= b[3] + 2 for (cy, cx), b in zip(centers, row)):
                label = text
                break
        if label is None:
            continue
        crop = recog.prepare_input(img, line)
        if crop.shape[1] > max_width:
            continue
        out.append((crop.astype(np.float32), [DEFAULT_ALPHABET.index(c) + 1 for c in label]))
    return out


def collate(items):
    wmax = max(i[0].shape[1] for i in items)
    wg = ((wmax + 49) // 50) * 50
    x = np.full((len(items), 1, 64, wg), -0.5, np.float32)
    for k, (img, _) in enumerate(items):
        x[k, 0, :, : img.shape[1]] = img
    targets = torch.tensor([c for _, ids in items for c in ids], dtype=torch.long)
    lens = torch.tensor([len(ids) for _, ids in items], dtype=torch.long)
    collate.widths = [it[0].shape[1] for it in items]
    return torch.from_numpy(x), targets, lens


def greedy(logp):
    lab = logp.argmax(-1).T.tolist()  # [N][T]
    outs = []
    for seq in lab:
        s, last = [], 0
        for l in seq:
            if l != last and l > 0:
                s.append(l)
            last = l
        outs.append(s)
    return outs


def train_rec(minutes: float, out_path: str, threads: int, resume: str | None = None, pipeline: bool = False,
              lr0: float = 1e-3):
    torch.set_num_threads(threads)
    torch.manual_seed(1)
    net = RecognitionNet().train()
    if resume and os.path.exists(resume):
        net.load_state_dict(torch.load(resume))
        log(f"resumed from {resume}")
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    ctc = nn.CTCLoss(blank=0, zero_infinity=True)
    t0 = time.time()
    step, seed = 0, 50_000
    pool = []
    gen = rec_lines_pipeline if pipeline else rec_lines
    val = [it for s in range(900_000, 900_003) for it in gen(s)][:48]
    best = -1.0
    while time.time() - t0 < minutes * 60:
        while len(pool) < 64:
            pool.extend(gen(seed))
            seed += 1
        pool.sort(key=lambda it: it[0].shape[1] + 40 * np.random.rand())
        k = int(np.random.randint(0, max(1, len(pool) - 8)))
        items = pool[k:k + 8]
        del pool[k:k + 8]
        x, targets, lens = collate(items)
        frac = min(1.0, (time.time() - t0) / (minutes * 60))
        for g in opt.param_groups:
            g["lr"] = lr0 * (0.5 * (1 + np.cos(np.pi * frac))) + 2e-5
        logp = net(x)  # [T, N, C]
        T = logp.shape[0]
        # characters decoded at x >= line width are discarded by the pipeline (recognition.rs:278):
        # forbid non-blank emissions in the right padding so every character is emitted inside the line
        mask = torch.zeros_like(logp)
        for k, wk in enumerate(collate.widths):
            tv = max((wk - 4) // 4, 1)
            mask[tv:, k, 1:] = -1e4
        logp = logp + mask
        loss = ctc(logp, targets, torch.full((len(items),), T, dtype=torch.long), lens)
        opt.zero_grad()
        loss.backward()
        nn.utils.clip_grad_norm_(net.parameters(), 5.0)
        opt.step()
        if step % 20 == 0:
            log(f"rec step {step} loss {loss.item():.4f} lr {opt.param_groups[0]['lr']:.2e}")
        if step % 100 == 99:
            net.eval()
            with torch.no_grad():
                ok = 0
                for i in range(0, len(val), 8):
                    xv, _, _ = collate(val[i:i + 8])
                    lp = net(xv)
                    for k, wk in enumerate(collate.widths):
                        lp[max(-(-wk // 4), 1):, k, 1:] = -1e4  # what the pipeline would drop
                    for got, (_, ids) in zip(greedy(lp), val[i:i + 8]):
                        ok += int(got == ids)
            acc = ok / len(val)
            log(f"rec step {step} val exact-line accuracy {acc:.3f}")
            if acc >= best:
                best = acc
                export_recognition(net, out_path)
                torch.save(net.state_dict(), out_path + ".pt")
            net.train()
        step += 1
    log(f"rec: best val accuracy {best:.3f} after {step} steps -> {out_path}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["det", "rec"])
    ap.add_argument("--minutes", type=float, default=10)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--out", default=None)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--pipeline", action="store_true")
    ap.add_argument("--lr", type=float, default=1e-3)
    a = ap.parse_args()
    d = default_model_dir()
    os.makedirs(d, exist_ok=True)
    if a.which == "det":
        train_det(a.minutes, a.out or os.path.join(d, "text-detection.trained.onnx"), a.threads)
    else:
        train_rec(a.minutes, a.out or os.path.join(d, "text-recognition.trained.onnx"), a.threads, a.resume, a.pipeline, a.lr)
