#!/usr/bin/env python
"""bench.py -- pages/sec of the OCR hot path (detect -> line-group -> recognise -> CTC).

    python bench.py --gpus N --steps K --warmup W          # the B200 arm (this repo)
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[2]): full pipeline on a batch of 8 synthetic 1024x768 RGB pages per
GPU.  One "step" = one pass of the whole hot path over that batch.  Rank r of N processes its own
8 pages (weak scaling); with N > 1 the recognised text of every rank is gathered to rank 0 over NCCL
inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PAGE_H, PAGE_W, BATCH = 768, 1024, 8
METRIC = "pages/sec end-to-end (detect+recognise)"
CPU_THREADS = 32  # torch-CPU threads of the oracle port (more only adds contention on these small ops)
WORKLOAD = "full pipeline (detect+line-group+recognise+CTC), batch=8 1024x768 synthetic pages per GPU"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fp:
            d = json.load(fp)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clocks / clock-event reasons sampled DURING the timed region (B200_PROFILING.md's clocks
    line).  NVML is initialised in `prepare()` -- before the warm-up -- because attaching a new
    NVML client (what spawning `nvidia-smi` does) can stall the GPU for >100 ms; the timed region
    then only sees light in-process queries every 50 ms.  Falls back to an `nvidia-smi -lms` child
    started in prepare() whose samples are filtered to the timed window."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.nvml = None
        self.handle = None
        self.proc = None
        self.samples = []   # (t, sm_mhz, sm_max_mhz, [reasons])
        self.active = False
        self.alive = False
        self.t0 = self.t1 = None

    def prepare(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:  # noqa: BLE001
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.alive = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {}
        for nm, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                         ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                         ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                         ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            alt = attr.replace("ClocksEventReason", "ClocksThrottleReason")
            v = getattr(n, attr, None) or getattr(n, alt, None)
            if v is not None:
                bits[nm] = int(v)
        while self.alive:
            if self.active:
                try:
                    sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                    try:
                        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    except Exception:  # noqa: BLE001
                        r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    self.samples.append((time.perf_counter(), sm, self.smax, [k for k, b in bits.items() if r & b]))
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.05)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 9:
                continue
            try:
                self.samples.append((time.perf_counter(), float(f[1]), float(f[2]),
                                     [nm for nm, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue

    def start(self):
        self.t0 = time.perf_counter()
        self.active = True

    def stop(self):
        self.t1 = time.perf_counter()
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.active = False
        self.alive = False
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        win = [s for s in self.samples if self.t0 <= s[0] <= self.t1 + 0.12]
        sm = [s[1] for s in win]
        smax = [s[2] for s in win]
        reasons = sorted({r for s in win for r in s[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_traffic(kernel: str):
    """DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r01b_conv_traffic.json")
    try:
        with open(path) as fp:
            d = json.load(fp)
        return d if d.get("kernel") == kernel else None
    except Exception:  # noqa: BLE001
        return None


def make_batch(rank: int):
    from tools.synth import make_page
    return [np.ascontiguousarray(make_page(200 + rank * BATCH + i, PAGE_H, PAGE_W)[0]) for i in range(BATCH)]


# =============================================================================================
# reference arm: the reference's CPU path (oracle port; rten cannot be built here -- no Rust)
# =============================================================================================
# The CPU port is single-threaded Python around torch-CPU networks; pages are independent, so the
# host cores are used the way the GPU arm uses the GPU: one worker process per page of the batch,
# torch threads split between the workers.
_ORACLE = None


def _oracle_worker_init(det, rec, threads):
    global _ORACLE
    import torch
    torch.set_num_threads(max(1, threads))
    from oracle.engine import OcrEngine as OEngine, OcrEngineParams as OParams
    from oracle.onnx_eval import OnnxModel
    _ORACLE = OEngine(OParams(detection_model=OnnxModel(det), recognition_model=OnnxModel(rec)))


def _oracle_worker_page(page):
    return _ORACLE.get_text(_ORACLE.prepare_input(page, "hwc"))


class OraclePool:
    """`workers` processes, each holding the oracle engine (models loaded once)."""

    def __init__(self, det, rec):
        import multiprocessing as mp
        self.cores = min(os.cpu_count() or 1, CPU_THREADS)
        self.workers = max(1, min(BATCH, self.cores))
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.workers, initializer=_oracle_worker_init,
                             initargs=(det, rec, self.cores // self.workers))
        self.pool.map(_oracle_worker_page, [np.zeros((64, 64, 3), np.uint8)] * self.workers)  # imports, model load

    def run_batch(self, pages):
        return self.pool.map(_oracle_worker_page, list(pages), chunksize=1)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tools.models import ensure_models
    det, rec = ensure_models()
    pages = make_batch(0)
    pool = OraclePool(det, rec)
    for _ in range(args.warmup):
        pool.run_batch(pages)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.run_batch(pages)
    dt = time.perf_counter() - t0
    pool.close()
    value = BATCH * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": 0, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"one {BATCH}-page batch, one page per worker process",
                   "weights": "models/*.onnx (synthetic stand-ins; the reference's weights are not available offline)"},
        "cpu_baseline": {"value": value, "unit": "pages/s", "cores": pool.cores, "kind": "port",
                         "sample": f"{args.steps} batches of {BATCH} pages over {pool.workers} worker processes x "
                                   f"{max(1, pool.cores // pool.workers)} torch threads, oracle port (torch-CPU fp32 nets + "
                                   "python/numpy post-processing); NOT rten -- no Rust toolchain or weights in this "
                                   "environment"},
        "e2e": {"value": value, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# =============================================================================================
# B200 arm
# =============================================================================================
def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import ocrs_b200 as ob
    from ocrs_b200.dist import gather_texts
    from tools.models import ensure_models

    det, rec = ensure_models()
    eng = ob.OcrEngine(ob.OcrEngineParams(detection_model=det, recognition_model=rec, device=local))
    # `--in-flight 2`: a second engine instance (own streams and buffers) on the same GPU lets the host
    # phases of one batch (layout analysis, result assembly) overlap the kernels of the other
    engines = [eng] + [ob.OcrEngine(ob.OcrEngineParams(detection_model=det, recognition_model=rec, device=local))
                       for _ in range(max(1, args.in_flight) - 1)]
    pages = make_batch(rank)
    # host copies in pinned memory (e2e) and device copies (kernel-only `value`)
    pinned = [torch.from_numpy(p).pin_memory() for p in pages]
    resident = [t.cuda(local) for t in pinned]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_text(results):
        """NCCL gather of the recognised text to rank 0 (two-phase: lengths, then padded bytes)."""
        if world > 1:
            gather_texts(results, device=f"cuda:{local}")

    def step_resident(e=None):
        e = e or eng
        inputs = [e.prepare_input_device(t.data_ptr(), 0, ob.DimOrder.Hwc, PAGE_H, PAGE_W, 3) for t in resident]
        return e.ocr_batch_text(inputs)

    def step_e2e(e=None):
        e = e or eng
        inputs = [e.prepare_input(ob.ImageSource(t.numpy(), ob.DimOrder.Hwc)) for t in pinned]
        return e.ocr_batch_text(inputs)

    step_done = []  # completion time of every step (diagnostic: largest gap is reported)

    def run_steps(fn, steps):
        """Runs `steps` batches, at most len(engines) in flight (one host thread per engine); the text
        of every finished batch is gathered to rank 0."""
        if len(engines) == 1:
            for _ in range(steps):
                gather_text(fn(eng))
                step_done.append(time.perf_counter())
            return
        import queue
        todo = queue.Queue()
        for i in range(steps):
            todo.put(i)
        done = queue.Queue()
        errs = []

        def worker(e):
            try:
                while True:
                    try:
                        todo.get_nowait()
                    except queue.Empty:
                        return
                    done.put(fn(e))
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
                done.put(None)

        ts = [threading.Thread(target=worker, args=(e,)) for e in engines]
        [t.start() for t in ts]
        for _ in range(steps):
            res = done.get()
            step_done.append(time.perf_counter())
            if res is None:
                break
            gather_text(res)  # collectives stay on the main thread, in completion order
        [t.join() for t in ts]
        if errs:
            raise errs[0]

    gaps = []  # per timed() call: largest interval between two step completions, ms

    def timed(fn, steps):
        barrier()
        launches0 = ob.kernel_launch_count()
        tb0 = [e.transfer_bytes() for e in engines]
        eng.timer_start()
        t0 = time.perf_counter()
        del step_done[:]
        step_done.append(t0)
        run_steps(fn, steps)
        gaps.append(max(b - a for a, b in zip(step_done, step_done[1:])) * 1e3 if len(step_done) > 1 else 0.0)
        for e2 in engines[1:]:
            e2.timer_start()  # (orders a marker behind everything enqueued on that engine's stream)
            e2.timer_stop()
        ms = eng.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(ms, wall) if len(engines) > 1 else ms  # several streams: the host clock brackets them all
        barrier()
        tb1 = [e.transfer_bytes() for e in engines]
        h2d0 = sum(t[0] for t in tb0); d2h0 = sum(t[1] for t in tb0)
        h2d1 = sum(t[0] for t in tb1); d2h1 = sum(t[1] for t in tb1)
        t = torch.tensor([ms, wall], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), ob.kernel_launch_count() - launches0, (h2d1 - h2d0) / steps, (d2h1 - d2h0) / steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()  # attach the NVML client before the warm-up, not inside the timed region
    # warm-up runs the whole step, including the gather (the first NCCL collective builds the
    # communicator, ~100 ms, and must not land in the timed region)
    for e in engines:
        for _ in range(args.warmup):
            gather_text(step_resident(e))
        for _ in range(max(1, args.warmup // 2)):
            gather_text(step_e2e(e))

    if rank == 0:
        sampler.start()
    eng.profile(reset=True)  # clear host-section timers
    ms, wall, launches, _, _ = timed(step_resident, args.steps)
    host_real = {k: round(v["ms"] / args.steps, 3) for k, v in eng.profile(reset=True).items() if k.startswith("host/")}
    ms_e2e, wall_e2e, _, h2d, d2h = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel roofline: profile one more pass with CUDA events around every operator ----
    eng.set_profiling(True)
    eng.stats(reset=True)  # word / line counts below are those of the profiled passes only
    barrier()
    res = None
    for _ in range(max(1, min(3, args.steps))):
        res = step_resident(eng)
    prof = eng.profile(reset=True)
    eng.set_profiling(False)
    stats = eng.stats(reset=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    total_pages = BATCH * world
    value = total_pages * args.steps / (ms / 1e3)
    e2e_value = total_pages * args.steps / (ms_e2e / 1e3)

    # dominant kernel = the profiled operator class with the largest share of device time
    ops = {k: v for k, v in prof.items()
           if not k.startswith("stage/") and not k.startswith("host/") and not k.endswith("(total)")}
    host_ms = {k: round(v["ms"] / max(1, min(3, args.steps)), 3) for k, v in prof.items() if k.startswith("host/")}
    dom_name, dom = max(ops.items(), key=lambda kv: kv[1]["ms"]) if ops else ("none", None)
    stage_ms = {k: round(v["ms"] / max(1, min(3, args.steps)), 3) for k, v in prof.items() if k.startswith("stage/")}
    roofline = None
    if dom is not None and dom["launches"] > 0:
        sec_per_launch = dom["ms"] / 1e3 / dom["launches"]
        if dom["flops"] > 0:
            achieved = dom["flops"] / dom["launches"] / sec_per_launch / 1e12
            roofline = {"kernel": dom_name, "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"],
                        "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None,
                        "peak_source": peaks["source"] + " bf16 sustained (kernel timed inside the step)",
                        "share_of_step": dom["ms"] / max(sum(v["ms"] for v in ops.values()), 1e-9),
                        "launches_per_step": dom["launches"] / max(1, min(3, args.steps)),
                        "algorithmic_bytes_per_launch": dom["bytes"] / dom["launches"],
                        "note": "achieved = fp32-equivalent conv/GEMM FLOPs (2*MACs) per launch / CUDA-event time; the "
                                "kernel issues 3 fp16 MMAs per product term (split operands), so its own ceiling is "
                                "peak/3; tensor pipe 51% active under ncu: the MMA thread waits 25% of its time for the "
                                "accumulator promotion and spends 20% in issue overhead (profiles/r01b_ncu_summary.md)"}
            tr = measured_traffic(dom_name)
            if tr is not None:
                roofline["traffic"] = tr["dram_bytes_per_algorithmic_byte"] * roofline["algorithmic_bytes_per_launch"]
                roofline["traffic_source"] = ("DRAM bytes (read+write) per algorithmic byte = %.3f from " % tr["dram_bytes_per_algorithmic_byte"]
                                              + tr["source"] + ", applied to this run's algorithmic bytes per launch")
        else:
            achieved = dom["bytes"] / dom["launches"] / sec_per_launch / 1e9
            roofline = {"kernel": dom_name, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"]}

    # ---- CPU baseline: oracle port on a bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            pool = OraclePool(det, rec)
            n_rep = 2
            t0 = time.perf_counter()
            for _ in range(n_rep):
                ref_text = pool.run_batch(pages)
            dt = time.perf_counter() - t0
            pool.close()
            cpu = {"value": n_rep * BATCH / dt, "unit": "pages/s", "cores": pool.cores, "kind": "port",
                   "sample": f"{n_rep} x the {BATCH}-page batch over {pool.workers} worker processes x "
                             f"{max(1, pool.cores // pool.workers)} torch threads, oracle port (torch-CPU fp32 nets + "
                             "python/numpy post-processing; NOT rten)",
                   "text_identical_to_gpu": [res[i] for i in range(BATCH)] == list(ref_text)}
        except Exception as ex:  # noqa: BLE001  (worker processes unavailable: time two pages in this process)
            from oracle.engine import OcrEngine as OEngine, OcrEngineParams as OParams
            from oracle.onnx_eval import OnnxModel
            cores = min(os.cpu_count() or 1, CPU_THREADS)
            torch.set_num_threads(cores)
            ora = OEngine(OParams(detection_model=OnnxModel(det), recognition_model=OnnxModel(rec)))
            n_sample = 2
            t0 = time.perf_counter()
            ref_text = [ora.get_text(ora.prepare_input(pages[i], "hwc")) for i in range(n_sample)]
            dt = time.perf_counter() - t0
            cpu = {"value": n_sample / dt, "unit": "pages/s", "cores": cores, "kind": "port",
                   "sample": f"first {n_sample} pages of the batch, sequentially, through the oracle port (worker pool failed: "
                             f"{type(ex).__name__}); NOT rten",
                   "text_identical_to_gpu": [res[i] for i in range(n_sample)] == ref_text}

    line = {
        "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "pages_per_gpu": BATCH, "page_hw": [PAGE_H, PAGE_W], "batches_in_flight": len(engines),
                   "weights": "models/*.onnx", "l2": "inputs per step (18.9 MB u8 + activations >> 126 MB L2 over a step)",
                   "words_per_page": stats["words"] / max(1, BATCH * max(1, min(3, args.steps))),
                   "lines_per_page": stats["lines"] / max(1, BATCH * max(1, min(3, args.steps)))},
        "e2e": {"value": e2e_value, "unit": "pages/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "stage_ms_per_step": stage_ms, "host_ms_per_step": host_real, "host_ms_per_step_serial_profile": host_ms,
        "wall_ms_per_step": wall / args.steps,
        "max_step_gap_ms": {"value": round(gaps[0], 2), "e2e": round(gaps[1], 2)} if len(gaps) >= 2 else None,
        "op_ms_per_step": {k: round(v["ms"] / max(1, min(3, args.steps)), 3) for k, v in ops.items()},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=2, help="batches in flight per GPU (engine instances)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
