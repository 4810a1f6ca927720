#!/usr/bin/env python
"""Benchmark of the hot path: training tokens/s of the NLPScratch Transformer step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|...]

A "step" is one full-sequence training step (forward, softmax-CE, backward, [all-reduce], global-norm
clip, Adam) over one batch of synthetic token ids.  Default workload = BASELINE.json configs[3]
("c4": 6 layers, d_model 512, 8 heads, ff 2048, seq 512, batch 64 per GPU, vocab 8192), the
configuration the metric is quoted on; weak scaling (per-GPU batch fixed).

Prints ONE JSON line (rank 0).  `value` = device-resident tokens/s; `e2e` = the same metric through
the public API with pinned-host -> device copies of x,y and a device -> host read of every step's loss
inside the timed region.  `roofline` = the dominant kernel (tcgen05 FFN GEMM) against the tcgen05 TF32
peak measured in the same run; `cpu_baseline` = the unmodified reference (baseline/_ref, installed by
tools/install_reference.py; the NumPy oracle port only if it is absent) on the host cores; `other` = the
FP32-faithful arithmetic modes and the other BASELINE.json configurations, a few steps each.
`--impl reference` times that CPU implementation alone, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (L, E, h, F, S, B/gpu, V)
    "c4": (6, 512, 8, 2048, 512, 64, 8192),
    "c4-32k": (6, 512, 8, 2048, 512, 64, 32768),
    "c5": (12, 768, 12, 3072, 2048, 8, 8192),
    "c3": (3, 256, 8, 1024, 320, 16, 32768),
    "c2": (3, 256, 8, 1024, 320, 16, 574),
    "tiny": (2, 64, 2, 128, 32, 4, 120),
    # the reference's own default script (BASELINE.json configs[0]): prefix dataset, last-token loss,
    # driven through the train_transformer loop; S is max_len, sequences are trimmed per batch
    "c1": (3, 256, 8, 1024, 320, 16, 574),
}
METRIC = "train tokens/s"


def flops_per_token(L, E, F, S, V):
    return L * (24.0 * E * E + 12.0 * E * F + 6.0 * S * E) + 6.0 * E * V


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index
        self.begin, self.end = 0, None

    # nvidia-smi's start-up and tear-down (NVML initialisation) can stall CUDA calls of other processes for tens of
    # milliseconds, so the sampler is started BEFORE the warm-up and stopped AFTER the last timed leg; only the samples
    # taken between mark_begin() and mark_end() -- the timed region -- are reported.
    def wait_ready(self, seconds=3.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < seconds:
            time.sleep(0.02)

    def mark_begin(self):
        self.begin = len(self.lines)

    def mark_end(self):
        self.end = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        window = self.lines[self.begin:self.end] if self.end is not None else self.lines[self.begin:]
        if not window:                       # a timed region shorter than one sampling period: nearest samples
            window = self.lines[max(0, self.begin - 1):(self.end + 1 if self.end is not None else None)]
        for ln in window:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def host_threads():
    """Give the BLAS under NumPy every host core, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would turn the CPU arm into a one-core measurement.  Returns the thread count actually in effect."""
    want = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)                       # stays in effect (no context manager): whole process
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return 1


def reference_model(wl, dropout):
    """(kind, model, step_fn): the UNMODIFIED reference (baseline/_ref, or /root/reference in the build container) when it
    is installed -- kind "reference" -- else the oracle port."""
    L, E, h, F, S, B, V = wl
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from install_reference import reference_path
    ref = reference_path()
    np.random.seed(0)
    if ref is not None:
        sys.dont_write_bytecode = True
        if ref not in sys.path:
            sys.path.insert(0, ref)
        from Transformer.transformer import Transformer as RefTransformer
        from utils.function import clip_grads
        model = RefTransformer(V, E, h, F, L, S, dropout_p=dropout)

        def step(x, y):                                      # transformer.py:635-641 with the clip of utils/train.py:122-129
            logits, cache = model.forward(x)
            loss, dlogits = model.calculate_loss_dlogits(logits, y)
            grads = model.backward(cache, dlogits)
            clip_grads(list(grads.values()), 1.0, lib=model.np)
            model.step(grads, 1e-4)
            return float(loss)
        return "reference", model, step
    from oracle.transformer_oracle import OracleTransformer, full_sequence_step
    model = OracleTransformer(V, E, h, F, L, S, dropout_p=dropout)
    return "port", model, lambda x, y: full_sequence_step(model, x, y, lr=1e-4, clip=1.0)[0]


def cpu_sample_batch(wl):
    """Sequences per CPU step: the bounded sample of the workload (a c4 step at batch 64 would take ~50 s)."""
    L, E, h, F, S, B, V = wl
    return max(1, min(B, 4 if S <= 512 else 1))


def run_reference(args, wl):
    """The reference's own CPU implementation of the step on the host cores, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    L, E, h, F, S, B, V = wl
    sample_B = cpu_sample_batch(wl)
    kind, model, step = reference_model(wl, 0.1)
    rng = np.random.default_rng(1234)
    x = rng.integers(1, V, (sample_B, S)).astype(np.int32)
    y = rng.integers(1, V, (sample_B, S)).astype(np.int32)
    budget = 240.0
    t_start = time.perf_counter()
    for _ in range(args.warmup):
        step(x, y)
        if time.perf_counter() - t_start > budget / 3:
            break
    done, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        step(x, y)
        done += 1
        if time.perf_counter() - t_start > budget:
            break
    dt = time.perf_counter() - t0
    tok_s = done * sample_B * S / dt
    what = ("the unmodified reference (Transformer/transformer.py, NumPy %s, f64 activations under numpy>=2)" % np.__version__
            if kind == "reference" else "oracle port (NumPy, f64 activations like the reference under numpy>=2)")
    sample = "%s, batch %d x seq %d per step (sample of the batch-%d workload), %d steps, %d BLAS threads" % (
        what, sample_B, S, B, done, threads)
    cfg = workload_config(args.workload, wl, 1, "reference CPU path, sampled at batch %d" % sample_B)
    cfg["batch_per_gpu"] = cfg["global_batch"] = sample_B
    cfg["sample_of_batch_per_gpu"] = B
    line = {"impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(name, wl, n_gpus, note):
    L, E, h, F, S, B, V = wl
    return {"workload": "%s: %dL d_model %d heads %d ff %d seq %d vocab %d batch %d/GPU full-sequence LM step (%s)" % (
        name, L, E, h, F, S, V, B, note), "layers": L, "d_model": E, "heads": h, "ff": F, "seq_len": S, "vocab": V,
        "batch_per_gpu": B, "global_batch": B * n_gpus, "parallelism": "dp%d" % n_gpus,
        "l2": "per-step working set (activations, GBs) exceeds the 126 MB L2; no explicit flush"}


def cpu_baseline_sample(wl, seconds=20.0):
    threads = host_threads()
    L, E, h, F, S, B, V = wl
    sample_B = cpu_sample_batch(wl)
    kind, model, step = reference_model(wl, 0.1)
    rng = np.random.default_rng(1234)
    x = rng.integers(1, V, (sample_B, S)).astype(np.int32)
    y = rng.integers(1, V, (sample_B, S)).astype(np.int32)
    t0, done = time.perf_counter(), 0
    while done < 1 or (time.perf_counter() - t0 < seconds and done < 8):
        step(x, y)
        done += 1
    dt = time.perf_counter() - t0
    return {"value": done * sample_B * S / dt, "unit": "tokens/s", "cores": threads, "kind": kind,
            "sample": "%s, batch %d x seq %d (sample of the batch-%d workload), %d steps in %.1f s (f64 activations, OpenBLAS, %d threads)" % (
                "unmodified reference" if kind == "reference" else "oracle port", sample_B, S, B, done, dt, threads)}


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from nlpscratch_b200 import Transformer, _lib
    from nlpscratch_b200.array import DeviceArray

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    distributed = world > 1
    real_stdout = sys.stdout
    if distributed:
        # NCCL's INFO lines (communicator size, rings, NVLS) must be visible so that the driver can verify the rank count,
        # and stdout must stay the one JSON line: NCCL writes to the process's stdout (file descriptor 1), so descriptor 1
        # is pointed at stderr for the whole run and the JSON line is written to a saved copy of the real stdout.
        # Only the INIT subsystem is logged: nothing is printed inside the timed region.
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):     # the box's default is VERSION
            os.environ["NCCL_DEBUG"] = os.environ.get("NLPS_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        sys.stdout.flush()
        real_stdout = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    L, E, h, F, S, B, V = wl
    np.random.seed(1234)
    model = Transformer(V, E, h, F, L, S, dropout_p=args.dropout, precision=args.precision, device=local)
    dp = None
    if distributed:
        from nlpscratch_b200.dp import DataParallel
        dp = DataParallel(model)
        dp.broadcast_parameters(0)
        if dp.backend == "native":
            r_, w_, n_ = C.c_int(), C.c_int(), C.c_int()
            _lib.call("nlps_comm_info", model._h, C.byref(r_), C.byref(w_), C.byref(n_))
            print("[nlps] NCCL communicator owned by libnlps_b200.so: rank %d nranks %d (ncclCommCount), %d bucket exchanges per step"
                  % (r_.value, w_.value, n_.value), file=sys.stderr, flush=True)
    rng = np.random.default_rng(1234 + rank)
    nbatches = 4
    xs = [rng.integers(1, V, (B, S)).astype(np.int32) for _ in range(nbatches)]
    ys = [rng.integers(1, V, (B, S)).astype(np.int32) for _ in range(nbatches)]
    xd = [DeviceArray.from_host(a) for a in xs]
    yd = [DeviceArray.from_host(a) for a in ys]
    lr, clip = 1e-4, 1.0

    def step(i):
        if dp is not None:
            return dp.train_step(xd[i % nbatches], yd[i % nbatches], lr, clip)
        return model.train_step(xd[i % nbatches], yd[i % nbatches], lr, clip)

    def fence():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # before the warm-up: its start-up must not fall into the timed region
    for i in range(args.warmup):
        loss = step(i)
    fence()
    first_loss = float(loss) if args.warmup else None
    if rank == 0:
        sampler.wait_ready()
    model.launch_count(reset=True)
    if dp is not None:
        dp.launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    sampler.mark_begin()
    ev0.record()
    for i in range(args.steps):
        loss = step(i)
    ev1.record()
    fence()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = model.launch_count() + (dp.launches if dp else 0)
    last_loss = float(loss)
    if distributed:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tokens = B * S * world * args.steps
    value = tokens / (ms * 1e-3)
    if args.quick:
        if rank == 0:
            sampler.stop()
            print(json.dumps({"metric": METRIC, "value": value, "ms_per_step": ms / args.steps, "gpu_launches": int(launches),
                              "quick": True, "loss_last": last_loss}), file=real_stdout, flush=True)
        if distributed:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- end to end: pinned host -> device copies of x,y and a loss read-back inside every step ----
    pin = []
    for a in xs + ys:
        p = C.c_void_p()
        _lib.call("nlps_host_alloc_pinned", C.byref(p), C.c_size_t(a.nbytes))
        C.memmove(p.value, a.ctypes.data, a.nbytes)
        pin.append(p)
    x_in, y_in = DeviceArray.empty((B, S), np.int32), DeviceArray.empty((B, S), np.int32)
    x_in.bounds = y_in.bounds = (1, V - 1)
    nbytes = xs[0].nbytes

    # The loss of EVERY step is read back to the host, through a pinned slot and an event, one step behind: step i's copy is
    # enqueued right behind step i, and the host reads step i-1's value while the GPU runs step i (the last one after the
    # loop, inside the timed region) -- the read-back the contract asks for, without idling the GPU on it.
    from nlpscratch_b200 import HostScalarSlot
    slots = [HostScalarSlot(model), HostScalarSlot(model)]
    host_losses = []

    def e2e_step(i, first):
        _lib.call("nlps_memcpy_h2d", _lib.vp(x_in.ptr), pin[i % nbatches], C.c_size_t(nbytes), _lib.stream_ptr())
        _lib.call("nlps_memcpy_h2d", _lib.vp(y_in.ptr), pin[nbatches + i % nbatches], C.c_size_t(nbytes), _lib.stream_ptr())
        l = dp.train_step(x_in, y_in, lr, clip) if dp is not None else model.train_step(x_in, y_in, lr, clip)
        slots[i % 2].fetch(l)
        if not first:
            host_losses.append(slots[(i - 1) % 2].value())       # device -> host read of the previous step's loss

    def e2e_loop(n):
        for i in range(n):
            e2e_step(i, i == 0)
        host_losses.append(slots[(n - 1) % 2].value())

    e2e_loop(min(3, max(1, args.warmup)))
    fence()
    host_losses.clear()
    e_steps = max(10, args.steps)
    t0 = time.perf_counter()
    ev0.record()
    e2e_loop(e_steps)
    ev1.record()
    fence()
    e_ms = ev0.elapsed_time(ev1)
    assert len(host_losses) == e_steps and all(np.isfinite(host_losses)), host_losses
    if distributed:
        t = torch.tensor([e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    e2e_value = B * S * world * e_steps / (e_ms * 1e-3)
    model.synchronize()                                           # surfaces anything the kernels flagged on the way

    # ---- data parallel: every replica must hold the same weights after the timed steps ------------------------------
    replicas_identical = None
    if distributed:
        import hashlib
        model.synchronize()
        digest = hashlib.sha256(model.flat("param").get().tobytes()).hexdigest()
        digests = [None] * world
        dist.all_gather_object(digests, digest)
        replicas_identical = len(set(digests)) == 1

    # ---- roofline of the dominant kernel (the tcgen05 GEMM), timed alone with CUDA events; the denominator is the
    #      tcgen05 kind::tf32 issue ceiling MEASURED in this run (csrc/tf32_peak.cu), one denominator for both fractions ----
    roof = None
    if rank == 0:
        peak = measured_tf32_peak(_lib, local)
        roof = gemm_roofline(torch, _lib, B * S, E, F, args.precision, peak)
        fpt = flops_per_token(L, E, F, S, V)
        ach = fpt * value / world / 1e12
        roof["step"] = {"algorithmic_tflop_per_step_per_gpu": fpt * B * S / 1e12, "achieved_tflops_per_gpu": ach,
                        "frac_of_tf32_peak": ach / peak["tflops"], "tf32_peak_tflops": peak["tflops"],
                        "frac_of_half_bf16_sustained": ach / (measured_peaks()["bf16_sustained"] / 2.0),
                        "peak_note": peak["note"]}
    # ---- other modes / configurations, a few steps each (device-resident), so that the FP32-faithful modes and
    #      BASELINE.json configs[2] / configs[4] get numbers from the same driver-run command -------------------------
    other = None
    if not args.no_other and args.workload == "c4" and args.precision == "tf32":
        model.close()
        model = None
        other = measure_others(args, torch, dist if distributed else None, rank, world, local)
    clocks = sampler.stop() if rank == 0 else None      # after the last GPU-timed leg (see ClockSampler)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(wl)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"tf32": "tf32", "fp32": "fp32", "tf32x3": "tf32x3"}[args.precision],
            "data": "synthetic", "config": workload_config(args.workload, wl, world,
                                                           "dropout %.2f, Adam lr 1e-4, clip 1.0" % args.dropout),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": int(2 * nbytes),
                    "d2h_bytes_per_step": 4, "steps": e_steps, "ms_per_step": e_ms / e_steps,
                    "loss_read": "every step's loss reaches the host inside the timed region: asynchronous copy into a pinned "
                                 "slot behind the step, read one step later (the last one after the loop)"},
            "roofline": roof, "cpu_baseline": cpu, "loss_first": first_loss, "loss_last": last_loss}
    if replicas_identical is not None:
        line["replicas_identical"] = replicas_identical
    if other is not None:
        line["other"] = other
    print(json.dumps(line), file=real_stdout, flush=True)


def measured_tf32_peak(_lib, device):
    """tcgen05 kind::tf32 issue ceiling of THIS GPU, measured now (operands resident in shared memory, accumulators in
    TMEM, no TMA / epilogue / global traffic): the better of the single-CTA (M128 N256) and CTA-pair (M256 N256) forms."""
    best = None
    for group in (1, 2):
        tf, ms = C.c_float(), C.c_float()
        _lib.call("nlps_tf32_mma_peak", C.c_int(device), C.c_int(group), C.c_int(4000), C.c_int(3), C.byref(tf), C.byref(ms))
        if best is None or tf.value > best[0]:
            best = (tf.value, group, ms.value)
    half_bf16 = measured_peaks()
    if not best[0] > 0:                       # should not happen on a B200; keep the line well-formed and say so
        return {"tflops": half_bf16["bf16_burst"] / 2, "cta_group": 0, "ms": 0.0,
                "note": "tcgen05 tf32 probe returned nothing: FALLBACK 1/2 of the %s burst BF16 peak" % half_bf16["source"]}
    return {"tflops": best[0], "cta_group": best[1], "ms": best[2],
            "note": "tcgen05.mma kind::tf32 issue ceiling measured in this run (nlps_tf32_mma_peak, cta_group::%d, %.2f ms "
                    "burst); for comparison 1/2 of the %s BF16 cuBLAS peaks: %.0f burst / %.0f sustained TFLOP/s" % (
                        best[1], best[2], half_bf16["source"], half_bf16["bf16_burst"] / 2, half_bf16["bf16_sustained"] / 2)}


def measure_others(args, torch, dist, rank, world, local, steps=5, warmup=2):
    """{name: {tokens/s, ms_per_step, ...}} for the FP32-faithful arithmetic modes on the headline workload and for the
    other BASELINE.json configurations, device-resident, `steps` timed steps each after `warmup` (graph capture included
    in the warm-up: eager, capture, replay).  Under torchrun every rank runs them (data parallel); loss read once."""
    from nlpscratch_b200 import Transformer
    from nlpscratch_b200.array import DeviceArray
    plan = [("c4/tf32x3", "c4", "tf32x3"), ("c4/fp32", "c4", "fp32"), ("c2/fp32", "c2", "fp32"), ("c2/tf32", "c2", "tf32"),
            ("c3/tf32", "c3", "tf32"), ("c5/tf32", "c5", "tf32"), ("c4-32k/tf32", "c4-32k", "tf32")]
    if world > 1:
        plan = [("c5/tf32", "c5", "tf32")]            # BASELINE.json configs[4]: the long-context model, data parallel
    out = {}
    for name, wname, prec in plan:
        L, E, h, F, S, B, V = WORKLOADS[wname]
        np.random.seed(1234)
        m = Transformer(V, E, h, F, L, S, dropout_p=args.dropout, precision=prec, device=local)
        dp = None
        if dist is not None:
            from nlpscratch_b200.dp import DataParallel
            dp = DataParallel(m)
            dp.broadcast_parameters(0)
        rng = np.random.default_rng(99 + rank)
        xd = [DeviceArray.from_host(rng.integers(1, V, (B, S)).astype(np.int32)) for _ in range(2)]
        yd = [DeviceArray.from_host(rng.integers(1, V, (B, S)).astype(np.int32)) for _ in range(2)]
        run = (lambda i: dp.train_step(xd[i % 2], yd[i % 2], 1e-4, 1.0)) if dp else (lambda i: m.train_step(xd[i % 2], yd[i % 2], 1e-4, 1.0))
        n_warm = warmup + 2                           # eager, capture, then replays
        for i in range(n_warm):
            loss = run(i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = run(n_warm + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        fpt = flops_per_token(L, E, F, S, V)
        tok_s = B * S * world * steps / (ms * 1e-3)
        out[name] = {"value": tok_s, "unit": "tokens/s", "ms_per_step": ms / steps, "steps": steps, "warmup": n_warm,
                     "precision": prec, "achieved_tflops_per_gpu": fpt * tok_s / world / 1e12, "loss": float(loss),
                     "config": "%dL d%d h%d ff%d seq%d vocab%d batch%d/GPU dropout %.2f" % (L, E, h, F, S, V, B, args.dropout)}
        m.close()
        del m, dp, xd, yd
    if world == 1:
        out["c1/tf32"] = driver_epoch(args, torch)
    return out


def driver_epoch(args, torch):
    """BASELINE.json configs[0] on the device: one epoch of the driver loop (nlpscratch_b200.train.train_transformer: fused
    last-token steps, device-side evaluation) over the reference's mini dataset, after one warm-up epoch."""
    from nlpscratch_b200 import Transformer
    from nlpscratch_b200.train import train_transformer
    L, E, h, F, S, B, V = WORKLOADS["c1"]
    X, Y, data_name = mini_dataset(V, S)
    np.random.seed(0)
    m = Transformer(V, E, h, F, L, S, dropout_p=args.dropout, precision="tf32")
    quiet = lambda *_a, **_k: None
    train_transformer(m, X, Y, learning_rate=1e-4, nepoch=1, evaluation_loss_after=1, batch_size=B, print_every=0,
                      clip_grad_norm=1.0, log=quiet)
    torch.cuda.synchronize()
    stats = train_transformer(m, X, Y, learning_rate=1e-4, nepoch=1, evaluation_loss_after=1, batch_size=B, print_every=0,
                              clip_grad_norm=1.0, log=quiet)
    m.close()
    return {"value": stats["tokens"] / stats["train_seconds"], "unit": "non-pad tokens/s", "ms_per_step": 1e3 * stats["train_seconds"] / stats["steps"],
            "steps": stats["steps"], "precision": "tf32", "eval_loss_before_epoch": stats["losses"][-1][1],
            "config": "driver loop, last-token loss, %s, %d prefixes x max_len %d, B%d, 3L d256 h8 ff1024 V%d" % (data_name, len(Y), S, B, V)}


def gemm_roofline(torch, _lib, N, E, F, precision, peak_info):
    """FFN up-projection GEMM (N,E)x(E,F), the most frequent shape of the step, timed alone."""
    from nlpscratch_b200.array import DeviceArray, b200np
    prec = _lib.PREC[precision]
    a = DeviceArray.from_host(np.random.default_rng(0).standard_normal((N, E)).astype(np.float32))
    b = DeviceArray.from_host(np.random.default_rng(1).standard_normal((E, F)).astype(np.float32))
    c = DeviceArray.empty((N, F), np.float32)
    stream = C.c_void_p(0)

    def launch():
        _lib.call("nlps_gemm", stream, C.c_int(prec), C.c_int(N), C.c_int(F), C.c_int(E), _lib.vp(a.ptr),
                  C.c_int64(E), C.c_int(0), _lib.vp(b.ptr), C.c_int64(F), C.c_int(0), _lib.vp(c.ptr), C.c_int64(F),
                  None, C.c_int(0), None, C.c_int64(0), None, C.c_int64(0), C.c_int(0))
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * N * E * F
    peak = peak_info["tflops"]
    ach = flops / (ms * 1e-3) / 1e12
    traffic, src = None, None
    for cand in ("r02_gemm_ffn1_ncu.json", "r01_gemm_ffn1_ncu.json"):
        try:    # DRAM bytes of the same launch: NOT measured in this run (that needs ncu) -- from the committed capture
            with open(os.path.join(ROOT, "profiles", cand)) as fh:
                m = json.load(fh)
            if (N, E, F) == (32768, 512, 2048):
                traffic = (float(m["dram__bytes_read.sum"]["value"]) + float(m["dram__bytes_write.sum"]["value"])) * 1e6
                src = cand
                break
        except Exception:
            continue
    return {"bound": "tensor", "kernel": "gemm_tf32_pair_kernel (tcgen05 cta_group::2 kind::tf32) %dx%dx%d" % (N, F, E),
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
            "traffic_note": "NOT from this run: dram__bytes_read+write of this launch in the committed `ncu --set full` capture "
                            "profiles/%s; algorithmic bytes %.1f MB" % (src, (N * E + E * F + N * F) * 4 / 1e6),
            "ms_per_launch": ms, "peak_note": peak_info["note"]}


def synthetic_prefix_dataset(V, max_len, n_sent=69, seed=0):
    """Every prefix of every sentence, right-padded (transformer_train.py:47-60), on synthetic sentences
    with the mini corpus' statistics (69 sentences -> ~1.4k prefixes, ~20k non-PAD tokens)."""
    rng = np.random.default_rng(seed)
    X, Y = [], []
    for _ in range(n_sent):
        n = int(rng.integers(8, 36))
        ids = [2] + [int(v) for v in rng.integers(3, V, n)]
        for i in range(1, len(ids)):
            X.append(ids[:i][-max_len:] + [0] * max(0, max_len - i))
            Y.append(ids[i])
    return np.array(X, dtype=np.int32), np.array(Y, dtype=np.int32)


def mini_dataset(V, S):
    """The reference's own dataset for BASELINE.json configs[0]: the word-level prefix dataset of
    dataset/tagged_train_mini.txt (1434 prefixes x max_len 320, V 574; integer arrays committed as
    tests/golden/dataset_mini.npz); synthetic sentences of the same statistics only if that file is missing."""
    path = os.path.join(ROOT, "tests", "golden", "dataset_mini.npz")
    if os.path.exists(path):
        with np.load(path) as z:
            if int(z["vocab"]) == V and z["X"].shape[1] == S:
                return z["X"], z["Y"], "dataset/tagged_train_mini.txt (word level)"
    X, Y = synthetic_prefix_dataset(V, S)
    return X, Y, "synthetic sentences with the mini corpus' statistics"


def run_driver(args, wl):
    """Workload c1: the reference driver loop (last-token loss, B16, clip 1.0, Adam 1e-4) on its own dataset."""
    import torch
    from nlpscratch_b200 import Transformer
    from nlpscratch_b200.train import train_transformer
    from oracle.transformer_oracle import OracleTransformer, last_token_step
    L, E, h, F, S, B, V = wl
    X, Y, data_name = mini_dataset(V, S)
    nonpad = int((X != 0).sum())
    if args.impl == "reference":
        np.random.seed(0)
        oracle = OracleTransformer(V, E, h, F, L, S, dropout_p=0.1)
        t0, tok, steps = time.perf_counter(), 0, 0
        for i in range(0, len(Y), B):
            last_token_step(oracle, X[i:i + B], Y[i:i + B], lr=1e-4, clip=1.0)
            tok += int((X[i:i + B] != 0).sum()); steps += 1
            if time.perf_counter() - t0 > 60:
                break
        dt = time.perf_counter() - t0
        line = {"impl": "reference", "metric": METRIC, "value": tok / dt, "unit": "non-pad tokens/s", "n_gpus": 1,
                "steps": steps, "warmup": 0, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "c1: reference driver loop, last-token loss, %d prefixes x max_len %d, B%d" % (len(Y), S, B)},
                "cpu_baseline": {"value": tok / dt, "unit": "non-pad tokens/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": "%d batches of the epoch" % steps},
                "e2e": {"value": tok / dt, "unit": "non-pad tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    np.random.seed(0)
    model = Transformer(V, E, h, F, L, S, dropout_p=args.dropout, precision=args.precision)
    quiet = lambda *_a, **_k: None
    train_transformer(model, X, Y, learning_rate=1e-4, nepoch=1, evaluation_loss_after=1, batch_size=B,
                      print_every=0, clip_grad_norm=1.0, log=quiet)          # warm-up epoch
    model.launch_count(reset=True)
    torch.cuda.synchronize()
    stats = train_transformer(model, X, Y, learning_rate=1e-4, nepoch=max(1, args.steps // 10), evaluation_loss_after=10 ** 9,
                              batch_size=B, print_every=0, clip_grad_norm=1.0, log=quiet)
    tok_s = stats["tokens"] / stats["train_seconds"]
    line = {"metric": METRIC, "value": tok_s, "unit": "non-pad tokens/s", "n_gpus": 1, "steps": stats["steps"], "warmup": 90,
            "ms_per_step": 1e3 * stats["train_seconds"] / stats["steps"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": data_name,
            "config": {"workload": "c1: reference driver loop (nlpscratch_b200.train.train_transformer), last-token loss, "
                                   "%d prefixes x max_len %d (%d non-PAD tokens/epoch), B%d, host->device copy of the dataset "
                                   "and shuffles inside the timed region" % (len(Y), S, nonpad, B)},
            "gpu_launches": model.launch_count(), "graph_steps_eager_captured_replayed": list(model.graph_stats()),
            "e2e": {"value": tok_s, "unit": "non-pad tokens/s", "h2d_bytes_per_step": int(X.nbytes / max(1, stats["steps"])),
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32", "tf32x3"])
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the secondary modes / configurations (\"other\" key)")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (for ncu launch lists)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.workload == "c1":
        if int(os.environ.get("RANK", "0")) == 0:
            run_driver(args, wl)
        return
    if args.impl == "reference":
        run_reference(args, wl)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, wl)


if __name__ == "__main__":
    main()
