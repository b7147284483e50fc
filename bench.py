#!/usr/bin/env python
"""Benchmark of the LEAStereo hot path on B200 - BASELINE.json's metric: stereo pairs/sec at KITTI 384x1248, D=192.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = one forward of ``LEAStereo(left, right)`` over a batch of B synthetic stereo pairs per GPU (native 2D
feature net + the CUDA hot path: cost volume -> 3D matching net -> disparity head; B = 8 by default - throughput
saturates from 4 pairs on, BASELINE configs[2] sweeps batch 1-64).  Pairs are independent, so N GPUs each run their own batch with no
data-path collective (weak scaling); the only collectives of the inference arm are the timing barrier and the
max-over-ranks of the device time.

Printed JSON line (rank 0): the task contract's keys, plus
  roofline      the 3D-conv kernel family: algorithmic FLOPs (sum 2*M*N*K over the reference's 102 convs, SURVEY.md
                8d) / summed CUDA-event durations of the launches that execute them, against the measured bf16 peak
                (``frac``), next to the same figure over the FLOPs actually launched (``frac_as_launched``);
  kernels       per-kernel-class times; cost-volume (standalone) and disparity-head entries against the HBM peak and,
                for the head, the exp-throughput bound;
  e2e           same metric through the public API with HOST buffers: 8-bit image pairs in pinned memory ->
                ``pipeline.InputPipeline`` (upload, z-normalisation, pad: predict.py:144-184) -> ``LEAStereo`` ->
                disparity maps back in pinned host memory, all inside the timed region;
  parity        the timed configuration (same engine options) against the reference's own fp32 output on the KITTI
                fixture tests/golden/large_cal_384x1248_d192.npz (north-star tolerance);
  gpu_baseline  the oracle port - the reference's op sequence on stock PyTorch / cuDNN - on the SAME GPU and batch,
                with ``allow_tf32`` off and on, each with its own tolerance result (SURVEY.md 0.1 "practical bar");
  cpu_baseline  the oracle (CPU port of the reference path) timed on the host cores on a bounded sample;
  train         BASELINE configs[4] (288x576, batch 4 per GPU): fwd + bwd + NCCL gradient all-reduce + Adam, at every N.
``--impl reference`` times the reference's own CPU implementation of the path (the oracle port - the Python reference
cannot travel to the GPU box) on rank 0.
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"name": "KITTI2015 384x1248 maxdisp=192 inference (BASELINE.json configs[2])", "H": 384, "W": 1248,
            "maxdisp": 192}
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"]))}, "measured"
        except Exception:  # noqa: BLE001
            pass
    return dict(FALLBACK_PEAKS), "fallback"


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML every 10 ms; nvidia-smi fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # NVML clocks-event (throttle) reason bits
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index, self._stop, self._t = index, threading.Event(), None
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _run(self):
        while not self._stop.is_set():
            try:
                if self.nvml is not None:
                    n = self.nvml
                    self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                    try:
                        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:  # noqa: BLE001
                        r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for name, bit in self.BITS.items():
                        if r & bit:
                            self.reasons.add(name)
                    self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                    self._stop.wait(0.01)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.sm.append(float(f[0])); self.max_mhz = float(f[1])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": round(max(self.power), 1) if self.power else None,
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------------
def build_model(maxdisp, device, options):
    from leastereo_b200 import LEAStereo, default_args
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = LEAStereo(default_args(maxdisp=maxdisp, cuda=True), device)
    model = model.to(device).eval()
    model.engine_options = dict(options)
    return model


def synthetic_pairs(B, H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)


def oracle_state_dict():
    """Random-init weights of the architecture (seed 0), for the CPU arms."""
    from leastereo_b200 import LEAStereo, default_args
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = LEAStereo(default_args(maxdisp=WORKLOAD["maxdisp"], cuda=False), "cpu")
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def time_oracle(rows, steps, warmup):
    """Times oracle.leastereo_forward on a band of `rows` image rows of the KITTI pair; returns s/step list."""
    from oracle import leastereo_oracle as O
    sd = oracle_state_dict()
    left, right = synthetic_pairs(1, rows, WORKLOAD["W"])
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.leastereo_forward(sd, left, right, WORKLOAD["maxdisp"])
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    return ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H = WORKLOAD["H"]
    # bounded sample: calibrate on a 96-row band; use the full pair only if the whole run stays within ~4 minutes
    t_band = time_oracle(96, 1, 0)[0]
    est_full = t_band * (H / 96.0) * (args.steps + args.warmup)
    rows = H if est_full <= 240.0 else 96
    frac = rows / float(H)
    ts = time_oracle(rows, args.steps, args.warmup)
    total = sum(ts)
    value = frac * len(ts) / total
    sample = "%d-row band of one 384x1248 pair per step (%.2f pair), D=192, fp32 PyTorch-CPU oracle port" % (rows, frac)
    line = {"impl": "reference", "metric": "stereo pairs/sec, KITTI 384x1248 D=192", "value": value, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
def profile_kernels(model, left, right):
    """Eager pass with a CUDA-event pair around every launch of the hot path; returns per-kind aggregates."""
    from leastereo_b200 import engine
    from leastereo_b200.kernels import get_ops
    ops = get_ops()
    opt = engine._options(model)
    B, _, H, W = left.shape
    H3, W3, C = (H - 1) // 3 + 1, (W - 1) // 3 + 1, model.matching.initial_fm
    D3 = int(model.maxdisp / 3)
    plan = engine.get_plan(model.matching, B, (D3, H3, W3), left.device, opt, ops)
    plan.refresh_params()
    fplan = next((p for p in engine._plans(model.feature).values() if isinstance(p, engine.FeaturePlan)), None)
    if fplan is None:
        with torch.no_grad():
            fx, fy = model.extract_features(left, right)
        fx, fy = fx.float().contiguous(), fy.float().contiguous()
    stream = torch.cuda.current_stream()
    records = []

    def timed(name, kind, fn, flops=0.0, nbytes=0.0):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        records.append([name, kind, a, b, flops, nbytes])

    for rep in range(2):       # first repetition warms up, second is reported
        records.clear()
        torch.cuda.synchronize()
        # let the host run ahead of the GPU: with an idle GPU a small kernel's event pair would time the host launch
        # path (~20 us per ctypes call) instead of the kernel
        torch.cuda._sleep(int(2.5e7))
        cv_bytes = 4.0 * B * (2 * C * H3 * W3 + 2 * C * D3 * H3 * W3)
        if fplan is not None:
            fplan.img[:B].copy_(left); fplan.img[B:].copy_(right)
            for s in fplan.steps:
                timed(s.name, "feature_" + s.kind, lambda s=s: fplan.run_step(s), s.flops, s.bytes)
        elif plan.fxp is not None:
            pk_bytes = 4.0 * B * 2 * C * H3 * W3 * 2
            timed("pack_features", "pack_features",
                  lambda: (ops.pack(fx, opt["planes"], out=plan.fxp), ops.pack(fy, opt["planes"], out=plan.fyp)), 0.0,
                  pk_bytes)
        else:
            timed("cost_volume", "cost_volume", lambda: ops.cost_volume_planes(fx, fy, model.maxdisp, opt["planes"],
                                                                              out=plan.cost), 0.0, cv_bytes)
        for s in plan.steps:
            timed(s.name, s.kind, lambda s=s: plan.run_step(s), s.flops, s.bytes)
        dh_bytes = 4.0 * B * (D3 * H3 * W3 + 9 * H3 * W3)
        timed("disp_head", "disp_head", lambda: ops.disp_head(plan.mat, model.maxdisp), 0.0, dh_bytes)
        torch.cuda.synchronize()
    agg = {}
    for name, kind, a, b, flops, nbytes in records:
        d = agg.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["launches"] += 1; d["ms"] += a.elapsed_time(b); d["flops"] += flops; d["bytes"] += nbytes
    per_launch = [(name, kind, a.elapsed_time(b), flops, nbytes) for name, kind, a, b, flops, nbytes in records]
    agg["_reference_conv_flops"] = plan.reference_conv_flops()
    return agg, per_launch


# ------------------------------------------------------------------------------------------------------------
# legs that stand beside the headline number (rank 0 unless stated)
# ------------------------------------------------------------------------------------------------------------
KITTI_FIXTURE = os.path.join(ROOT, "tests", "golden", "large_cal_384x1248_d192.npz")


def load_kitti_fixture():
    """The reference's own fp32 run of KITTI 384x1248 D=192 (calibrated BN statistics, seeded inputs / weights):
    tests/golden/make_golden_large.py.  Returns (state-dict overrides, left, right, reference disparity) for pair 0."""
    import numpy as np
    z = np.load(KITTI_FIXTURE)
    bn = {k[3:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("bn/")}
    g = torch.Generator().manual_seed(1)
    left = torch.randn(1, 3, WORKLOAD["H"], WORKLOAD["W"], generator=g)
    right = torch.randn(1, 3, WORKLOAD["H"], WORKLOAD["W"], generator=g)
    return bn, left, right, torch.from_numpy(z["disp0"].copy())


def tolerance(d, ref):
    diff = (d.double().cpu() - ref.double()).abs()
    frac, mean = float((diff <= 0.1).double().mean()), float(diff.mean())
    return {"frac_within_0p1px": round(frac, 6), "mean_abs_px": round(mean, 6), "max_abs_px": round(float(diff.max()), 4),
            "ok": bool(frac >= 0.999 and mean <= 0.01)}


def parity_leg(device, options):
    """The TIMED configuration (same engine options) on the KITTI fixture against the reference's fp32 disparity."""
    if not os.path.exists(KITTI_FIXTURE):
        return {"unavailable": "tests/golden/large_cal_384x1248_d192.npz missing"}
    bn, left, right, ref = load_kitti_fixture()
    model = build_model(WORKLOAD["maxdisp"], device, options)
    sd = model.state_dict()
    sd.update({k: v.to(device) for k, v in bn.items()})
    model.load_state_dict(sd)
    model.engine_options = dict(options)
    with torch.no_grad():
        d = model(left.to(device), right.to(device))
    out = tolerance(d, ref)
    out.update(fixture="tests/golden/large_cal_384x1248_d192.npz (unmodified reference, fp32 CPU, calibrated regime of SURVEY 8d)",
               tolerance="|dd| <= 0.1 px on >= 99.9 % of pixels and mean |dd| <= 0.01 px (BASELINE.json north_star)")
    del model
    torch.cuda.empty_cache()
    return out


def gpu_baseline_leg(device, B, steps=3, warmup=2):
    """Stock PyTorch + cuDNN on the same B200: the oracle port (the reference's op sequence, functional, fp32 tensors)
    run on ``cuda`` for the same KITTI batch, with ``allow_tf32`` off and on (SURVEY.md 0.1 / BASELINE.md 4.5), each
    with its tolerance against the reference's fp32 CPU output on the KITTI fixture."""
    from oracle import leastereo_oracle as O
    H, W, maxdisp = WORKLOAD["H"], WORKLOAD["W"], WORKLOAD["maxdisp"]
    sd = {k: v.to(device) for k, v in oracle_state_dict().items()}
    left, right = synthetic_pairs(B, H, W, seed=1)
    left, right = left.to(device), right.to(device)
    fixture = load_kitti_fixture() if os.path.exists(KITTI_FIXTURE) else None
    out = {"what": "oracle port (reference op sequence: F.conv3d / batch_norm / interpolate / softmin) on cuda, "
                   "torch %s, cudnn.benchmark=True, batch %d" % (torch.__version__, B), "unit": "pairs/s"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.benchmark = True
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            try:
                with torch.no_grad():
                    for _ in range(warmup):
                        O.leastereo_forward(sd, left, right, maxdisp)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        O.leastereo_forward(sd, left, right, maxdisp)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / steps
                    entry = {"value": round(B / (ms * 1e-3), 3), "ms_per_step": round(ms, 3), "allow_tf32": tf32}
                    if fixture is not None:
                        bn, fl, fr, ref = fixture
                        sdc = dict(sd)
                        sdc.update({k: v.to(device) for k, v in bn.items()})
                        entry["tolerance_vs_reference_fp32"] = tolerance(
                            O.leastereo_forward(sdc, fl.to(device), fr.to(device), maxdisp), ref)
                out[name] = entry
            except Exception as e:  # noqa: BLE001
                out[name] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    return out


def train_leg(device, rank, world, dist, steps=4, warmup=5, batch=4, H=288, W=576, maxdisp=192):
    """BASELINE configs[4]: fwd + bwd + gradient all-reduce (ONE NCCL all-reduce of the flat fp32 bucket over NVLink)
    + Adam at 288x576, batch 4 per GPU, every rank; device time, max over ranks (train.py:153-160)."""
    from leastereo_b200 import LEAStereo, default_args
    from leastereo_b200.pipeline import FlatAdam, masked_smooth_l1_loss
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = LEAStereo(default_args(maxdisp=maxdisp, cuda=True), device).to(device).train()
    opt = FlatAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(101 + rank)
    left = torch.randn(batch, 3, H, W, generator=g).to(device)
    right = torch.randn(batch, 3, H, W, generator=g).to(device)
    target = (torch.rand(batch, H, W, generator=g) * maxdisp * 0.5).to(device)
    ar = [torch.cuda.Event(enable_timing=True) for _ in range(2 * (steps + warmup))]
    losses = []

    def step(k):
        opt.zero_grad()
        disp = model(left, right)
        loss = masked_smooth_l1_loss(disp, target, maxdisp)
        loss.backward()
        ar[2 * k].record()
        opt.allreduce(world)
        ar[2 * k + 1].record()
        opt.step()
        losses.append(loss.detach())

    for k in range(warmup):
        step(k)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(warmup, warmup + steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1) / steps
    ar_us = 1e3 * sum(ar[2 * k].elapsed_time(ar[2 * k + 1]) for k in range(warmup, warmup + steps)) / steps
    t = torch.tensor([ms, ar_us], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {"workload": "BASELINE.json configs[4]: train fwd+bwd+all-reduce+Adam, SceneFlow crop %dx%d D=%d, batch %d per GPU"
                       % (H, W, maxdisp, batch),
           "ms_per_step": round(float(t[0]), 3), "value": round(batch * world / (float(t[0]) * 1e-3), 3), "unit": "pairs/s",
           "n_gpus": world, "steps": steps, "warmup": warmup, "scaling": "weak",
           "allreduce_us_per_step": round(float(t[1]), 1) if world > 1 else 0.0,
           "allreduce_bytes": int(opt.grad.numel() * 4), "nccl_ranks": world if world > 1 else 0,
           "loss_first_last": [round(float(losses[0]), 5), round(float(losses[-1]), 5)],
           "max_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    del model, opt
    from leastereo_b200 import training
    training._TRAIN_PLANS.clear()
    torch.cuda.empty_cache()
    return out


def standalone_kernels(device, B, peaks):
    """The two HBM-roofline kernels of BASELINE.json's north_star, timed on their own (CUDA events, best of 5 after a
    warm-up; inputs larger than L2 or rewritten between runs): the standalone bit-exact cost volume and the fused
    disparity head."""
    from leastereo_b200.kernels import get_ops
    ops = get_ops()
    H3, W3, C, D3, maxdisp = 128, 416, 32, 64, WORKLOAD["maxdisp"]
    g = torch.Generator().manual_seed(3)
    fx = torch.randn(B, C, H3, W3, generator=g).to(device)
    fy = torch.randn(B, C, H3, W3, generator=g).to(device)
    mat = (torch.randn(B, 1, D3, H3, W3, generator=g) * 3).to(device)

    def best(fn, n=5):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts)

    out = {}
    cv_ms = best(lambda: ops.cost_volume_f32(fx, fy, maxdisp))
    cv_bytes = 4.0 * B * (2 * C * H3 * W3 + 2 * C * D3 * H3 * W3)
    out["cost_volume_standalone"] = {"ms": round(cv_ms, 4), "algorithmic_bytes": cv_bytes,
                                     "GBps_algorithmic": round(cv_bytes / (cv_ms * 1e-3) / 1e9, 1),
                                     "frac_hbm_peak": round(cv_bytes / (cv_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                                     "note": "lea_cost_volume_f32 (reference layout, bit-exact; includes the allocation of "
                                             "its output by torch); on the forward path the volume is never built"}
    dh_ms = best(lambda: ops.disp_head(mat, maxdisp))
    dh_bytes = 4.0 * B * (D3 * H3 * W3 + 9 * H3 * W3)
    exps = float(B) * maxdisp * 9 * H3 * W3
    sm_count = torch.cuda.get_device_properties(device).multi_processor_count
    mufu_peak = sm_count * 16 * 1.965e9            # 16 exp / clk / SM (sm_90 rate, assumed) at the max SM clock
    out["disp_head_standalone"] = {"ms": round(dh_ms, 4), "algorithmic_bytes": dh_bytes,
                                   "GBps_algorithmic": round(dh_bytes / (dh_ms * 1e-3) / 1e9, 1),
                                   "frac_hbm_peak": round(dh_bytes / (dh_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                                   "algorithmic_exps": exps, "Gexp_per_s": round(exps / (dh_ms * 1e-3) / 1e9, 1),
                                   "frac_exp_peak": round(exps / (dh_ms * 1e-3) / mufu_peak, 4),
                                   "note": "fused head: 192 soft-min terms per output pixel make it exp/FMA-bound, not "
                                           "HBM-bound (SURVEY 8d: <= ~12 % of the HBM roofline is reachable)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8,
                    help="stereo pairs per GPU per step (BASELINE configs[2] sweeps batch 1-64; throughput saturates from 4 on, "
                         "8 keeps the timed region of a 20-step run near one second)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--conv", default=os.environ.get("LEA_CONV", "tc"), choices=["tc", "simt"])
    ap.add_argument("--planes", type=int, default=int(os.environ.get("LEA_PLANES", "2")))
    ap.add_argument("--mma-terms", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--quick", action="store_true", help="development: only the headline arm (no baselines / train / parity)")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch profile to this JSON file")
    ap.add_argument("--knob", action="append", default=[], help="engine option override name=int (development)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.quick:
        args.no_cpu_baseline = args.no_gpu_baseline = args.no_train = args.no_parity = True

    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    import __graft_entry__ as entry
    if rank == 0:
        with contextlib.redirect_stdout(sys.stderr):
            entry.build()
    if dist is not None:
        dist.barrier()
    from leastereo_b200 import engine
    from leastereo_b200.kernels import get_ops
    from leastereo_b200.pipeline import InputPipeline
    ops = get_ops()

    H, W, maxdisp, B = WORKLOAD["H"], WORKLOAD["W"], WORKLOAD["maxdisp"], args.batch
    options = {"planes": args.planes, "conv": args.conv, "mma_terms": args.mma_terms, "assume_frozen": True,
               "cuda_graph": not args.no_graph}
    for kv in args.knob:
        k, v = kv.split("=")
        options[k] = int(v)
    model = build_model(maxdisp, device, options)
    left_h, right_h = synthetic_pairs(B, H, W, seed=1 + rank)
    left, right = left_h.to(device), right_h.to(device)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        model(left, right)                       # builds the plan, packs weights
        torch.cuda.synchronize()
        n0 = ops.launches
        disp = model(left, right)
        torch.cuda.synchronize()
        launches_per_step = ops.launches - n0

        # Both timed arms call the public API, ``LEAStereo.forward``: from its third call with one input shape the
        # module replays its own CUDA graph of the launch list (engine option ``cuda_graph``; ``--no-graph`` turns it
        # off for per-launch profiles).  The bench only checks that the graph exists and reproduces the eager result.
        graph_on = False
        if not args.no_graph:
            out3 = model(left, right)
            torch.cuda.synchronize()
            fplans = [q for q in engine._plans(model.feature).values() if isinstance(q, engine.FeaturePlan)]
            graph_on = bool(fplans) and fplans[-1].__dict__.get("_graph", {}).get("graph") is not None
            if not graph_on:
                print("[bench] the module did not capture its CUDA graph; running eagerly", file=sys.stderr)
            elif not torch.equal(out3, disp):
                raise RuntimeError("graph replay differs from the eager run")

        def step_device():
            return model(left, right)

        # ---- device-resident throughput ----
        for _ in range(args.warmup):
            step_device()
        barrier()
        sampler = ClockSampler(local_rank)
        with sampler:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(args.steps):
                step_device()
            ev1.record()
            barrier()
        ms_total = ev0.elapsed_time(ev1)

        # ---- end to end through the public API, HOST buffers on both sides ----
        # What predict.py does per pair (PIL image -> z-normalise -> pad -> upload -> model -> download), with the
        # repo's public pieces: 8-bit images sit in pinned host memory; every step ``InputPipeline.submit`` copies the
        # step's B pairs to the device on a side stream (2*H*W*3 bytes per pair), ``next`` normalises + pads them
        # (predict.py:144-184 as kernels) straight into the model's input tensors, the model runs, and the B disparity
        # maps are copied into pinned host memory.  Upload of step k+1 and download of step k overlap step k's kernels.
        import numpy as np
        rng = np.random.RandomState(7 + rank)
        Hi, Wi = 375, 1242                                    # a KITTI 2015 frame; padded to 384 x 1248 (predict_kitti15.sh)
        host_sets = [(torch.from_numpy(rng.randint(0, 256, (B, Hi, Wi, 3)).astype(np.uint8)),
                      torch.from_numpy(rng.randint(0, 256, (B, Hi, Wi, 3)).astype(np.uint8))) for _ in range(2)]
        pipe = InputPipeline(Hi, Wi, H, W, device, depth=2, batch=B)
        cs = torch.cuda.current_stream()
        xs = torch.cuda.Stream()
        out_stage = [torch.empty((B, H, W), dtype=torch.float32, device=device) for _ in range(2)]
        out_hs = [torch.empty((B, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        ev_d2h = [torch.cuda.Event() for _ in range(2)]

        def run_e2e(n):
            pipe.submit(*host_sets[0])
            for k in range(n):
                if k + 1 < n:
                    pipe.submit(*host_sets[(k + 1) % 2])     # next step's 8-bit upload overlaps this step's kernels
                pipe.next(dst=(left, right))                  # normalise + pad into the model's (graph's) input tensors
                o = step_device()
                s_ = k % 2
                if k >= 2:
                    cs.wait_event(ev_d2h[s_])                 # the slot's previous download has left the device buffer
                out_stage[s_].copy_(o)
                ev_out[s_].record(cs)
                with torch.cuda.stream(xs):
                    xs.wait_event(ev_out[s_])
                    out_hs[s_].copy_(out_stage[s_], non_blocking=True)
                    ev_d2h[s_].record(xs)
            cs.wait_stream(xs)

        run_e2e(2)
        barrier()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record()
        run_e2e(args.steps)
        ev3.record()
        barrier()
        ms_e2e = ev2.elapsed_time(ev3)
        # the last step's host result must be the model's output for that step's normalised images
        check = step_device().cpu()
        if not torch.allclose(out_hs[(args.steps - 1) % 2], check, atol=1e-3, rtol=0) or not torch.isfinite(check).all():
            raise RuntimeError("end-to-end result differs from the device-resident result")
        h2d_per_step, d2h_per_step = pipe.h2d_bytes_per_batch, B * H * W * 4
        left.copy_(left_h); right.copy_(right_h)               # restore the seeded fp32 inputs for the profile below

        t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(t[0]), float(t[1])

        agg, per_launch = (None, None)
        if rank == 0:
            agg, per_launch = profile_kernels(model, left, right)
        opt_view = engine._options(model)
        plan = next(iter(engine._plans(model.matching).values()))
        plan_facts = {"collapse_stem0": plan.fxy3 is not None, "fuse_cv": plan.fxp is not None,
                      "fuse_head": any(s.kind == "head_taps" for s in plan.steps),
                      "feature_planes": int(opt_view.get("feature_planes", 3)),
                      "accum_split": int(opt_view.get("accum_split", 0)),
                      "arena_GB": round(getattr(plan, "arena_bytes", 0) / 2 ** 30, 2)}

    # ---- legs beside the headline ----
    engine.invalidate_cached_plans(model)
    del model, plan
    torch.cuda.empty_cache()
    train = None
    if not args.no_train:
        try:
            train = train_leg(device, rank, world, dist)
        except Exception as e:  # noqa: BLE001
            train = {"error": str(e)[:300]}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks, peak_src = load_peaks()
    pairs = B * world * args.steps
    value = pairs / (ms_total / 1e3)
    e2e_value = pairs / (ms_e2e / 1e3)
    ref_conv_flops = agg.pop("_reference_conv_flops")
    conv_kinds = [k for k in agg if k.startswith("conv")]        # 3D matching-net convs only ("feature_*" excluded)
    # everything that executes the reference's 3D convs: the conv launches plus the two kernels that finish the
    # algebraically rewritten ones (collapsed stem0, low-res head contraction)
    conv_time_kinds = conv_kinds + [k for k in ("stem0_assemble", "head_taps") if k in agg]
    # the 1x1x1 convs fused into the down-sampling resample launches (HBM-bound, CUDA cores, < 0.2 % of the FLOPs) are
    # left out of both the numerator and the denominator
    ref_conv_flops -= agg.get("resample_conv1x1", {}).get("flops", 0.0)
    conv_ms = sum(agg[k]["ms"] for k in conv_time_kinds)
    conv_flops = sum(agg[k]["flops"] for k in conv_kinds)
    conv_tflops = ref_conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    launched_tflops = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    step_ms_profile = sum(v["ms"] for v in agg.values())
    kernels = {}
    for kind, v in agg.items():
        e = {"launches": v["launches"], "ms": round(v["ms"], 4), "share_of_hot_path": round(v["ms"] / step_ms_profile, 4)}
        if v["flops"] > 0:
            e["TFLOPs"] = round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2)
        if v["bytes"] > 0 and v["ms"] > 0:
            e["GBps_algorithmic"] = round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)
            e["frac_hbm_peak"] = round(e["GBps_algorithmic"] / peaks["hbm_gbs"], 4)
        kernels[kind] = e
    try:
        kernels.update(standalone_kernels(device, B, peaks))
    except Exception as e:  # noqa: BLE001
        kernels["standalone_error"] = str(e)[:200]
    # DRAM traffic of the dominant kernel cannot be measured outside a profiler: it is the figure of this round's
    # ncu capture of this very command (profiles/r02_conv_tc_traffic.json), named as such - or null.
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_conv_tc_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/r02_conv_tc_traffic.json: " + tj.get("captured", "ncu")
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "tensor", "kernel": "lea_conv_tc_kernel" if args.conv == "tc" else "lea_conv3_simt_kernel",
                "achieved": round(conv_tflops, 2), "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": round(conv_tflops / peaks["bf16_tflops_sustained"], 4),
                "achieved_as_launched": round(launched_tflops, 2),
                "frac_as_launched": round(launched_tflops / peaks["bf16_tflops_sustained"], 4),
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch_1pair": int(sum(agg[k]["bytes"] for k in conv_kinds) / max(1, B) /
                                                          max(1, sum(agg[k]["launches"] for k in conv_kinds))),
                "peak_source": "%s bf16 dense, sustained (kernel timed inside a long step)" % peak_src,
                "algorithmic_flops_per_step": ref_conv_flops, "conv_launches_per_step": sum(agg[k]["launches"] for k in conv_kinds),
                "launched_flops_per_step": conv_flops, "conv_ms_per_step": round(conv_ms, 4),
                "note": "achieved = the reference's algorithmic FLOPs (sum of 2*M*N*K over ITS 3D conv list, SURVEY 8d: 1344 GFLOP "
                        "per pair) / summed CUDA-event durations of every launch that executes those convs (tensor-core conv "
                        "launches + collapsed-stem0 assemble + head-taps kernels); achieved_as_launched counts only the "
                        "2*M*N*K actually launched after the exact algebraic rewrites named in config.rewrites (27 % of the "
                        "reference's FLOPs - stem0 on the un-masked volume, last_3 on the up-sampled volume - are never "
                        "issued); split-precision modes issue planes*(planes+1)/2 tensor-core MACs per launched MAC; "
                        "traffic = ncu DRAM bytes per conv launch, averaged over the launches of one step"}
    line = {"metric": "stereo pairs/sec, KITTI 384x1248 D=192", "value": round(value, 4), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {1: "bf16", 2: "bf16x3 (2 bf16 planes, fp32 accumulate)", 3: "bf16x6 (3 bf16 planes, fp32 accumulate)"}[
                args.planes] if args.conv == "tc" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "pairs_per_gpu_per_step": B, "conv": args.conv, "planes": args.planes,
                       "mma_terms": args.mma_terms, "cuda_graph": graph_on,
                       "parallelism": "pairs sharded, no collective",
                       "rewrites": {"collapse_stem0": plan_facts["collapse_stem0"], "fuse_cost_volume": plan_facts["fuse_cv"],
                                    "fuse_head": plan_facts["fuse_head"],
                                    "note": "exact algebraic rewrites of the plan (DESIGN.md 3): stem0 separates into 2-D convs "
                                            "where the cost volume is un-masked; last_3's channel contraction runs before "
                                            "upsample_6"},
                       "tc_fold": os.environ.get("LEA_TC_FOLD", "2") + " (2 = incl. the one-tile fold of the 32-output convs, which adds a "
                                  "neighbour-tap a_lo*W_lo term <= 2^-18 relative, DESIGN.md 4.1; gated by the `parity` leg)",
                       "feature_planes": plan_facts["feature_planes"], "accum_split": plan_facts["accum_split"],
                       "activation_arena_GB": plan_facts["arena_GB"],
                       "l2": "per-step activation working set (>4 GB of planes volumes) exceeds the 126 MB L2; no flush needed",
                       "weights": "random init seed 0", "feature_net": "native kernels (fused stems + tcgen05 convs on depth-1 volumes), inside the timed step"},
            "e2e": {"value": round(e2e_value, 4), "unit": "pairs/s", "h2d_bytes_per_step": h2d_per_step,
                    "d2h_bytes_per_step": d2h_per_step,
                    "path": "pinned 8-bit 375x1242 image pairs -> pipeline.InputPipeline (H2D, z-normalise, pad to 384x1248: "
                            "predict.py:144-184) -> LEAStereo.forward -> disparity maps in pinned host memory"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": sampler.summary(), "roofline": roofline, "kernels": kernels}
    if train is not None:
        line["train"] = train
    if args.dump_launches:
        json.dump([list(r) for r in per_launch], open(args.dump_launches, "w"), indent=0)
    if world == 1:
        if not args.no_parity:
            try:
                line["parity"] = parity_leg(device, options)
            except Exception as e:  # noqa: BLE001
                line["parity"] = {"error": str(e)[:300]}
        if not args.no_gpu_baseline:
            try:
                line["gpu_baseline"] = gpu_baseline_leg(device, B)
            except Exception as e:  # noqa: BLE001
                line["gpu_baseline"] = {"error": str(e)[:300]}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            t_band = time_oracle(96, 1, 1)[0]
            rows = 384 if t_band * 4 <= 30.0 else 96
            ts = time_oracle(rows, 1, 0) if rows != 96 else [t_band]
            frac = rows / 384.0
            line["cpu_baseline"] = {"value": round(frac / ts[0], 5), "unit": "pairs/s", "cores": cores, "kind": "port",
                                    "sample": "%d-row band of one 384x1248 pair (%.2f pair), D=192, 1 forward of the fp32 "
                                              "PyTorch-CPU oracle port" % (rows, frac)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
