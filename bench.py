#!/usr/bin/env python
"""bench.py -- BLIP-NLVR pruned forward (p = 0.5) images/sec on B200, the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one BLIP_NLVR.forward(train=False) over this rank's 32 synthetic pairs (64 images 384x384, 20-token
sentences, seeded random weights) at the temperature calibrated on the oracle for p = 0.5
(tests/golden/calib_nlvr_p50_b32.npz). Weak scaling: every rank owns its own 32 pairs; weights are broadcast once from
rank 0. The path has no exchange step: like the reference's evaluation loop (compress_nlvr_dtp.py:82-104, one
`synchronize_between_processes` after the last batch) the ranks run their batches independently and the results are
gathered ONCE, after the last step, inside the timed region.

  value     whole-job images/s with the inputs already in HBM (CUDA events, max over ranks, barrier + sync both sides)
  e2e       the same forward through the public module API with pinned-host inputs: H2D of images/ids + D2H of logits
            inside the timed region
  roofline  the dominant kernel (found in an instrumented warm-up step, after the one-off weight preparation), its algorithmic FLOPs / its CUDA-event time
  cpu_baseline  the oracle (CPU restatement of the reference forward, oracle/dtp_oracle.py) on this box's host cores,
            bounded sample; `--impl reference` times the same thing as its own arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from bench_workloads import WORKLOADS  # noqa: E402


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class NvmlClockSampler:
    """SM clock and throttle reasons through NVML (the library nvidia-smi itself uses) from a background thread, every
    50 ms while the timed region runs. In-process NVML queries perturb the launch thread far less than forking
    `nvidia-smi -lms` next to it (measured: the subprocess cost rank 0 up to 10 % of a 24 ms step)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.sm, self.reasons, self.stop_flag, self.thread = [], set(), False, None

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        mx = None
        try:
            mx = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": mx, "samples": len(self.sm),
                "reasons": sorted(self.reasons), "source": "nvml"}


def make_clock_sampler(index: int):
    if os.environ.get("MADTP_CLOCKS", "nvml") == "nvml":
        try:
            return NvmlClockSampler(index)
        except Exception:
            pass
    return ClockSampler(index)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (fallback when NVML's
    Python binding is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own forward on the host cores (reference arm and cpu_baseline leg)
# ---------------------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """`--impl reference`: the UNMODIFIED reference modules of the chosen configuration (staged byte for byte in
    oracle/_ref by oracle/make_ref.py, imported under the third-party shims of oracle/ref_shims.py) on all host threads;
    the oracle port when the reference tree is absent. Rank 0 only."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    w = WORKLOADS[args.config]()
    rate, sec, kind, sample = w.cpu_reference(args.sample, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = f"{sample}, PyTorch-CPU fp32, {cores} threads"
    line = {"impl": "reference", "metric": w.metric, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.workload, "baseline_config": w.config, "temperature": w.temperature,
                       "sample": sample},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpu_baseline_subprocess(config: int, sample: int, steps: int, warmup: int):
    """The cpu_baseline leg of the main arm: the reference arm in its OWN process (no CUDA context, no sampler thread
    next to it -- round 1's in-process leg read 28 % low for that reason)."""
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", str(config), "--steps", str(steps),
           "--warmup", str(warmup), "--sample", str(sample)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    except subprocess.TimeoutExpired:
        return {"value": None, "unit": "images/s", "cores": None, "kind": "unavailable", "sample": "timed out"}
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    return {"value": None, "unit": "images/s", "cores": None, "kind": "unavailable", "sample": r.stderr[-300:]}


# ---------------------------------------------------------------------------------------------------------------
# roofline bookkeeping
# ---------------------------------------------------------------------------------------------------------------
def algorithmic_flops(name, meta):
    """Algorithmic FLOPs of one launch from its recorded shape arguments."""
    if name == "madtp_attn_fwd":
        B, H, Nq, Nk = meta
        return 4.0 * B * H * Nq * Nk * 64            # QK^T and PV, 2 FLOPs per MAC
    if name == "madtp_attn_tc_fwd":
        B, H, N = meta
        return 4.0 * B * H * N * N * 64
    if name == "madtp_attn_tc_stats":
        B, H, N = meta
        return 2.0 * B * H * N * N                   # max over heads + column sum; the QK^T recompute is not credited
    if name == "madtp_gemm_qkv":
        M, K, heads = meta
        return 2.0 * M * (3 * heads * 64) * K
    if name.startswith("madtp_gemm"):
        _, M, N, K = meta
        return 2.0 * M * N * K
    if name == "madtp_attn_stats":
        B, H, N = meta
        return 2.0 * B * H * N * N                   # max over heads + column sum; the QK^T recompute is not credited
    return 0.0



def ncu_traffic(kernel_key: str):
    """DRAM bytes per launch of the dominant kernel class from the kept `ncu --set full` summary (profiles/
    ncu_traffic.json, written by profiles/summarize.py from the capture of this same command), or None."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return None
    try:
        return json.loads(p.read_text()).get(kernel_key)
    except Exception:
        return None


_JSON_OUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: libraries that print there (NCCL's version banner, torchrun notices) are sent
    to stderr for the rest of the process."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="madtp_b200", choices=["madtp_b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS),
                    help="BASELINE.json configuration (2 = the headline BLIP-NLVR forward)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sample", "--sample-pairs", type=int, default=0, dest="sample",
                    help="reference arm: pairs / images per step (0 = the full batch)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the "
                                                            "captured CUDA graph (device-resident lengths either way)")
    ap.add_argument("--host-lengths", action="store_true", help="round-1 path: read topk_num back once per layer")
    ap.add_argument("--streams", type=int, default=2, help="batches in flight: consecutive steps alternate between this "
                    "many CUDA streams, each with its own captured graph and buffers (madtp_b200.pipeline.StreamPool); "
                    "1 = every step on the current stream")
    ap.add_argument("--value-lane", default="f16", choices=["f16", "split"],
                    help="split: diagnostic -- the ViT's value lane (attention output projection, FFN) also runs on the "
                         "error-compensated fp16 hi/lo planes with an fp32 context and the exact GELU, to show the "
                         "free-running keep-mask agreement without the value lane's fp16 rounding (slower)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "madtp_b200" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    from madtp_b200 import _lib, dist as mdist, vit as mvit
    from madtp_b200.pipeline import InputPrefetcher, StreamPool
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: madtp_b200 has no CPU fallback")
    _lib.load()
    rank, local_rank, world = mdist.init("nccl")
    if world != args.gpus:
        raise RuntimeError(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    w = WORKLOADS[args.config]()

    # weights: rank 0 draws the seeded state dict, everyone else receives it over NCCL
    model = w.build(dev, rank)
    mdist.broadcast_parameters(model, src=0)
    mvit.device_lengths_enabled(not args.host_lengths)
    if args.value_lane == "split":
        from madtp_b200 import functional as mfn
        mfn.value_lane_split(True)
    use_graph = w.graphable and not (args.no_graph or args.host_lengths)
    w.enable_graphs(use_graph)

    host = tuple(t.pin_memory() for t in w.host_inputs(rank))          # this rank's own batch
    resident = tuple(t.to(dev) for t in host)
    result_shape = None

    n_streams = args.streams if (w.graphable and not (args.no_graph or args.host_lengths)) else 1
    pool = StreamPool(dev, n_streams) if n_streams > 1 else None
    step_no = [0]

    def step_resident():
        if pool is None:
            return w.step(resident)
        i = step_no[0]
        step_no[0] += 1
        with pool.stream(i, wait_current=False):     # the whole step (one graph replay) on its own stream
            return w.step(resident)

    def join_streams():
        if pool is not None:
            pool.join()

    def fork_streams():
        if pool is not None:
            cur = torch.cuda.current_stream()
            for s in pool.streams:
                s.wait_stream(cur)

    feeder = InputPrefetcher(dev, host, depth=args.streams + 1)
    e2e_state = {"i": 0, "start": 0, "end": 0, "out_h": None}

    def step_e2e():
        """Public API with HOST inputs: every step's pinned-host -> HBM copy (issued one step ahead on a side stream,
        inside the timed region) and the device -> host read of the result."""
        i = e2e_state["i"]
        if i == e2e_state["start"]:
            feeder.submit(i, host)
        if i + 1 < e2e_state["end"]:
            feeder.submit(i + 1, host)

        def body():
            out = w.step(feeder.acquire(i))
            feeder.release(i)
            if e2e_state["out_h"] is None:
                e2e_state["out_h"] = [torch.empty(out.shape, dtype=out.dtype).pin_memory() for _ in range(max(1, n_streams))]
            e2e_state["out_h"][i % max(1, n_streams)].copy_(out, non_blocking=True)
            return out
        if pool is None:
            out = body()
        else:
            with pool.stream(i, wait_current=False):   # H2D wait, forward and D2H of step i on stream i % n
                out = body()
        e2e_state["i"] = i + 1
        return out

    class per_kernel_mode:
        """Per-kernel CUDA-event timing needs every launch to come from Python with its exact shape arguments: the
        instrumented passes run the SAME kernels with host-side lengths (one read-back per layer, the round-1 path;
        results are bit-identical, tests/test_device_lengths_gpu.py) instead of replaying the graph."""

        def __enter__(self):
            mvit.device_lengths_enabled(False)

        def __exit__(self, *exc):
            mvit.device_lengths_enabled(not args.host_lengths)

    # ---- warm-up; one warm-up step is fully instrumented: it records every launch's exact shape arguments ----
    with per_kernel_mode():
        step_resident()                   # first call also prepares the GEMM-ready weight copies (one-off)
        torch.cuda.synchronize()
        timer = _lib.LaunchTimer()
        _lib.set_launch_timer(timer)
        step_resident()
        torch.cuda.synchronize()
        _lib.set_launch_timer(None)
    shapes_pass = timer.records            # (kernel, e0, e1, shape arguments) in launch order, host-side lengths
    for _ in range(max(args.warmup, 3) * max(1, args.streams)):   # graph mode: arena passes, the capture, then replays
        step_resident()
    torch.cuda.synchronize()

    by_rank = []                      # per timed region: every rank's own ms per step (the reported time is their max)

    def timed(fn, steps, only=None):
        t = _lib.LaunchTimer(only=only) if only else None
        mdist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.launch_count() + w.graph_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_own = torch.cuda.Event(enable_timing=True)
        _lib.set_launch_timer(t)
        e0.record()
        fork_streams()
        last = None
        for _ in range(steps):
            last = fn()
        join_streams()
        e_own.record()                       # this rank's own steps are done (diagnostics: ms_per_step_by_rank)
        mdist.all_gather_rows(last)          # the job's one collective: every rank's last results, after the last step
        e1.record()
        _lib.set_launch_timer(None)
        torch.cuda.synchronize()
        mdist.barrier()
        by_rank.append(mdist.gather_over_ranks(e0.elapsed_time(e_own) / steps, dev))
        ms = mdist.max_over_ranks(e0.elapsed_time(e1), dev)
        return ms, _lib.launch_count() + w.graph_launches() - l0, t

    def backlogged_pass(steps):
        mdist.barrier()
        torch.cuda.synchronize()
        recs, total = [], 0.0
        held = w.suspend_graphs()          # device-resident lengths stay on: no read-back drains the queue mid-step
        try:
            for _ in range(2):             # arena passes of the un-graphed path
                w.step(resident)
            torch.cuda.synchronize()
            for _ in range(steps):
                t = _lib.LaunchTimer()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(60_000_000)          # ~30 ms: the host needs ~10 ms to queue 410 launches + events
                _lib.set_launch_timer(t)
                e0.record()
                w.step(resident)
                e1.record()
                _lib.set_launch_timer(None)
                torch.cuda.synchronize()
                recs.append([(n, a.elapsed_time(b)) for (n, a, b, _) in t.records])
                total += e0.elapsed_time(e1)
        finally:
            _lib.set_launch_timer(None)
            w.resume_graphs(held)
        return recs, total

    sampler = make_clock_sampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, _ = timed(step_resident, args.steps)
    # (the clock sampler covers exactly the timed region above)
    # Per-kernel device times: the same steps once more, launched from Python with CUDA events around every launch
    # (kept out of the pass that produces `value`). The GPU is held back by a spin kernel at the start of each step so
    # that the host has queued the whole step before the first kernel runs: the events then bracket kernel time, not
    # the host's launch gaps (which dominate the ~5 us text-encoder kernels otherwise).
    clocks = sampler.stop() if sampler else None
    timed_records, ms_instr = backlogged_pass(min(args.steps, 10))
    e2e_state.update(i=0, start=0, end=1)
    step_e2e()
    torch.cuda.synchronize()
    e2e_state.update(start=e2e_state["i"], end=e2e_state["i"] + args.steps)
    ms_e2e, _, _ = timed(step_e2e, args.steps)

    units_per_step = w.units * world
    value = units_per_step * args.steps / (ms / 1e3)
    e2e_value = units_per_step * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel class: shapes from the host-length pass, times from the backlogged pass ----
    def per_name(records, field):
        d = {}
        for r in records:
            d.setdefault(r[0], []).append(r[field])
        return d
    metas = per_name(shapes_pass, 3)
    times = {}                                                   # kernel -> per-launch ms summed over the timed steps
    for step_recs in timed_records:
        for name, lst in per_name(step_recs, 1).items():
            acc = times.setdefault(name, [0.0] * len(lst))
            if len(acc) == len(lst):
                for i, v in enumerate(lst):
                    acc[i] += v
    n_timed = max(len(timed_records), 1)
    prof = {k: {"launches": len(v), "ms": sum(v) / n_timed} for k, v in times.items()}
    top = max(prof.items(), key=lambda kv: kv[1]["ms"])[0]
    busy_ms = sum(v["ms"] for v in prof.values())                # GPU-busy time of one step: the sum of its kernels
    matched = top in metas and len(metas[top]) == len(times[top])
    if matched:
        top_meta, top_ms = metas[top], [v / n_timed for v in times[top]]
        timed_in = ("a second pass of the same K steps, launched from Python with CUDA events around every launch and "
                    "the GPU held back at the start of each step until the host has queued it (events bracket kernel "
                    "time, not host launch gaps); shape arguments from an instrumented warm-up step with host-side "
                    "lengths (same kernels in the same order, bit-identical results); the kernels inside the CUDA graph "
                    "of the `value` pass cannot be bracketed by events")
    else:   # the two passes launched this kernel a different number of times: fall back to the host-length pass alone
        top_meta = metas.get(top, [])
        top_ms = [e0.elapsed_time(e1) for (n, e0, e1, _) in shapes_pass if n == top]
        timed_in = "one instrumented warm-up step with host-side lengths (includes host launch gaps)"
    class_ms = sum(top_ms)
    flops = sum(algorithmic_flops(top, m) for m in top_meta)
    achieved = flops / (class_ms / 1e3) / 1e12 if class_ms > 0 else 0.0
    traffic = ncu_traffic(top) if args.config == 2 else None
    roofline = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                "traffic_detail": traffic,
                "peak_source": peaks["source"] + " (sustained bf16, kernel timed inside a long step)",
                "launches_per_step": len(top_ms), "ms_per_step": class_ms,
                "share_of_step": class_ms / busy_ms if busy_ms > 0 else None,
                "gpu_busy_ms_per_step": busy_ms, "instrumented_ms_per_step": ms_instr / n_timed,
                "algorithmic_tflop_per_launch": flops / max(len(top_ms), 1) / 1e12,
                "timed_in": timed_in,
                "note": "algorithmic FLOPs per launch as in DESIGN.md section 6 (error-compensation passes and the "
                        "second QK^T pass are not credited); share_of_step = this class / the sum of all kernels of a step"}
    # the dominant kernel class by launch shape: where inside the class the time goes
    by_shape = {}
    for meta, tms in zip(top_meta, top_ms):
        d = by_shape.setdefault(tuple(meta), [0, 0.0])
        d[0] += 1
        d[1] += tms
    shapes = []
    for meta, (cnt, tms) in sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:12]:
        fl = algorithmic_flops(top, meta) * cnt
        shapes.append({"shape": list(meta), "launches_per_step": cnt, "us_per_launch": 1e3 * tms / cnt,
                       "tflops": fl / (tms / 1e3) / 1e12 if tms > 0 else 0.0,
                       "share_of_class": tms / class_ms if class_ms > 0 else 0.0})
    roofline["by_shape"] = shapes
    step_flops = w.step_flops()                               # oracle trajectory, per rank
    step_roofline = {"algorithmic_tflop_per_step": step_flops / 1e12,
                     "achieved": step_flops / (ms / args.steps / 1e3) / 1e12, "peak": peaks["tflops_sustained"],
                     "unit": "TFLOP/s", "frac": step_flops / (ms / args.steps / 1e3) / 1e12 / peaks["tflops_sustained"]}
    breakdown = {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in
                 sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_subprocess(args.config, 0, 3, 1)

    parity = None
    out = step_resident()
    join_streams()
    torch.cuda.synchronize()
    if rank == 0:
        parity = w.parity(out)
        execution = ("CUDA graph replay, device-resident token counts (0 host read-backs per step)" +
                     (f", {n_streams} batches in flight on {n_streams} CUDA streams (pipeline.StreamPool)" if pool else "")
                     if use_graph else
                     "Python-issued launches, " + ("one read-back per pruned layer" if args.host_lengths or args.config == 4
                                                   else "device-resident token counts, one read-back per encoder call"))
        cfg = {"workload": w.workload, "baseline_config": w.config, "temperature": w.temperature,
               "parallelism": f"batch-shard x{world}", "execution": execution,
               "value_lane": "f16 operands (default)" if args.value_lane == "f16" else
               "DIAGNOSTIC: error-compensated fp16 hi/lo planes + fp32 context + exact GELU in the ViT (--value-lane split)",
               "collectives": "none per step; one all_gather of the last step's results closes the timed region",
               "l2": "per-step working set (weights + activations) exceeds the 126 MB L2; no flush"}
        cfg.update(w.config_extra())
        line = {"metric": w.metric, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_by_rank": by_rank[0],
                "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": "f32-accurate scoring lane (error-compensated fp16 hi/lo planes on tcgen05, fp32 accumulate) + f16 value lane",
                "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in host)),
                        "d2h_bytes_per_step": int(e2e_state["out_h"][0].numel() * e2e_state["out_h"][0].element_size())},
                "gpu_launches": int(launches),
                "roofline": roofline, "step_roofline": step_roofline, "kernel_ms_one_step": breakdown,
                "parity": parity, "cpu_baseline": cpu_baseline}
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
