"""Per-launch timeline of one bench step (development aid): every C-ABI launch of a warm step, in launch order, with
its shape arguments and its CUDA-event time (Python-issued launches with host-side lengths, the same kernels the graph
replays). Averages over --steps instrumented steps.

    python scripts/step_profile.py [--config 2] [--steps 3] > gpurun_out/step_profile.txt
"""
import argparse
import sys
from collections import OrderedDict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench_workloads import WORKLOADS  # noqa: E402
from madtp_b200 import _lib, vit as mvit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_lib.load()
w = WORKLOADS[args.config]()
model = w.build(dev, 0)
mvit.device_lengths_enabled(False)
w.enable_graphs(False)
inputs = tuple(t.to(dev) for t in w.host_inputs(0))
for _ in range(3):
    w.step(inputs)
torch.cuda.synchronize()
runs = []
for _ in range(args.steps):
    timer = _lib.LaunchTimer()
    _lib.set_launch_timer(timer)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w.step(inputs)
    e1.record()
    _lib.set_launch_timer(None)
    torch.cuda.synchronize()
    runs.append(([(n, tuple(m), a.elapsed_time(b) * 1e3) for (n, a, b, m) in timer.records], e0.elapsed_time(e1)))
base = runs[0][0]
print(f"# config {args.config}: {len(base)} launches, instrumented step {sum(r[1] for r in runs) / len(runs):.2f} ms")
tot = OrderedDict()
for i, (name, meta, _) in enumerate(base):
    us = sum(r[0][i][2] for r in runs if len(r[0]) == len(base)) / sum(1 for r in runs if len(r[0]) == len(base))
    print(f"{i:4d} {name:28s} {str(list(meta)):34s} {us:8.1f}")
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += us
print("# totals")
for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:28s} {c:4d} {us / 1e3:8.3f} ms")
print(f"{'sum':28s} {sum(c for c, _ in tot.values()):4d} {sum(u for _, u in tot.values()) / 1e3:8.3f} ms")
