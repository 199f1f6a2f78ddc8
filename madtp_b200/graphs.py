"""CUDA-graph execution of a forward with device-resident lengths.

With the pruned token counts kept on the device (functional.dtp_finish, include/madtp_b200.h "Device-resident lengths")
the launch sequence of a forward no longer depends on the data: the same ~400 kernel launches, with the same arguments,
for every batch of a given shape. `GraphedCall` runs such a forward twice for warm-up inside an `_lib.Arena` (every
buffer the wrappers allocate becomes a persistent, zero-initialised buffer handed out in call order), captures the third
run in a torch.cuda.CUDAGraph and replays it from then on: one graph launch per forward instead of 400 Python-issued
launches and 24 blocking read-backs (the reference: one `.item()` per pruned layer, models/vit.py:145).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from . import _lib as L


class GraphedCall:
    """fn(*static_inputs) -> (outputs, trajectories): captured once, replayed on fresh inputs of the same shapes."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 2):
        self.static_inputs = [t.clone() for t in example_inputs]
        self.arena = L.Arena()
        self.fn = fn
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.arena.frozen = True
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        # thread-local capture mode: other threads of the process (NCCL's watchdog, bench.py's NVML clock sampler) may
        # keep calling the CUDA API while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.outputs, self.trajectories = self._run()
        self.launches = L.launch_count() - n0      # kernels of libmadtp_b200.so inside one replay of the graph
        self.replays = 0

    def _run(self):
        self.arena.begin()
        prev = L.set_arena(self.arena)
        try:
            return self.fn(*self.static_inputs)
        finally:
            L.set_arena(prev)

    def matches(self, inputs: Sequence[torch.Tensor]) -> bool:
        return len(inputs) == len(self.static_inputs) and all(
            a.shape == b.shape and a.dtype == b.dtype and a.device == b.device
            for a, b in zip(inputs, self.static_inputs))

    def __call__(self, *inputs: torch.Tensor):
        """Copies `inputs` into the captured buffers (device or pinned-host sources, asynchronously) and replays.
        Returns the captured output tensors: they are overwritten by the next replay."""
        for dst, src in zip(self.static_inputs, inputs):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        for t in self.trajectories:
            t.invalidate()
        return self.outputs


class GraphedForward:
    """CUDA-graph execution of a forward with device-resident lengths (see blip_nlvr.BLIP_NLVR.enable_cuda_graphs):
    one captured graph, with its own persistent buffers, per entry point, input shapes, temperature and stream."""
    _graphs = None

    def enable_cuda_graphs(self, enable: bool = True):
        """Capture the pruned evaluation forward per input shape and replay it on later calls. The graphs bake in the
        addresses of the weights and of their GEMM-ready copies: call `reset_cuda_graphs()` after changing weights."""
        self._graphs = {} if enable else None
        return self

    def reset_cuda_graphs(self):
        if self._graphs is not None:
            self._graphs = {}

    def _device_path(self, temperature, *tensors):
        from .vit import device_lengths_enabled
        return temperature > 0 and device_lengths_enabled() and all(t.is_cuda for t in tensors)

    def _run_device(self, tag, fn, inputs, temperature, extra=()):
        """fn(*inputs) -> (outputs, trajectories) with data-independent launches: inside an arena (persistent buffers,
        Python-issued launches), or as a replay of the captured graph when enable_cuda_graphs(True). `extra`: further
        host-side values the launch sequence depends on (they become part of the arena / graph key)."""
        inputs = [t.contiguous() for t in inputs]
        key = (tag,) + tuple(tuple(t.shape) for t in inputs) + (float(temperature),) + tuple(extra)
        if self._graphs is None:
            with L.arena_for(self, key):    # clones: the arena's buffers are reused by the next call with this key
                return tuple(None if o is None else o.clone() for o in fn(*inputs)[0])
        key = key + (torch.cuda.current_stream().cuda_stream,) + L.mode_key()
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = GraphedCall(fn, inputs)
        return g(*inputs)
