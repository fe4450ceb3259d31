"""CUDA-graph replay of a policy step.

A policy step is ~230 kernel launches issued from Python through ctypes; at the 200M / 256-episode size the GPU never
starves, but at the small configurations (VIMA-20M, 64 episodes: ~2 ms of kernels) and in `forward_step` (33 rows per
episode) the launch path is the bound.  Shapes are static from step to step, every kernel takes its arguments by value
(tensor maps are `__grid_constant__` parameters, the tiny grouped GEMMs carry their descriptors in parameter space), and
nothing on the step synchronises with the host after the first call, so the whole step captures into one graph:

    g = GraphedStep(lambda obs: step(obs), example_obs)   # warms up, then captures on a side stream
    out = g(new_obs)                                        # copies new_obs into the static inputs, replays

The outputs are the tensors the captured call returned (static storage, overwritten by the next replay).
"""
from __future__ import annotations

from typing import Callable

import torch

from . import _C


def _map(fn, x, y=None):
    if isinstance(x, dict):
        return {k: _map(fn, v, None if y is None else y[k]) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_map(fn, v, None if y is None else y[i]) for i, v in enumerate(x))
    return fn(x) if y is None else fn(x, y)


class GraphedStep:
    def __init__(self, fn: Callable, example_inputs, *, warmup: int = 3):
        """fn(inputs) -> nest of tensors; `example_inputs`: nest (dict / list / tensor) of CUDA tensors with the step's shapes."""
        self.fn = fn
        self.static_in = _map(lambda t: t.clone(), example_inputs)
        dev = next(iter(self._leaves(self.static_in))).device
        self.ctx = _C.Context.get(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # packs weights, fills caches, runs the first-call host checks
                fn(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = self.ctx.launches
        with torch.cuda.graph(self.graph):
            self.static_out = fn(self.static_in)
        self.kernels_per_replay = self.ctx.launches - n0  # vima:: kernels inside the graph (torch's own copies come on top)
        self.replays = 0

    @staticmethod
    def _leaves(x):
        if isinstance(x, dict):
            for v in x.values():
                yield from GraphedStep._leaves(v)
        elif isinstance(x, (list, tuple)):
            for v in x:
                yield from GraphedStep._leaves(v)
        else:
            yield x

    def __call__(self, inputs):
        if inputs is not self.static_in:
            _map(lambda dst, src: dst.copy_(src, non_blocking=True) if dst.data_ptr() != src.data_ptr() else dst, self.static_in, inputs)
        self.graph.replay()
        self.replays += 1
        return self.static_out

    def load_inputs(self, src):
        """Copy `src` (same nest; pinned host or device tensors) straight into the graph's static input buffers, asynchronously on
        the current stream; follow with `self(self.static_in)` to replay without a second device-to-device copy."""
        _map(lambda dst, s: dst.copy_(s, non_blocking=True), self.static_in, src)
        return self.static_in

    def describe(self) -> dict:
        return {"vima_kernels_per_replay": int(self.kernels_per_replay),
                "note": "the step is captured once (torch.cuda.CUDAGraph) and replayed; inputs are copied into static buffers"}
