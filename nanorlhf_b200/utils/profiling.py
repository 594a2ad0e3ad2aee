"""Phase timers (CUDA events), NVTX ranges, memory high-water marks -> ``time/*`` / ``mem/*`` metrics.

The reference only prints wall-clock ``s/episode`` (/root/reference/GRPO/grpo_trainer.py:726) and
imports an unused memory profiler.  Here every phase of an update (rollout, reward,
logprob, advantage, train, comm, ckpt) is bracketed with CUDA events on the current stream and an
NVTX range, with no host synchronisation until the metrics are read at the end of the update.
"""
from __future__ import annotations

import contextlib
import os
import time
from collections import OrderedDict
from typing import Dict

import torch


class PhaseTimer:
    def __init__(self, device=None, nvtx: bool = False):
        self.device = torch.device(device) if device is not None else None
        self.cuda = self.device is not None and self.device.type == "cuda"
        self.nvtx = nvtx and self.cuda
        self._events: "OrderedDict[str, list]" = OrderedDict()
        self._wall: Dict[str, float] = {}

    @contextlib.contextmanager
    def phase(self, name: str):
        if self.nvtx:
            torch.cuda.nvtx.range_push(name)
        if self.cuda:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self._wall[name] = self._wall.get(name, 0.0) + (time.perf_counter() - t0)
            if self.cuda:
                e.record()
                self._events.setdefault(name, []).append((s, e))
            if self.nvtx:
                torch.cuda.nvtx.range_pop()

    def collect(self, reset: bool = True) -> Dict[str, float]:
        """Seconds per phase (device time when on CUDA, else wall)."""
        out: Dict[str, float] = {}
        if self.cuda:
            torch.cuda.synchronize(self.device)
            for k, evs in self._events.items():
                out[f"time/{k}_s"] = sum(s.elapsed_time(e) for s, e in evs) / 1e3
            for k, v in self._wall.items():
                out[f"time/{k}_wall_s"] = v
            out["mem/peak_allocated_gb"] = torch.cuda.max_memory_allocated(self.device) / 2**30
            out["mem/peak_reserved_gb"] = torch.cuda.max_memory_reserved(self.device) / 2**30
        else:
            for k, v in self._wall.items():
                out[f"time/{k}_s"] = v
        if reset:
            self._events.clear()
            self._wall.clear()
        return out


def debug_sync_enabled() -> bool:
    """``NANORLHF_DEBUG_SYNC=1``: sync + finite-check after each phase."""
    return os.environ.get("NANORLHF_DEBUG_SYNC", "0") == "1"


def check_finite(name: str, *tensors):
    if not debug_sync_enabled():
        return
    for t in tensors:
        if t is not None and t.is_floating_point() and not torch.isfinite(t).all():
            raise FloatingPointError(f"non-finite values after {name}")
    if torch.cuda.is_available():
        torch.cuda.synchronize()
