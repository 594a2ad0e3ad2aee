"""Failure detection: fault injection hooks, per-rank heartbeat, watchdog.

The reference has none of this.  ``NANORLHF_FAULT=rank:phase:step`` makes the named
rank raise (``:exit`` suffix -> ``os._exit(17)``) when it enters ``phase`` of update ``step`` --
used by the distributed tests to check that surviving ranks exit non-zero within the watchdog
timeout and that a restart resumes from the last checkpoint.
"""
from __future__ import annotations

import json
import os
import threading
import time
from typing import Optional


class InjectedFault(RuntimeError):
    pass


def maybe_inject(rank: int, phase: str, step: int):
    spec = os.environ.get("NANORLHF_FAULT")
    if not spec:
        return
    parts = spec.split(":")
    if len(parts) < 3:
        return
    r, p, s = parts[0], parts[1], parts[2]
    if int(r) == rank and p == phase and int(s) == step:
        if len(parts) > 3 and parts[3] == "exit":
            os._exit(17)
        raise InjectedFault(f"injected fault at rank={rank} phase={phase} step={step}")


class Heartbeat:
    """Writes ``<dir>/heartbeat_<rank>.json`` at each phase boundary; a watchdog thread aborts the
    process if no beat lands within ``timeout_s`` (a hung collective / kernel never spins forever)."""

    def __init__(self, directory: str, rank: int, timeout_s: float = 1800.0, enable_watchdog: bool = True):
        self.path = os.path.join(directory, f"heartbeat_{rank}.json")
        os.makedirs(directory, exist_ok=True)
        self.rank, self.timeout_s = rank, timeout_s
        self._last = time.time()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        if enable_watchdog and timeout_s > 0:
            self._thread = threading.Thread(target=self._watch, daemon=True)
            self._thread.start()

    def beat(self, phase: str, step: int):
        self._last = time.time()
        tmp = self.path + ".tmp"
        with open(tmp, "w") as f:
            json.dump({"rank": self.rank, "phase": phase, "step": step, "time": self._last}, f)
        os.replace(tmp, self.path)

    def _watch(self):
        while not self._stop.wait(min(5.0, self.timeout_s / 4)):
            if time.time() - self._last > self.timeout_s:
                print(f"[watchdog] rank {self.rank}: no heartbeat for {self.timeout_s}s -- aborting", flush=True)
                os._exit(98)

    def close(self):
        self._stop.set()
