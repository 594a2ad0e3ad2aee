"""nvidia-smi clock / throttle-reason sampler for benchmark records (B200_PROFILING.md "clocks line").

Every number this repo reports (``bench.py``, ``bench/*_bench.py``, ``baseline/stack_bench.py``) carries the
clocks seen *during* its timed region: median SM clock under load, the maximum SM clock, and the set of throttle
reasons that were active.  A run that saw ``hw_slowdown`` / ``hw_thermal_slowdown`` / ``sw_thermal_slowdown`` is not
a valid measurement; ``sw_power_cap`` on a 1 kW part under a dense GEMM is normal and only noted.
"""
from __future__ import annotations

import statistics
import subprocess
import threading

_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
      "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
_REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")


class ClockSampler:
    def __init__(self, gpu_index: int = 0, period_ms: int = 200):
        self.rows, self.proc, self.idx, self.period = [], None, gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={_Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", str(self.period)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
        ok = [r for r in self.rows if len(r) >= 9]
        sm = [float(r[1]) for r in ok if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in ok if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in ok if r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in ok:
            for name, v in zip(_REASONS, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.result = self.stop()
        return False
