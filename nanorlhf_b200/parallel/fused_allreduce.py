"""Host side of K-AR: symmetric-memory allocation + the fused all-reduce/AdamW launch.

``torch.distributed._symmetric_memory`` is used *only* as the allocator / handle exchanger (CUDA VMM
allocation, fabric-handle exchange over the bootstrap process group, NVLS multicast object, signal
pads); the data path is csrc/comm.cu -- peer / multicast loads and stores issued from inside the
kernel.  There is no NCCL call on this path.
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch
import torch.distributed as dist

from ..ops import native


class FusedAllReduceAdam:
    def __init__(self, comm, use_multicast: str = "auto", max_blocks: int = 148 * 4):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.comm = comm
        self.group = dist.group.WORLD
        self.handles: Dict[int, object] = {}
        self.max_blocks = max_blocks
        self.use_multicast = os.environ.get("NANORLHF_NVLS", use_multicast)
        self.bytes_reduced = 0
        # every cross-rank wait is bounded: a rank that died or diverged turns into a device-side trap (CUDA error -> non-zero
        # exit -> restart from the last checkpoint) instead of an infinite spin that only the job scheduler can end
        self.barrier_timeout_ms = int(float(os.environ.get("NANORLHF_BARRIER_TIMEOUT_S", "900")) * 1000)
        self.wait_events = []      # (start, end) around the opening barrier of every K-AR step: time spent waiting for the slowest rank
        native.load()

    def alloc(self, n: int, dtype: torch.dtype, device) -> torch.Tensor:
        """Symmetric (peer-mapped) flat buffer.  Collective: every rank must call with the same size."""
        t = self.symm_mem.empty(n, dtype=dtype, device=device)
        t.zero_()
        hdl = self.symm_mem.rendezvous(t, self.group.group_name)
        self.handles[t.data_ptr()] = hdl
        return t

    def handle(self, t: torch.Tensor):
        return self.handles[t.data_ptr()]

    def register(self, flats: List):
        self.flats = flats

    def _mc(self, hdl) -> int:
        if self.use_multicast in ("0", "off", "never"):
            return 0
        if self.use_multicast == "auto" and self.comm.world_size <= 2:
            return 0          # measured (profiles/dist_check_2gpu_r1.json): at N=2 the P2P kernel is 1.6x faster than NVLS

        return int(getattr(hdl, "multicast_ptr", 0) or 0)

    @torch.no_grad()
    def allreduce_adam(self, f, hp: dict, scale: float):
        """reduce-scatter(grad) + AdamW on the owned shard + all-gather(param), one kernel."""
        gh, ph = self.handle(f.grad), self.handle(f.param)
        from .optimizer import shard_bounds
        world, rank = self.comm.world_size, self.comm.rank
        lo, hi = shard_bounds(f.padded, world, rank)
        n = hi - lo
        mc_g, mc_p = self._mc(gh), self._mc(ph)
        use_mc = bool(mc_g and mc_p)
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        gh.barrier(channel=0, timeout_ms=self.barrier_timeout_ms)                      # every rank's backward has finished writing its gradients
        w1.record()
        self.wait_events.append((w0, w1))
        native._count()
        native.ext().allreduce_adam(list(gh.buffer_ptrs), list(ph.buffer_ptrs), mc_g, mc_p, f.exp_avg, f.exp_avg_sq,
                                    lo, n, rank, hp["lr"], hp["beta1"], hp["beta2"], hp["eps"], hp["wd"], hp["step"],
                                    scale, use_mc, self.max_blocks, getattr(f, "master", None))
        ph.barrier(channel=1, timeout_ms=self.barrier_timeout_ms)                      # updated parameters are visible on every rank
        self.bytes_reduced += f.padded * f.grad.element_size()

    def pop_wait_ms(self) -> float:
        ms = 0.0
        for a, b in self.wait_events:
            b.synchronize()
            ms += a.elapsed_time(b)
        self.wait_events.clear()
        return ms

    @torch.no_grad()
    def allreduce_(self, t: torch.Tensor, scale: float = 1.0):
        """In-place fused sum all-reduce of a symmetric bf16 buffer (A/B baseline for dist.all_reduce)."""
        h = self.handle(t)
        world, rank = self.comm.world_size, self.comm.rank
        per = (t.numel() // world) // 8 * 8
        lo = rank * per
        n = per if rank < world - 1 else t.numel() - lo
        h.barrier(channel=0, timeout_ms=self.barrier_timeout_ms)
        native._count()
        native.ext().allreduce_sum(list(h.buffer_ptrs), lo, n, rank, scale, self.max_blocks)
        h.barrier(channel=1, timeout_ms=self.barrier_timeout_ms)
        return t
