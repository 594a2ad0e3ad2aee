"""Control-plane communicator.

The reference never calls ``torch.distributed`` directly -- accelerate does (SURVEY.md section 2.4,
call sites N1-N5).  Here one small ``Comm`` object owns the process group.  It is used for
rendezvous, barriers, the tiny packed metrics reduction (N4 -> one all-reduce instead of ten
all-gathers) and as the *baseline* implementation (``comm="nccl"``) of the two bandwidth paths;
the product paths are the fused kernels in ``parallel/fused_allreduce.py`` (K-AR) and
``parallel/weight_sync.py`` (K-BC) which run over symmetric memory.

Backends: ``nccl`` on CUDA, ``gloo`` on CPU (multi-process tests run under gloo on the dev box),
``single`` when WORLD_SIZE == 1.
"""
from __future__ import annotations

import datetime
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class Comm:
    def __init__(self, rank: int = 0, world_size: int = 1, local_rank: int = 0, device: Optional[torch.device] = None,
                 backend: str = "single", group=None):
        self.rank, self.world_size, self.local_rank = rank, world_size, local_rank
        self.device = device or torch.device("cpu")
        self.backend, self.group = backend, group

    # ---- construction ----------------------------------------------------------------------
    @classmethod
    def from_env(cls, device: Optional[torch.device] = None, timeout_s: float = 1800.0) -> "Comm":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if device is None:
            device = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
        if device.type == "cuda":
            torch.cuda.set_device(device)
        if world == 1:
            return cls(0, 1, local_rank, device, "single")
        backend = "nccl" if device.type == "cuda" else "gloo"
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            kw = {"device_id": device} if backend == "nccl" else {}
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=timeout_s), **kw)
        return cls(rank, world, local_rank, device, backend, dist.group.WORLD)

    @property
    def is_main(self) -> bool:
        return self.rank == 0

    # ---- collectives -----------------------------------------------------------------------
    def barrier(self):
        if self.world_size > 1:
            if self.backend == "nccl":
                dist.barrier(device_ids=[self.device.index])
            else:
                dist.barrier()

    def broadcast_object(self, obj, src: int = 0):
        if self.world_size == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    def all_reduce_(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
        if self.world_size > 1:
            ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
            if op == "mean":
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                t /= self.world_size
            else:
                dist.all_reduce(t, op=ops[op])
        return t

    def all_gather_cat(self, t: torch.Tensor) -> torch.Tensor:
        if self.world_size == 1:
            return t
        t = t.contiguous()
        if t.dim() == 0:
            t = t.reshape(1)
        # leading sizes may differ per rank (e.g. the number of micro-buckets under data-dependent batching):
        # agree on them first, pad to the longest, trim after the gather
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = [torch.empty_like(n) for _ in range(self.world_size)]
        dist.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        mx = max(sizes)
        if t.shape[0] < mx:
            t = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))], 0)
        out = [torch.empty_like(t) for _ in range(self.world_size)]
        dist.all_gather(out, t)
        return torch.cat([o[:k] for o, k in zip(out, sizes)], 0)

    def broadcast_(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.world_size > 1:
            dist.broadcast(t, src=src)
        return t

    def broadcast_module_(self, module, src: int = 0, only_frozen: bool = False) -> int:
        """Make a module's parameters and buffers identical to rank ``src``'s (what DDP's constructor does for everything
        it wraps).  Tensors are coalesced per dtype into flat staging buffers so a 1.5B model is a handful
        of NCCL broadcasts.  ``only_frozen`` skips trainable parameters (the optimizer broadcasts its flat buffers itself).
        Returns the number of bytes broadcast."""
        if self.world_size == 1 or module is None:
            return 0
        seen, by_dtype = set(), {}
        for t in list(module.parameters()) + list(module.buffers()):
            if id(t) in seen or t.numel() == 0 or (only_frozen and getattr(t, "requires_grad", False)):
                continue
            seen.add(id(t))
            by_dtype.setdefault((t.dtype, t.device), []).append(t)
        total = 0
        for (_dt, _dev), ts in by_dtype.items():
            chunk, n = [], 0
            for t in ts + [None]:
                if t is not None:
                    chunk.append(t)
                    n += t.numel()
                if chunk and (t is None or n >= (1 << 28)):          # <= ~0.5 GB (bf16) staging at a time
                    flat = torch.cat([c.detach().reshape(-1) for c in chunk])
                    dist.broadcast(flat, src=src)
                    off = 0
                    with torch.no_grad():
                        for c in chunk:
                            c.copy_(flat[off:off + c.numel()].view(c.shape))
                            off += c.numel()
                    total += flat.numel() * flat.element_size()
                    chunk, n = [], 0
        return total

    def reduce_scalars(self, values: Dict[str, float], op: str = "mean") -> Dict[str, float]:
        """Pack a dict of python floats into ONE vector and all-reduce it."""
        if self.world_size == 1:
            return dict(values)
        keys = sorted(values)
        t = torch.tensor([float(values[k]) for k in keys], dtype=torch.float64,
                         device=self.device if self.backend == "nccl" else "cpu")
        self.all_reduce_(t, op)
        return {k: float(v) for k, v in zip(keys, t.tolist())}

    def max_int(self, v: int) -> int:
        """Agree on a rank-invariant count (e.g. #micro-steps under data-dependent batching)."""
        if self.world_size == 1:
            return int(v)
        t = torch.tensor([int(v)], dtype=torch.int64, device=self.device if self.backend == "nccl" else "cpu")
        self.all_reduce_(t, "max")
        return int(t.item())

    def close(self):
        if self.world_size > 1 and dist.is_initialized():
            dist.destroy_process_group()
