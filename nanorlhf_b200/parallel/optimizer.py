"""Flat-buffer AdamW with the data-parallel gradient reduction folded into ``step()``.

Reference: HF ``Trainer.create_optimizer`` -> ``torch.optim.AdamW`` over ``requires_grad`` params
(/root/reference/GRPO/grpo_trainer.py:258), DDP's bucketed NCCL all-reduce on the accumulation
boundary (:690) and a separate AdamW step (:692, K18); PPO overrides
``create_optimizer`` for two learning rates (/root/reference/PPO/ppo_trainer.py:341-402).

B200-first design:

* every param group lives in ONE contiguous flat parameter buffer, ONE flat gradient buffer
  (``p.grad`` are views, autograd accumulates in place) and flat moment buffers, so the whole
  optimizer step is a single kernel launch over ~0.5 G elements instead of hundreds;
* with ``world_size > 1`` and ``comm="fused"`` the flat gradient buffer is allocated in symmetric
  memory and ``step()`` launches K-AR (parallel/fused_allreduce.py): each rank reduces its 1/N
  slice straight out of its peers' gradient buffers over NVLink (or NVLS ``multimem.ld_reduce``),
  applies AdamW to that slice in registers and writes the updated bf16 parameters into every
  peer's parameter buffer -- reduce-scatter + Adam + all-gather in one kernel, moments sharded
  ZeRO-1 style.  ``comm="nccl"`` keeps the baseline: ``dist.all_reduce`` then a local step.
* moments are fp32 by default (``optimizer_state_dtype="bf16"`` reproduces the reference's
  bf16-moment behaviour);
* bf16 parameters get an **fp32 master copy of the owned shard** next to the moments (``master_weights``): at
  the shipped lr = 6e-6 an Adam step (~lr) is far below half a bf16 ulp of a typical weight (|w| ~ 0.02 ->
  ulp 1.2e-4), so without it ``round(p - lr * u)`` == p and training silently stalls.  peft keeps adapter
  weights in fp32 for the same reason; here the model keeps bf16 compute copies and the optimizer owns the
  fp32 truth (sharded 1/N per rank under fused DP, so it costs 4 bytes x 540 M / N).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from .. import ops
from ..ops import reference as ref


class _Flat:
    """One param group's flat storage."""

    def __init__(self, params: List[torch.nn.Parameter], state_dtype: torch.dtype, grad_alloc=None,
                 param_alloc=None, shard: Optional[range] = None):
        self.params = params
        self.dtype = params[0].dtype
        self.device = params[0].device
        sizes = [p.numel() for p in params]
        pad = 0
        self.numel = sum(sizes)
        # round up so every rank's shard is 16-byte aligned for vectorised kernels
        self.padded = (self.numel + 1023) // 1024 * 1024
        alloc_p = param_alloc or (lambda n, dt, dev: torch.zeros(n, dtype=dt, device=dev))
        alloc_g = grad_alloc or (lambda n, dt, dev: torch.zeros(n, dtype=dt, device=dev))
        self.param = alloc_p(self.padded, self.dtype, self.device)
        self.grad = alloc_g(self.padded, self.dtype, self.device)
        off = 0
        self.offsets = []
        with torch.no_grad():
            for p, n in zip(params, sizes):
                self.param[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.param[off:off + n].view(p.shape)
                p.grad = self.grad[off:off + n].view(p.shape)
                self.offsets.append(off)
                off += n
        self.state_dtype = state_dtype
        self.exp_avg = None
        self.exp_avg_sq = None
        self.master = None
        self.shard = shard

    def ensure_state(self, lo: int, hi: int, master: bool = False):
        if self.exp_avg is None:
            self.exp_avg = torch.zeros(hi - lo, dtype=self.state_dtype, device=self.device)
            self.exp_avg_sq = torch.zeros(hi - lo, dtype=self.state_dtype, device=self.device)
        if master and self.master is None and self.dtype != torch.float32:
            self.master = self.param[lo:hi].float()

    def rebind_grads(self):
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad[off:off + p.numel()].data_ptr():
                p.grad = self.grad[off:off + p.numel()].view(p.shape)


class FusedAdamW(torch.optim.Optimizer):
    """AdamW over flat buffers; gradient averaging across data-parallel ranks happens inside step()."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 state_dtype: torch.dtype = torch.float32, comm=None, comm_mode: str = "fused",
                 master_weights: bool = True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.comm = comm
        self.world = comm.world_size if comm is not None else 1
        self.rank = comm.rank if comm is not None else 0
        self.comm_mode = comm_mode if self.world > 1 else "none"
        if self.comm_mode == "fused" and self.world not in (2, 4, 8):
            self.comm_mode = "nccl"          # csrc/comm.cu instantiates the P2P kernel for 2 / 4 / 8 ranks
        self.state_dtype = state_dtype
        self.master_weights = master_weights
        self.max_grad_norm: Optional[float] = None      # set by the trainer (config.max_grad_norm)
        self.last_grad_norm: Optional[float] = None
        self.comm_events: List[tuple] = []      # (start, end) CUDA events around every collective of step()
        self._flats: List[_Flat] = []
        self._fused = None
        self._step = 0
        self.grad_scale = 1.0            # extra multiplier applied to gradients inside the kernel
        # comm="nccl" bucketed overlap (enable_bucket_overlap): per flat a list of [lo, hi, params pending, launched]
        self._buckets: Optional[List[List[list]]] = None
        self._armed = False
        self._inflight: list = []                # async work handles of the buckets launched from hooks
        self._order: List[tuple] = []            # global launch sequence of (flat index, bucket index)
        self._next = 0
        self.overlap_launched = 0                # buckets whose all-reduce was issued from a backward hook (statistics)
        grad_alloc = param_alloc = None
        if self.comm_mode == "fused":
            from .fused_allreduce import FusedAllReduceAdam
            self._fused = FusedAllReduceAdam(comm)
            grad_alloc = self._fused.alloc
            param_alloc = self._fused.alloc
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            if not ps:
                self._flats.append(None)
                continue
            if len({p.dtype for p in ps}) != 1 or len({p.device for p in ps}) != 1:
                raise ValueError("all params of one group must share dtype and device")
            self._flats.append(_Flat(ps, state_dtype, grad_alloc, param_alloc))
        if self._fused is not None:
            self._fused.register([f for f in self._flats if f is not None])
        if self.world > 1:
            # replicas must start identical (DDP broadcasts rank 0's parameters at construction);
            # one flat broadcast per group over the bootstrap process group
            for f in self.flats:
                comm.broadcast_(f.param, 0)

    # ---- comm="nccl": DDP-style bucketed all-reduce overlapped with the backward -----------------
    def enable_bucket_overlap(self, bucket_bytes: int = 25 << 20) -> bool:
        """What DDP's reducer does for the reference (/root/reference/GRPO/grpo_trainer.py:690): gradients are
        grouped into ~25 MB buckets in reverse parameter order (the order the backward produces them); on the micro-step the
        trainer arms with ``arm_overlap()`` a post-accumulate hook per parameter counts its bucket down and the bucket's
        all-reduce is issued asynchronously the moment it is complete, while the backward of the earlier layers still runs.
        ``step()`` waits for the handles and reduces whatever was not launched (parameters without a gradient, eager
        fall-backs).  Buckets are contiguous ranges of the flat gradient buffer, so there is no copy in or out.
        Only the NCCL / gloo path: K-AR (comm="fused") reduces inside the optimizer kernel; CUDA-graph replays fire no hooks
        (then nothing is launched early and ``step()`` reduces everything, as before)."""
        if self.comm_mode != "nccl" or self._buckets is not None:
            return self._buckets is not None
        self._buckets = []
        for fi, f in enumerate(self.flats):
            esz = f.grad.element_size()
            buckets, hi, members, nbytes = [], f.numel, [], 0
            for pi in range(len(f.params) - 1, -1, -1):
                members.append(pi)
                nbytes += f.params[pi].numel() * esz
                if nbytes >= bucket_bytes or pi == 0:
                    lo = f.offsets[pi]
                    buckets.append([lo, hi, len(members), False, list(members)])
                    hi, members, nbytes = lo, [], 0
            self._buckets.append(buckets)
            for bi, b in enumerate(buckets):
                for pi in b[4]:
                    f.params[pi].register_post_accumulate_grad_hook(self._make_hook(fi, bi))
        self._order = [(fi, bi) for fi, buckets in enumerate(self._buckets) for bi in range(len(buckets))]
        self._next = len(self._order)
        return True

    def _make_hook(self, fi: int, bi: int):
        def hook(_param):
            if not self._armed:
                return
            b = self._buckets[fi][bi]
            b[2] -= 1
            if b[2] == 0:
                self._launch_ready()
        return hook

    def _launch_ready(self):
        """Issue the all-reduces strictly in the global bucket sequence (flat by flat, last parameters first), whatever order
        the hooks complete in: every rank -- including one whose window had no micro-step at all and reduces everything inside
        ``step()`` -- then enqueues the same collectives in the same order."""
        import torch.distributed as dist
        while self._next < len(self._order):
            fi, bi = self._order[self._next]
            b = self._buckets[fi][bi]
            if b[2] != 0:
                break
            f = self.flats[fi]
            self._inflight.append(dist.all_reduce(f.grad[b[0]:b[1]], op=dist.ReduceOp.SUM, async_op=True))
            b[3] = True
            self._next += 1
            self.overlap_launched += 1

    def arm_overlap(self):
        """Call right before the LAST micro-step's backward of an accumulation window (earlier micro-steps must not reduce:
        DDP's ``no_sync``)."""
        if self._buckets is None:
            return
        self._armed = True
        self._next = 0
        for buckets in self._buckets:
            for b in buckets:
                b[2], b[3] = len(b[4]), False

    def _finish_overlap(self) -> bool:
        """Wait for the early all-reduces and reduce, in sequence, the buckets that were not launched (parameters without a
        gradient this step, CUDA-graph replays, a window without micro-steps).  True if the gradients are now reduced."""
        if self._buckets is None:
            return False
        import torch.distributed as dist
        if not self._armed:                       # nothing armed this window: the whole sequence is reduced here
            self._next = 0
        self._armed = False
        for work in self._inflight:
            work.wait()
        self._inflight.clear()
        for fi, bi in self._order[self._next:]:
            b = self._buckets[fi][bi]
            dist.all_reduce(self.flats[fi].grad[b[0]:b[1]], op=dist.ReduceOp.SUM)
        self._next = len(self._order)
        return True

    # ---- bookkeeping -----------------------------------------------------------------------
    @property
    def flats(self):
        return [f for f in self._flats if f is not None]

    def zero_grad(self, set_to_none: bool = False):
        for f in self.flats:
            f.grad.zero_()
            f.rebind_grads()

    def _shard_bounds(self, f: _Flat):
        """Slice of the flat buffer whose moments this rank owns (whole buffer unless fused DP)."""
        if self.comm_mode != "fused":
            return 0, f.padded
        return shard_bounds(f.padded, self.world, self.rank)

    def _clip_factor(self, already_reduced: bool = False) -> float:
        """Global-norm clipping needs the norm of the REDUCED gradient before any parameter moves, so the gradients are
        all-reduced first (NCCL) and this step then runs the local update on the already reduced buffers."""
        if self.world > 1 and not already_reduced:
            for f in self.flats:
                self.comm.all_reduce_(f.grad, "sum")
        norm = float(self.grad_norm()) * self.grad_scale / self.world
        self.last_grad_norm = norm
        return min(1.0, self.max_grad_norm / (norm + 1e-6))

    @torch.no_grad()
    def step(self, closure=None):
        self._step += 1
        clip, reduced = 1.0, False
        if self.comm_mode == "nccl" and self._finish_overlap():
            reduced = True                       # every bucket was all-reduced (most of them during the backward)
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            clip, reduced = self._clip_factor(already_reduced=reduced), self.world > 1
        for g, f in zip(self.param_groups, self._flats):
            if f is None:
                continue
            lo, hi = self._shard_bounds(f)
            f.ensure_state(lo, hi, self.master_weights)
            b1, b2 = g["betas"]
            hp = dict(lr=float(g["lr"]), beta1=b1, beta2=b2, eps=g["eps"], wd=g["weight_decay"], step=self._step)
            timed = self.world > 1 and f.param.is_cuda
            if timed:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            if self.comm_mode == "fused" and not reduced:
                self._fused.allreduce_adam(f, hp, clip * self.grad_scale / self.world)
            elif self.comm_mode == "fused":
                # clipped step: gradients already reduced; every rank updates its shard and the bf16 copies are re-gathered
                lo_, hi_ = lo, hi
                scale = clip * self.grad_scale / self.world
                if ops.use_native(f.param):
                    ops._nat().adamw_flat(f.param[lo_:hi_], f.grad[lo_:hi_], f.exp_avg, f.exp_avg_sq, scale=scale, master=f.master, **hp)
                else:
                    ref.adamw_step_(f.param[lo_:hi_], f.grad[lo_:hi_], f.exp_avg, f.exp_avg_sq, hp["lr"], b1, b2, hp["eps"], hp["wd"],
                                    self._step, grad_scale=scale, master=f.master)
                shards = self.comm.all_gather_cat(f.param[lo_:hi_].clone()) if (hi_ - lo_) * self.world == f.padded else None
                if shards is not None:
                    f.param.copy_(shards)
                else:                                       # ragged last shard: broadcast each owner's slice
                    for r in range(self.world):
                        a_, b_ = shard_bounds(f.padded, self.world, r)
                        self.comm.broadcast_(f.param[a_:b_], r)
            else:
                if self.comm_mode == "nccl":
                    if not reduced:
                        self.comm.all_reduce_(f.grad, "sum")
                    scale = clip * self.grad_scale / self.world
                    if timed:                   # baseline: only the all-reduce is communication
                        ev1.record()
                        self.comm_events.append((ev0, ev1))
                        timed = False
                else:
                    scale = clip * self.grad_scale
                if ops.use_native(f.param):
                    ops._nat().adamw_flat(f.param, f.grad, f.exp_avg, f.exp_avg_sq, scale=scale, master=f.master, **hp)
                else:
                    ref.adamw_step_(f.param, f.grad, f.exp_avg, f.exp_avg_sq, hp["lr"], b1, b2, hp["eps"], hp["wd"],
                                    self._step, grad_scale=scale, master=f.master)
            if timed:
                ev1.record()
                self.comm_events.append((ev0, ev1))
        return None

    def pop_comm_ms(self, split: bool = False):
        """Device time spent in step()'s collectives since the last call.  For K-AR that is the opening barrier (= waiting
        for the slowest rank to finish its backward: load imbalance, not communication), the fused kernel and the closing
        barrier; nothing overlaps it, so all of it is exposed.  ``split=True`` returns (communication, straggler wait).
        Synchronises on the recorded events."""
        ms = 0.0
        for a, b in self.comm_events:
            b.synchronize()
            ms += a.elapsed_time(b)
        self.comm_events.clear()
        wait = self._fused.pop_wait_ms() if self._fused is not None else 0.0
        return (ms - wait, wait) if split else ms

    def grad_norm(self) -> torch.Tensor:
        sq = sum((f.grad.float() ** 2).sum() for f in self.flats)
        return torch.sqrt(sq)

    # ---- state (checkpoint / offload) --------------------------------------------------------
    def state_tensors(self) -> Dict[str, torch.Tensor]:
        out = {}
        for i, f in enumerate(self._flats):
            if f is not None and f.exp_avg is not None:
                out[f"group{i}.exp_avg"] = f.exp_avg
                out[f"group{i}.exp_avg_sq"] = f.exp_avg_sq
                if f.master is not None:
                    out[f"group{i}.master"] = f.master
        return out

    def set_state_tensor(self, key: str, t: torch.Tensor):
        gi, name = key.split(".")
        setattr(self._flats[int(gi[5:])], name, t)

    def state_dict(self, full: bool = False):
        """``full=False``: this rank's shard (cheap, what the offload engine moves).  ``full=True``: the whole-buffer
        state gathered from the owners -- the world-size independent form written to ``optimizer.pt``; collective."""
        state = {k: v.detach() for k, v in self.state_tensors().items()}
        layout = {"world": self.world, "comm_mode": self.comm_mode}
        if full and self.comm_mode == "fused":
            gathered = {}
            for k, v in state.items():
                f = self._flats[int(k.split(".")[0][5:])]
                gathered[k] = self.comm.all_gather_cat(v)[:f.padded]
            state, layout = gathered, {"world": 1, "comm_mode": "full"}
        return {"step": self._step, **layout,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "state": {k: v.cpu() for k, v in state.items()}}

    def load_state_dict(self, sd):
        self._step = sd["step"]
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in saved.items() if k in ("lr", "betas", "eps", "weight_decay", "initial_lr")})
        same_layout = sd.get("world", 1) == self.world and sd.get("comm_mode", "none") == self.comm_mode
        for k, v in sd["state"].items():
            gi, name = int(k.split(".")[0][5:]), k.split(".")[1]
            f = self._flats[gi]
            lo, hi = self._shard_bounds(f)
            f.ensure_state(lo, hi, self.master_weights and name == "master")
            dst = getattr(f, name)
            if dst is None:
                continue
            if v.numel() == f.padded:                       # whole-buffer state: take the slice this rank owns
                dst.copy_(v[lo:hi].to(f.device))
            elif same_layout and v.numel() == dst.numel():
                dst.copy_(v.to(f.device))
            else:
                raise RuntimeError(f"optimizer state '{k}' has {v.numel()} elements: a per-rank shard written with a "
                                   f"different DP layout (world {sd.get('world')}); save with state_dict(full=True) to reshard")
        for f in self.flats:                                # the bf16 parameters are the rounded image of the master copy
            if f.master is not None and any(k.endswith(".master") for k in sd["state"]):
                lo, hi = self._shard_bounds(f)
                f.param[lo:hi].copy_(f.master.to(f.dtype))


def shard_bounds(padded: int, world: int, rank: int):
    """[lo, hi) of the flat buffer owned by ``rank``: equal slices rounded down to 8 elements (the kernels move
    16-byte vectors, any world size), the last rank takes the remainder."""
    per = (padded // world) // 8 * 8
    lo = rank * per
    return lo, (lo + per if rank < world - 1 else padded)


def build_param_groups(named_params: Iterable, weight_decay: float, lr: float):
    """Decay / no-decay split the way HF does it (no decay for biases and norm weights)."""
    decay, no_decay = [], []
    for n, p in named_params:
        if not p.requires_grad:
            continue
        (no_decay if (n.endswith("bias") or "norm" in n.lower()) else decay).append(p)
    groups = []
    if decay:
        groups.append({"params": decay, "weight_decay": weight_decay, "lr": lr})
    if no_decay:
        groups.append({"params": no_decay, "weight_decay": 0.0, "lr": lr})
    return groups
