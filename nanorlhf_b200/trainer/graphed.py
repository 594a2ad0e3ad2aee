"""CUDA-graph replay of one training micro-step (forward + loss + backward).

A 1.5B LoRA micro-step is ~3000 kernel launches; eager PyTorch needs ~110 ms of CPU time to issue
69 ms of GPU work, so the train phase is launch-bound.  The micro-step is pure device code once the
packing indices exist (``models.qwen2.build_logprob_plan`` holds every host sync), so it is captured
once per shape bucket and replayed:

* packed tokens are padded to a multiple of ``TOKEN_BUCKET`` with one dummy sequence of pad tokens,
  response rows to a multiple of ``ROW_BUCKET`` with entries that land in a dump row -- neither
  reaches the loss, so gradients are unchanged;
* gradients accumulate in place into the optimizer's flat buffer (static addresses), the
  per-micro-step statistics come back through one static vector;
* all graphs share one memory pool (they are replayed strictly one after another).

The first time a bucket is seen the step runs eagerly (that is also the warm-up the capture needs);
capture happens on the second occurrence.  The reference has no equivalent (HF ``Trainer``-style
eager loop, GRPO/grpo_trainer.py:560-640).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch

from ..models.qwen2 import build_logprob_plan, planned_response_logprobs
from ..utils import INVALID_LOGPROB

TOKEN_BUCKET = 256
ROW_BUCKET = 256
MAX_GRAPHS = 24


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Captured:
    __slots__ = ("graph", "plan", "mb", "out", "launches")


class GraphedMicroStep:
    def __init__(self, trainer):
        self.t = trainer
        self.seen: Dict[tuple, int] = {}
        self.graphs: "OrderedDict[tuple, _Captured]" = OrderedDict()
        self.pool = None
        self.stat_keys: Optional[list] = None
        self.replays = 0
        self.eager = 0
        self.disabled = False
        self.failed: Dict[tuple, int] = {}
        self.capture_failures = 0

    # ---- the device-only micro-step ---------------------------------------------------------------
    def _body(self, plan, mb, B, T_r, L, accum):
        t, a = self.t, self.t.args
        out_lp, out_ent, hidden = planned_response_logprobs(t.policy, plan, B, T_r, a.temperature, True,
                                                            max_seqlen=L)
        mb = dict(mb)
        mb["new_logprobs"] = torch.masked_fill(out_lp, mb["padding_mask"], INVALID_LOGPROB)
        if t.uses_value_model:                               # PPO: the critic's forward / backward rides in the same graph
            vm = t.model.value_model
            vh = hidden if vm is t.policy else vm.hidden_states(plan["ids"], plan["cu"], plan["pos"], L)
            vals = vm.values(vh.index_select(0, plan["vsel"]))
            out_v = torch.zeros((B + 1, T_r), dtype=torch.float32, device=vals.device)
            mb["vpred"] = out_v.index_put((plan["vr"], plan["vc"]), vals.float())[:B]     # padded entries land in dump row B
        loss, st = t.micro_loss(mb)
        (loss / accum).backward()
        with torch.no_grad():
            if a.stats_include_padding:
                st["entropy"] = out_ent.mean()
            else:
                m = (~mb["padding_mask"]).float()
                st["entropy"] = (out_ent * m).sum() / m.sum().clamp_min(1)
            if self.stat_keys is None:
                self.stat_keys = sorted(st)
            return torch.stack([st[k].detach().float().reshape(()) for k in self.stat_keys])

    # ---- padding to the bucket -----------------------------------------------------------------------
    @staticmethod
    def _pad_plan(plan, B, pad_id, T_b, R_b):
        dev = plan["ids"].device
        T, R = plan["ids"].numel(), plan["src"].numel()
        ids = torch.full((T_b,), pad_id, dtype=plan["ids"].dtype, device=dev)
        ids[:T] = plan["ids"]
        pos = torch.zeros(T_b, dtype=plan["pos"].dtype, device=dev)
        pos[:T] = plan["pos"]
        if T_b > T:
            pos[T:] = torch.arange(T_b - T, device=dev, dtype=pos.dtype)
        cu = torch.empty(B + 2, dtype=torch.int32, device=dev)
        cu[:B + 1] = plan["cu"]
        cu[B + 1] = T_b                                      # dummy sequence (possibly empty)
        src = torch.zeros(R_b, dtype=plan["src"].dtype, device=dev)
        src[:R] = plan["src"]
        tg = torch.zeros(R_b, dtype=plan["targets"].dtype, device=dev)
        tg[:R] = plan["targets"]
        r = torch.full((R_b,), B, dtype=plan["r"].dtype, device=dev)     # dump row
        r[:R] = plan["r"]
        c = torch.zeros(R_b, dtype=plan["c"].dtype, device=dev)
        c[:R] = plan["c"]
        return {"ids": ids, "cu": cu, "pos": pos, "src": src, "targets": tg, "r": r, "c": c}

    # ---- public ----------------------------------------------------------------------------------------
    def __call__(self, mb: dict, ctx: int, pad_id: int) -> torch.Tensor:
        """Runs the micro-step for ``mb`` (gradients accumulate into .grad); returns the stats vector
        (order ``self.stat_keys``)."""
        qr = mb["query_responses"]
        B, L = qr.shape
        T_r = L - ctx
        plan = build_logprob_plan(qr, ctx, pad_id)
        T_b = _round_up(plan["ids"].numel(), TOKEN_BUCKET)
        R_b = _round_up(max(plan["src"].numel(), 1), ROW_BUCKET)
        tens = {k: v for k, v in mb.items() if isinstance(v, torch.Tensor) and k != "query_responses"}
        accum = int(getattr(self.t, "_accum_steps", self.t.args.gradient_accumulation_steps))
        key = (T_b, R_b, B, L, accum, tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in tens.items())))
        padded = self._pad_plan(plan, B, pad_id, T_b, R_b)
        if self.t.uses_value_model:
            # value positions (columns ctx-1 .. L-2 of every real token, models/qwen2.py response_logprobs), padded to a bucket
            col, row = plan["col"], plan["row"]
            vsel = ((col >= ctx - 1) & (col <= L - 2)).nonzero(as_tuple=False).squeeze(1)
            nv = vsel.numel()
            V_b = _round_up(max(nv, 1), ROW_BUCKET)
            dev = vsel.device
            padded["vsel"] = torch.zeros(V_b, dtype=torch.long, device=dev)
            padded["vsel"][:nv] = vsel
            padded["vr"] = torch.full((V_b,), B, dtype=torch.long, device=dev)
            padded["vr"][:nv] = row[vsel]
            padded["vc"] = torch.zeros(V_b, dtype=torch.long, device=dev)
            padded["vc"][:nv] = col[vsel] + 1 - ctx
            key = key + (V_b,)
        cap = self.graphs.get(key)
        if cap is None:
            n = self.seen.get(key, 0)
            self.seen[key] = n + 1
            if n == 0 or self.disabled or len(self.graphs) >= MAX_GRAPHS or self.failed.get(key, 0) >= 2:
                self.eager += 1
                return self._body(padded, tens, B, T_r, L, accum)    # eager (also the capture warm-up)
            try:
                cap = self._capture(key, padded, tens, B, T_r, L, accum)
            except Exception as e:  # noqa: BLE001 -- a step that cannot be captured must still train
                import warnings
                # capture records, it does not execute: .grad is untouched.  A capture can be invalidated by something
                # transient (another thread's CUDA call), so the bucket gets one more try on its next occurrence; only
                # repeated failures turn the graphs off for good.
                self.failed[key] = self.failed.get(key, 0) + 1
                self.capture_failures += 1
                self.disabled = self.capture_failures >= 6
                warnings.warn(f"CUDA-graph capture of the micro-step failed ({type(e).__name__}: {e}); this step runs eagerly"
                              + ("; graphs are off from now on" if self.disabled else ""))
                self.eager += 1
                return self._body(padded, tens, B, T_r, L, accum)
        else:
            self.graphs.move_to_end(key)
        for k, v in padded.items():
            cap.plan[k].copy_(v)
        for k, v in tens.items():
            cap.mb[k].copy_(v)
        cap.graph.replay()
        self.replays += 1
        from ..ops import native
        native._count(cap.launches)
        return cap.out.clone()

    def _capture(self, key, padded, tens, B, T_r, L, accum) -> _Captured:
        from ..ops import native
        cap = _Captured()
        cap.plan = {k: v.clone() for k, v in padded.items()}
        cap.mb = {k: v.clone() for k, v in tens.items()}
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        torch.cuda.synchronize()
        before = native.launches()
        cap.graph = torch.cuda.CUDAGraph()
        # thread_local: CUDA calls of OTHER threads (checkpoint writer, pinned-memory prefetch) must not invalidate the capture;
        # the backward's launches from the autograd thread still land in the capturing stream
        with torch.cuda.graph(cap.graph, pool=self.pool, capture_error_mode="thread_local"):
            cap.out = self._body(cap.plan, cap.mb, B, T_r, L, accum)
        cap.launches = native.launches() - before
        native._count(-cap.launches)                # capture launched nothing; replays add it back
        self.graphs[key] = cap
        return cap
