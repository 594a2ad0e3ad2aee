"""LR schedules (the reference gets these from HF ``get_scheduler``).

``cosine_with_min_lr`` with ``min_lr_rate`` is the default of every entry script
(/root/reference/GRPO/grpo.py:119-121); ``reduce_lr_on_plateau`` is used by the value pre-fit
(/root/reference/PPO/ppo.py:92-94).  Schedulers are stepped once per *update* with horizon
``num_total_batches`` (/root/reference/GRPO/grpo_trainer.py:258-260,748).
"""
from __future__ import annotations

import math

import torch


def _lambda(kind: str, warmup: int, total: int, kw: dict):
    min_rate = float(kw.get("min_lr_rate", 0.0))
    cycles = float(kw.get("num_cycles", 0.5))

    def f(step: int) -> float:
        if warmup > 0 and step < warmup:
            return step / max(1, warmup)
        if kind == "constant":
            return 1.0
        prog = (step - warmup) / max(1, total - warmup)
        prog = min(max(prog, 0.0), 1.0)
        if kind == "linear":
            return max(0.0, 1.0 - prog)
        if kind == "cosine":
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * cycles * 2.0 * prog)))
        if kind == "cosine_with_min_lr":
            factor = 0.5 * (1.0 + math.cos(math.pi * cycles * 2.0 * prog))
            return factor * (1 - min_rate) + min_rate
        raise ValueError(f"unknown lr_scheduler_type {kind!r}")
    return f


def get_scheduler(kind: str, optimizer, num_warmup_steps: int, num_training_steps: int, scheduler_specific_kwargs=None):
    kw = dict(scheduler_specific_kwargs or {})
    if kind == "reduce_lr_on_plateau":
        return torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, **kw)
    if kind == "constant_with_warmup":
        kind = "constant"
    return torch.optim.lr_scheduler.LambdaLR(optimizer, _lambda(kind, num_warmup_steps, num_training_steps, kw))
