"""Metric sinks: stdout, JSONL (always), tensorboard, wandb (offline-capable).

Reference: ``self.log(metrics)`` -> HF ``on_log`` -> Wandb/TensorBoard callbacks per ``report_to``
(/root/reference/GRPO/grpo_trainer.py:747, GRPO/grpo.py:136); completions table printed with rich
and logged to wandb (:712-724).  Keys are listed in SURVEY.md App. B.
"""
from __future__ import annotations

import json
import os
import time
from typing import Any, Dict, Iterable, List, Optional

from .callbacks import TrainerCallback


def _as_list(report_to) -> List[str]:
    if report_to in (None, "none", "None", []):
        return []
    if isinstance(report_to, str):
        return [s.strip() for s in report_to.split(",") if s.strip() and s.strip() != "none"]
    return list(report_to)


class JsonlLogger(TrainerCallback):
    def __init__(self, path: str):
        self.path = path
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)

    def on_log(self, args, state, control, logs=None, **kw):
        if state.is_world_process_zero and logs is not None:
            with open(self.path, "a") as f:
                f.write(json.dumps({"time": time.time(), **logs}, default=float) + "\n")


class TensorBoardLogger(TrainerCallback):
    def __init__(self, logdir: str):
        self.logdir, self.w = logdir, None

    def on_train_begin(self, args, state, control, **kw):
        if state.is_world_process_zero and self.w is None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.w = SummaryWriter(self.logdir)
            except Exception as e:  # tensorboard is optional
                print(f"[metrics] tensorboard unavailable: {e}")

    def on_log(self, args, state, control, logs=None, **kw):
        if self.w is not None and logs:
            for k, v in logs.items():
                if isinstance(v, (int, float)):
                    self.w.add_scalar(k, v, state.global_step)
            self.w.flush()

    def on_train_end(self, args, state, control, **kw):
        if self.w is not None:
            self.w.close()
            self.w = None


class WandbLogger(TrainerCallback):
    """wandb in offline mode when there is no network (GPU box)."""

    def __init__(self, run_name: Optional[str] = None, config: Optional[dict] = None):
        self.run_name, self.config, self.run = run_name, config, None

    def on_train_begin(self, args, state, control, **kw):
        if not state.is_world_process_zero or self.run is not None:
            return
        try:
            import wandb
            os.environ.setdefault("WANDB_MODE", "offline")
            self.run = wandb.init(project=os.environ.get("WANDB_PROJECT", "nanorlhf"), name=self.run_name,
                                  config=self.config, dir=args.logging_dir, reinit=True)
        except Exception as e:
            print(f"[metrics] wandb unavailable: {e}")

    def on_log(self, args, state, control, logs=None, **kw):
        if self.run is not None and logs:
            self.run.log({k: v for k, v in logs.items() if isinstance(v, (int, float))}, step=state.global_step)

    def log_table(self, name: str, columns: List[str], rows: Iterable[Iterable[Any]]):
        if self.run is not None:
            import wandb
            self.run.log({name: wandb.Table(columns=columns, data=[list(r) for r in rows])})

    def on_train_end(self, args, state, control, **kw):
        if self.run is not None:
            self.run.finish()
            self.run = None


def reporting_callbacks(args, run_name: Optional[str] = None) -> List[TrainerCallback]:
    cbs: List[TrainerCallback] = [JsonlLogger(os.path.join(args.logging_dir, "metrics.jsonl"))]
    for name in _as_list(args.report_to):
        if name == "tensorboard":
            cbs.append(TensorBoardLogger(args.logging_dir))
        elif name == "wandb":
            cbs.append(WandbLogger(run_name, args.to_dict() if hasattr(args, "to_dict") else None))
        elif name in ("jsonl", "stdout"):
            pass
        else:
            raise ValueError(f"unknown report_to sink {name!r}")
    return cbs


def print_rich_table(columns: List[str], rows: List[List[Any]], title: str = "completions", max_rows: int = 5):
    """Console table of the first completions (reference: print_rich_table, grpo_trainer.py:720)."""
    try:
        from rich.console import Console
        from rich.table import Table
        t = Table(title=title, show_lines=True)
        for c in columns:
            t.add_column(c, overflow="fold")
        for r in rows[:max_rows]:
            t.add_row(*[str(x)[:400] for x in r])
        Console().print(t)
    except Exception:
        for r in rows[:max_rows]:
            print(dict(zip(columns, r)))
