"""Trainer state, control flags and callback plumbing.

The reference inherits these from HF ``Trainer`` (``TrainerState``/``OnlineTrainerState``,
``TrainerControl``, ``CallbackHandler``, ``DefaultFlowCallback``, ``EarlyStoppingCallback``:
import sites /root/reference/GRPO/grpo_trainer.py:49-59, use :266-273,:462,:749-752).
Events fired by the trainers: on_train_begin, on_step_end, on_log, on_save, on_evaluate,
on_train_end -- the same set the reference reaches.
"""
from __future__ import annotations

import dataclasses
import json
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional


@dataclass
class OnlineTrainerState:
    """HF ``TrainerState`` subset + ``episode`` (trl's OnlineTrainerState)."""
    epoch: float = 0.0
    global_step: int = 0
    max_steps: int = 0
    episode: int = 0
    num_train_epochs: float = 0.0
    logging_steps: int = 1
    eval_steps: int = 1
    save_steps: int = 1
    log_history: List[Dict[str, Any]] = field(default_factory=list)
    best_metric: Optional[float] = None
    best_model_checkpoint: Optional[str] = None
    is_local_process_zero: bool = True
    is_world_process_zero: bool = True
    stateful_callbacks: Dict[str, Any] = field(default_factory=dict)

    def save_to_json(self, path: str):
        with open(path, "w") as f:
            json.dump(dataclasses.asdict(self), f, indent=2, sort_keys=True, default=str)

    @classmethod
    def load_from_json(cls, path: str):
        with open(path) as f:
            d = json.load(f)
        return cls(**{k: v for k, v in d.items() if k in cls.__dataclass_fields__})


@dataclass
class TrainerControl:
    should_training_stop: bool = False
    should_epoch_stop: bool = False
    should_save: bool = False
    should_evaluate: bool = False
    should_log: bool = False


class TrainerCallback:
    def on_train_begin(self, args, state, control, **kw): ...
    def on_train_end(self, args, state, control, **kw): ...
    def on_step_end(self, args, state, control, **kw): ...
    def on_log(self, args, state, control, logs=None, **kw): ...
    def on_save(self, args, state, control, **kw): ...
    def on_evaluate(self, args, state, control, metrics=None, **kw): ...


class DefaultFlowCallback(TrainerCallback):
    """Sets should_log / should_save / should_evaluate from the *_steps fields."""

    def on_step_end(self, args, state, control, **kw):
        if args.logging_steps and state.global_step % state.logging_steps == 0:
            control.should_log = True
        if args.save_strategy == "steps" and state.save_steps and state.global_step % state.save_steps == 0:
            control.should_save = True
        if args.eval_strategy == "steps" and state.eval_steps and state.global_step % state.eval_steps == 0:
            control.should_evaluate = True
        if state.max_steps and state.global_step >= state.max_steps:
            control.should_training_stop = True
            if args.save_strategy != "no":
                control.should_save = True
        return control

    def on_train_end(self, args, state, control, **kw):
        return control


class ProgressCallback(TrainerCallback):
    def __init__(self):
        self.t0 = None

    def on_train_begin(self, args, state, control, **kw):
        self.t0 = time.time()

    def on_log(self, args, state, control, logs=None, **kw):
        if state.is_world_process_zero and logs is not None and not getattr(args, "quiet", False):
            shown = {k: (round(v, 5) if isinstance(v, float) else v) for k, v in logs.items()}
            print(shown, flush=True)


class EarlyStoppingCallback(TrainerCallback):
    """Counts non-improving evaluations; acts only in ``on_evaluate`` (so it is inert in the
    reference's RL loops, which never fire it -- SURVEY.md section 2.2 -- but live in the value
    pre-fit, PPO/ppo.py:84-110)."""

    def __init__(self, early_stopping_patience: int = 1, early_stopping_threshold: float = 0.0):
        self.patience, self.threshold, self.counter = early_stopping_patience, early_stopping_threshold, 0

    def on_train_begin(self, args, state, control, **kw):
        if getattr(args, "metric_for_best_model", None) is None:
            raise AssertionError("EarlyStoppingCallback requires metric_for_best_model")

    def check(self, args, state, control, value: float):
        better = (lambda a, b: a > b) if args.greater_is_better else (lambda a, b: a < b)
        if state.best_metric is None or (better(value, state.best_metric)
                                         and abs(value - state.best_metric) > self.threshold):
            self.counter = 0
        else:
            self.counter += 1

    def on_evaluate(self, args, state, control, metrics=None, **kw):
        name = args.metric_for_best_model
        if metrics is None or name not in metrics:
            return control
        self.check(args, state, control, metrics[name])
        if self.counter >= self.patience:
            control.should_training_stop = True
        return control

    def state(self):
        return {"args": {"early_stopping_patience": self.patience, "early_stopping_threshold": self.threshold},
                "attributes": {"early_stopping_patience_counter": self.counter}}


class CallbackHandler:
    def __init__(self, callbacks, model=None, processing_class=None, optimizer=None, lr_scheduler=None):
        self.callbacks = list(callbacks)
        self.model, self.processing_class = model, processing_class
        self.optimizer, self.lr_scheduler = optimizer, lr_scheduler

    def add_callback(self, cb):
        self.callbacks.append(cb() if isinstance(cb, type) else cb)

    def _fire(self, event, args, state, control, **kw):
        for cb in self.callbacks:
            r = getattr(cb, event)(args, state, control, model=self.model, processing_class=self.processing_class,
                                   optimizer=self.optimizer, lr_scheduler=self.lr_scheduler, **kw)
            if r is not None:
                control = r
        return control

    def on_train_begin(self, args, state, control):
        control.should_training_stop = False
        return self._fire("on_train_begin", args, state, control)

    def on_train_end(self, args, state, control):
        return self._fire("on_train_end", args, state, control)

    def on_step_end(self, args, state, control):
        control.should_log = control.should_save = control.should_evaluate = False
        return self._fire("on_step_end", args, state, control)

    def on_log(self, args, state, control, logs):
        control.should_log = False
        return self._fire("on_log", args, state, control, logs=logs)

    def on_save(self, args, state, control):
        control.should_save = False
        return self._fire("on_save", args, state, control)

    def on_evaluate(self, args, state, control, metrics):
        control.should_evaluate = False
        return self._fire("on_evaluate", args, state, control, metrics=metrics)


DEFAULT_CALLBACKS = [DefaultFlowCallback]
