"""HF-style checkpoint directories + real resume.

On-disk layout is the reference's (writer /root/reference/GRPO/grpo_trainer.py:
321-404, PPO/ppo_trainer.py:407-418):

    <output_dir>/checkpoint-<global_step>/
        adapter_config.json + adapter_model.safetensors      (use_lora)   | config.json + model.safetensors
        tokenizer files, training_args.bin, optimizer.pt, scheduler.pt,
        rng_state.pth | rng_state_<rank>.pth, trainer_state.json, value_model/ (PPO)

including the ``_old``-suffix quirk of the best-checkpoint bookkeeping (a metric whose name ends in
``_old`` describes the policy *before* this update, so the previous checkpoint is the one recorded
as best, :375-380) and rotation to ``save_total_limit`` that never deletes the best.  Unlike the
reference (which cannot resume: SURVEY.md 5.4) ``load_checkpoint`` restores model/adapter,
optimizer moments, scheduler, every RNG incl. the sampler-seed stream, dataloader position, episode.
"""
from __future__ import annotations

import os
import pickle
import random
import re
import shutil
from typing import Optional

import numpy as np
import torch

from ..sampler import engine as sampler_engine

PREFIX = "checkpoint"
_RE = re.compile(rf"^{PREFIX}-(\d+)$")


class AsyncCheckpointWriter:
    """Saving every update (the reference's ``save_steps=1``) must not stall the GPU: device tensors are snapshotted into
    reusable pinned host buffers by ``cudaMemcpyAsync`` on a side stream, and a background thread serialises them
    (safetensors / torch.save) once the copy event has fired -- the next update's rollout overlaps both.  The trainer
    calls ``wait_snapshot()`` before its next optimizer step (the only point where the snapshotted tensors change) and
    ``wait()`` before the next checkpoint / at the end of training."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.event = None
        self.thread = None
        self.error = None
        self._pinned = {}

    def snapshot(self, tensors: dict) -> dict:
        """Device -> pinned-host copies of ``tensors`` (dict of name -> tensor); returns host views valid after the event."""
        out = {}
        if not self.cuda:
            return {k: v.detach().clone() for k, v in tensors.items()}
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            for k, v in tensors.items():
                v = v.detach()
                if v.device.type != "cuda":
                    out[k] = v.clone()
                    continue
                buf = self._pinned.get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                    self._pinned[k] = buf
                buf.copy_(v, non_blocking=True)
                v.record_stream(self.stream)
                out[k] = buf
            self.event = torch.cuda.Event()
            self.event.record(self.stream)
        return out

    def submit(self, fn):
        import threading
        self.wait()
        ev = self.event

        def run():
            try:
                if ev is not None:
                    ev.synchronize()
                fn()
            except BaseException as e:  # noqa: BLE001 -- surfaced on the training thread by wait()
                self.error = e

        self.thread = threading.Thread(target=run, name="nanorlhf-ckpt-writer", daemon=False)
        self.thread.start()

    def wait_snapshot(self):
        if self.event is not None:
            self.event.synchronize()

    def wait(self):
        if self.thread is not None:
            self.thread.join()
            self.thread = None
        if self.error is not None:
            e, self.error = self.error, None
            raise RuntimeError("asynchronous checkpoint write failed") from e


def _model_tensors(model) -> dict:
    """The tensors ``save_pretrained`` would write: the peft adapter state dict, or the full state dict."""
    if hasattr(model, "adapter_state_dict"):
        return dict(model.adapter_state_dict())
    sd = dict(model.state_dict())
    if getattr(model.config, "tie_word_embeddings", False):
        sd.pop("lm_head.weight", None)
    return sd


def _save_one_model(model, path: str, host_state: Optional[dict] = None):
    """``host_state``: pre-fetched (pinned host) copies of ``_model_tensors(model)`` -- the async writer's path."""
    os.makedirs(path, exist_ok=True)
    if host_state is None:
        model.save_pretrained(path)
        return
    import json
    from dataclasses import asdict
    from ..models.hf_io import save_state_dict
    if hasattr(model, "adapter_state_dict"):
        cfg = asdict(model.peft_config)
        cfg["base_model_name_or_path"] = getattr(model, "name_or_path", "")
        with open(os.path.join(path, "adapter_config.json"), "w") as f:
            json.dump(cfg, f, indent=2)
        save_state_dict(host_state, os.path.join(path, "adapter_model.safetensors"))
    else:
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(model.config.to_dict(), f, indent=2)
        save_state_dict(host_state, os.path.join(path, "model.safetensors"))


def save_model(trainer, output_dir: str, host_states: Optional[dict] = None):
    """Policy only (adapter when LoRA), + ``value_model/`` for PPO, + tokenizer + training_args.bin."""
    if not trainer.comm.is_main:
        return
    hs = host_states or {}
    _save_one_model(trainer.policy, output_dir, hs.get("policy"))
    if trainer.uses_value_model and trainer.args.save_value_model and trainer.model.value_model is not None:
        _save_one_model(trainer.model.value_model, os.path.join(output_dir, "value_model"), hs.get("value"))
    if hasattr(trainer.tokenizer, "save_pretrained"):
        trainer.tokenizer.save_pretrained(output_dir)
    torch.save(trainer.args.to_dict(), os.path.join(output_dir, "training_args.bin"))      # torch.load-able like HF's


def _rng_state(device):
    st = {"python": random.getstate(), "numpy": np.random.get_state(), "cpu": torch.random.get_rng_state()}
    if device.type == "cuda":
        st["cuda"] = torch.cuda.random.get_rng_state(device)
    return st


def save_checkpoint(trainer, metrics=None):
    a, st, comm = trainer.args, trainer.state, trainer.comm
    folder = f"{PREFIX}-{st.global_step}"
    out = os.path.join(a.output_dir, folder)
    if comm.is_main:
        os.makedirs(out, exist_ok=True)
    comm.barrier()
    writer = getattr(trainer, "ckpt_writer", None)
    use_async = bool(getattr(a, "async_checkpoint", False)) and writer is not None and trainer.device.type == "cuda"
    if use_async:
        return _save_checkpoint_async(trainer, writer, out, metrics)
    save_model(trainer, out)
    if not a.save_only_model:
        # one world-size independent optimizer.pt (reference layout): under fused DP the ZeRO-1 shards of the moments /
        # fp32 master weights are gathered from their owners, so a run can resume on a different number of GPUs
        try:
            opt_sd = trainer.optimizer.state_dict(full=True)
        except TypeError:                       # a user-supplied torch optimizer
            opt_sd = trainer.optimizer.state_dict()
        if comm.is_main:
            torch.save(opt_sd, os.path.join(out, "optimizer.pt"))
        if comm.is_main:
            torch.save(trainer.lr_scheduler.state_dict(), os.path.join(out, "scheduler.pt"))
        rng = _rng_state(trainer.device)
        rng["sampler_seed_stream"] = sampler_engine.seed_stream_state()
        rng["trainer_np_rng"] = trainer._np_rng.get_state()
        rng["trainer_select_gen"] = trainer._select_gen.get_state()
        rng["dataloader"] = trainer.dataloader.state_dict()
        name = "rng_state.pth" if comm.world_size == 1 else f"rng_state_{comm.rank}.pth"
        torch.save(rng, os.path.join(out, name))
    # best-metric bookkeeping with the one-update lag for *_old metrics
    if metrics is not None and a.metric_for_best_model:
        name = a.metric_for_best_model
        if name in metrics:
            val = metrics[name]
            better = (lambda x, y: x > y) if a.greater_is_better else (lambda x, y: x < y)
            if st.best_metric is None or st.best_model_checkpoint is None or better(val, st.best_metric):
                st.best_metric = val
                if name.endswith("_old"):
                    prev = os.path.join(a.output_dir, f"{PREFIX}-{st.global_step - 1}")
                    st.best_model_checkpoint = prev if os.path.isdir(prev) else out
                else:
                    st.best_model_checkpoint = out
    if comm.is_main:
        st.save_to_json(os.path.join(out, "trainer_state.json"))
        rotate_checkpoints(a.output_dir, a.save_total_limit, st.best_model_checkpoint)
    comm.barrier()
    return out


def _update_best(trainer, out: str, metrics):
    a, st = trainer.args, trainer.state
    if metrics is not None and a.metric_for_best_model and a.metric_for_best_model in metrics:
        name, val = a.metric_for_best_model, metrics[a.metric_for_best_model]
        better = (lambda x, y: x > y) if a.greater_is_better else (lambda x, y: x < y)
        if st.best_metric is None or st.best_model_checkpoint is None or better(val, st.best_metric):
            st.best_metric = val
            prev = os.path.join(a.output_dir, f"{PREFIX}-{st.global_step - 1}")
            st.best_model_checkpoint = prev if (name.endswith("_old") and os.path.isdir(prev)) else out


def _save_checkpoint_async(trainer, writer: AsyncCheckpointWriter, out: str, metrics):
    """Same files as the synchronous path; the device -> host copies run on the writer's side stream and the serialisation
    in its thread.  Collective parts (gathering the ZeRO-1 optimizer shards) happen here, on every rank."""
    import copy
    import json  # noqa: F401
    a, st, comm = trainer.args, trainer.state, trainer.comm
    writer.wait()                                               # one checkpoint in flight at a time
    host = {}
    if comm.is_main:
        host["policy"] = writer.snapshot({f"p.{k}": v for k, v in _model_tensors(trainer.policy).items()})
        host["policy"] = {k[2:]: v for k, v in host["policy"].items()}
        if trainer.uses_value_model and a.save_value_model and trainer.model.value_model is not None:
            host["value"] = {k[2:]: v for k, v in writer.snapshot({f"v.{k}": v for k, v in _model_tensors(trainer.model.value_model).items()}).items()}
    opt_sd = None
    if not a.save_only_model:
        opt = trainer.optimizer
        if hasattr(opt, "state_tensors"):
            dev_state = {}
            for k, v in opt.state_tensors().items():           # gather shards on the device (collective), copy asynchronously
                f = opt._flats[int(k.split(".")[0][5:])]
                dev_state[k] = comm.all_gather_cat(v)[:f.padded] if getattr(opt, "comm_mode", "none") == "fused" else v
            if comm.is_main:
                host_state = {k[2:]: v for k, v in writer.snapshot({f"o.{k}": v for k, v in dev_state.items()}).items()}
                opt_sd = {"step": opt._step, "world": 1 if opt.comm_mode == "fused" else opt.world,
                          "comm_mode": "full" if opt.comm_mode == "fused" else opt.comm_mode,
                          "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in opt.param_groups], "state": host_state}
        elif comm.is_main:
            opt_sd = opt.state_dict()
    rng = _rng_state(trainer.device)
    rng["sampler_seed_stream"] = sampler_engine.seed_stream_state()
    rng["trainer_np_rng"] = trainer._np_rng.get_state()
    rng["trainer_select_gen"] = trainer._select_gen.get_state()
    rng["dataloader"] = trainer.dataloader.state_dict()
    sched_sd = copy.deepcopy(trainer.lr_scheduler.state_dict())
    _update_best(trainer, out, metrics)
    state_snapshot = copy.deepcopy(st)
    rank_rng_name = "rng_state.pth" if comm.world_size == 1 else f"rng_state_{comm.rank}.pth"

    def write():
        if not a.save_only_model:
            torch.save(rng, os.path.join(out, rank_rng_name))
        if not comm.is_main:
            return
        save_model(trainer, out, host)
        if not a.save_only_model:
            torch.save(opt_sd, os.path.join(out, "optimizer.pt"))
            torch.save(sched_sd, os.path.join(out, "scheduler.pt"))
        state_snapshot.save_to_json(os.path.join(out, "trainer_state.json"))     # written last: marks the checkpoint complete
        rotate_checkpoints(a.output_dir, a.save_total_limit, state_snapshot.best_model_checkpoint)

    writer.submit(write)
    return out


def list_checkpoints(output_dir: str):
    if not os.path.isdir(output_dir):
        return []
    found = []
    for d in os.listdir(output_dir):
        m = _RE.match(d)
        if m and os.path.isdir(os.path.join(output_dir, d)):
            found.append((int(m.group(1)), os.path.join(output_dir, d)))
    return [p for _, p in sorted(found)]


def rotate_checkpoints(output_dir: str, limit: Optional[int], best: Optional[str]):
    if not limit or limit <= 0:
        return
    cks = list_checkpoints(output_dir)
    if len(cks) <= limit:
        return
    # never delete the best: move it to just before the survivors (HF's rule)
    if best is not None and best in cks:
        cks.remove(best)
        cks.insert(max(len(cks) - limit + 1, 0), best)
    for p in cks[:max(0, len(cks) - limit)]:
        if p != best:
            shutil.rmtree(p, ignore_errors=True)


def find_resume_checkpoint(args) -> Optional[str]:
    cks = [c for c in list_checkpoints(args.output_dir) if os.path.exists(os.path.join(c, "trainer_state.json"))]
    return cks[-1] if cks else None


def _load_model_into(model, path: str):
    from ..models.hf_io import load_state_dict
    if hasattr(model, "load_adapter_state_dict") and os.path.exists(os.path.join(path, "adapter_model.safetensors")):
        model.load_adapter_state_dict(load_state_dict(os.path.join(path, "adapter_model.safetensors")))
        return
    sd = load_state_dict(os.path.join(path, "model.safetensors"))
    base = model
    if getattr(base.config, "tie_word_embeddings", False) and "lm_head.weight" not in sd and hasattr(base, "lm_head"):
        sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    with torch.no_grad():
        own = base.state_dict()
        for k, v in sd.items():
            if k in own:
                own[k].copy_(v.to(own[k].device, own[k].dtype))


def load_checkpoint(trainer, path: str):
    from ..utils.callbacks import OnlineTrainerState
    a, comm = trainer.args, trainer.comm
    _load_model_into(trainer.policy, path)
    vdir = os.path.join(path, "value_model")
    if trainer.uses_value_model and os.path.isdir(vdir):
        _load_model_into(trainer.model.value_model, vdir)
    opt_path = os.path.join(path, f"optimizer_rank{comm.rank}.pt")        # layout written by round-1 checkpoints
    if not os.path.exists(opt_path):
        opt_path = os.path.join(path, "optimizer.pt")
    if os.path.exists(opt_path):
        trainer.optimizer.load_state_dict(torch.load(opt_path, map_location="cpu", weights_only=False))
    sp = os.path.join(path, "scheduler.pt")
    if os.path.exists(sp):
        trainer.lr_scheduler.load_state_dict(torch.load(sp, map_location="cpu", weights_only=False))
    rp = os.path.join(path, "rng_state.pth" if comm.world_size == 1 else f"rng_state_{comm.rank}.pth")
    if os.path.exists(rp):
        rng = torch.load(rp, map_location="cpu", weights_only=False)
        random.setstate(rng["python"])
        np.random.set_state(rng["numpy"])
        torch.random.set_rng_state(rng["cpu"])
        if "cuda" in rng and trainer.device.type == "cuda":
            torch.cuda.random.set_rng_state(rng["cuda"], trainer.device)
        sampler_engine.set_seed_stream_state(rng["sampler_seed_stream"])
        trainer._np_rng.set_state(rng["trainer_np_rng"])
        trainer._select_gen.set_state(rng["trainer_select_gen"])
        trainer.dataloader.load_state_dict(rng["dataloader"])
    if hasattr(trainer, "_bump_policy_version"):
        trainer._bump_policy_version()            # a resident sampler must re-merge: its arena predates these weights
    old = OnlineTrainerState.load_from_json(os.path.join(path, "trainer_state.json"))
    for k in ("global_step", "episode", "epoch", "log_history", "best_metric", "best_model_checkpoint"):
        setattr(trainer.state, k, getattr(old, k))
