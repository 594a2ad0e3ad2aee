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


def _save_one_model(model, path: str):
    os.makedirs(path, exist_ok=True)
    model.save_pretrained(path)


def save_model(trainer, output_dir: str):
    """Policy only (adapter when LoRA), + ``value_model/`` for PPO, + tokenizer + training_args.bin."""
    if not trainer.comm.is_main:
        return
    _save_one_model(trainer.policy, output_dir)
    if trainer.uses_value_model and trainer.args.save_value_model and trainer.model.value_model is not None:
        _save_one_model(trainer.model.value_model, os.path.join(output_dir, "value_model"))
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
