"""The shared online-RL training loop.

The reference ships seven near-identical copies of this loop, one per algorithm
(/root/reference/GRPO/grpo_trainer.py:176-778 and siblings), each subclassing HF ``Trainer``.
Here there is one ``RLTrainer`` that owns the six phases of an update:

  1. tier optimizer state out of the way (optional)        -- runtime/offload.py   (K-OFF)
  2. rollout with the in-process sampler                   -- sampler/engine.py    (K1-K5, K-BC)
  3. reward callback ``reward_func(list[str], eos_token)``  -- reward/*
  4. no-grad log-prob (/value) pass, policy + ref          -- models/qwen2.py      (K9, K-LP)
  5. per-token rewards + advantage estimation              -- ops (K-GAE)
  6. optimisation: epochs x mini-batches x micro-batches   -- ops (K-LOSS), parallel/optimizer.py (K-AR)

and thin subclasses that override only the algorithm-specific hooks (``select_samples``,
``token_rewards``, ``advantages``, ``micro_loss``).  Public surface matches the reference
(SURVEY.md App. D): ``XTrainer(config, processing_class, policy, ref_policy, train_dataset,
reward_func=..., callbacks=..., [value_model])``, ``.train()``, ``.save_model()``,
``._save_checkpoint()``, ``.state / .control / .callback_handler``.
"""
from __future__ import annotations

import math
import os
import random
import time
from collections import defaultdict
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..config import RLConfig
from ..models.qwen2 import response_logprobs
from ..parallel.comm import Comm
from ..parallel.optimizer import FusedAdamW, build_param_groups
from ..runtime import checkpoint as ckpt
from ..runtime.offload import TieringEngine
from ..sampler import engine as sampler_engine
from ..utils import (INVALID_LOGPROB, disable_dropout_in_model, exact_div, masked_whiten, response_masks,
                     scatter_terminal_reward, truncate_response)
from ..utils.batching import create_batches
from ..utils.callbacks import (DEFAULT_CALLBACKS, CallbackHandler, OnlineTrainerState, ProgressCallback,
                               TrainerControl)
from ..utils.data import DataCollatorWithPadding, PromptLoader
from ..utils.faults import Heartbeat, maybe_inject
from ..utils.metrics import print_rich_table, reporting_callbacks
from ..utils.profiling import PhaseTimer, check_finite
from ..utils.schedules import get_scheduler


class PolicyAndValueWrapper(nn.Module):
    """Groups everything trainable (reference: grpo_trainer.py:82-88, ppo_trainer.py:87-99)."""

    def __init__(self, policy, value_model=None):
        super().__init__()
        self.policy = policy
        self.value_model = value_model

    def forward(self, *a, **kw):
        return self.policy(*a, **kw)


class RLTrainer:
    algo_name = "rl"
    samples_per_prompt_field: Optional[str] = None    # e.g. "grpo_sample_N"
    uses_value_model = False
    kl_in_reward = True                               # every algorithm except GRPO
    logs_policy_ratio_stats = True                    # RAFT drops approxkl/clipfrac/ratio

    # ------------------------------------------------------------------------------------------
    def __init__(self, config: RLConfig, processing_class, policy, ref_policy, train_dataset,
                 value_model=None, data_collator=None, eval_dataset=None, optimizers=(None, None),
                 callbacks=None, reward_func: Optional[Callable] = None, comm: Optional[Comm] = None,
                 device=None, accuracy_func: Optional[Callable] = None):
        if ref_policy is policy:
            raise ValueError("`policy` and `ref_policy` cannot be the same object; pass a copy")
        if reward_func is None:
            raise ValueError("reward_func is required")
        self.args = args = config
        self.processing_class = self.tokenizer = processing_class
        self.policy, self.ref_policy, self.value_model = policy, ref_policy, value_model
        self.reward_func, self.accuracy_func = reward_func, accuracy_func
        self.train_dataset, self.eval_dataset = train_dataset, eval_dataset
        self.train_dataset_len = len(train_dataset)
        self.data_collator = data_collator or DataCollatorWithPadding(processing_class)

        self.comm = comm or Comm.from_env(device)
        self.device = self.comm.device
        self._derive_batch_sizes()

        # run name with a timestamp agreed across ranks (reference broadcasts it: :241-243)
        stamp = self.comm.broadcast_object(int(time.time()), 0)
        args.run_name = args.run_name or f"{args.exp_name}__{args.seed}__{stamp}"
        self.local_seed = args.seed + self.comm.rank * 100003      # per-rank seed (:244)
        if args.stop_token == "eos":
            args.stop_token_id = processing_class.eos_token_id

        for m in (policy, ref_policy, value_model):
            if m is not None:
                disable_dropout_in_model(m)
        if args.logprob_top_p_consistent and args.top_p < 1.0:
            # experimental: policy and reference log-probs under the truncated softmax the sampler draws from
            for m in (policy, ref_policy):
                lm = getattr(m, "base_model", m)
                for mod in {id(x): x for x in (m, lm, getattr(lm, "model", None))}.values():
                    if mod is not None and hasattr(mod, "token_logprobs"):
                        mod.logprob_top_p = float(args.top_p)
        self.model = PolicyAndValueWrapper(policy, value_model if self.uses_value_model else None)
        self.model.to(self.device)
        ref_policy.eval()
        for p in ref_policy.parameters():
            p.requires_grad_(False)
        self.tiering = TieringEngine(self.device)
        self.tiering.register("ref", ref_policy, args.role_residency("ref"))
        # a model-based reward callback (reward/model_reward.py) joins the tiering engine as the "reward" role, so
        # ``offload_reward="host"`` parks the reward model in pinned host memory between reward phases (reference:
        # reward_model.to('cuda') / .to('cpu') around every call, GRPO/grpo.py:164,195)
        rm = getattr(reward_func, "rm", None)
        if isinstance(rm, nn.Module) and getattr(reward_func, "tiering", None) is None and next(rm.parameters()).device.type == self.device.type:
            self.tiering.register("reward", rm, args.role_residency("reward"))
            reward_func.tiering = self.tiering
        if args.gradient_checkpointing:
            policy.gradient_checkpointing_enable(args.gradient_checkpointing_kwargs)
            if self.model.value_model is not None:
                self.model.value_model.gradient_checkpointing_enable(args.gradient_checkpointing_kwargs)

        # Replicas must hold identical weights -- frozen ones included (base model under LoRA, reference policy, reward
        # model): the reference gets this from DDP's constructor broadcast.  Seeds alone are not enough:
        # device-side random init is not guaranteed to be bit-identical across processes.
        if self.comm.world_size > 1:
            for m in (policy, ref_policy, value_model if self.uses_value_model else None, getattr(reward_func, "rm", None)):
                if isinstance(m, nn.Module):
                    self.comm.broadcast_module_(m, 0)
        self.optimizer, self.lr_scheduler = optimizers
        if self.optimizer is None:
            self.optimizer = self.create_optimizer()
        if self.lr_scheduler is None:
            self.lr_scheduler = get_scheduler(args.lr_scheduler_type, self.optimizer, args.warmup_steps,
                                              args.num_total_batches, args.lr_scheduler_kwargs)
        if hasattr(self.optimizer, "max_grad_norm"):
            self.optimizer.max_grad_norm = args.max_grad_norm
        self.tiering.register_optimizer("optimizer", self.optimizer, args.role_residency("optimizer"))

        self.state = OnlineTrainerState(is_local_process_zero=self.comm.local_rank == 0,
                                        is_world_process_zero=self.comm.is_main)
        self.control = TrainerControl()
        cbs = [c() if isinstance(c, type) else c for c in DEFAULT_CALLBACKS]
        cbs += [ProgressCallback()] + reporting_callbacks(args, args.run_name) + list(callbacks or [])
        self.callback_handler = CallbackHandler(cbs, self.model, processing_class, self.optimizer, self.lr_scheduler)
        self.state.stateful_callbacks = {type(c).__name__: c.state() for c in cbs if hasattr(c, "state")}

        torch.manual_seed(args.seed)       # same shuffle on every rank, then shard
        self.dataloader = PromptLoader(train_dataset, args.local_batch_size, self.data_collator, seed=args.seed,
                                       rank=self.comm.rank, world_size=self.comm.world_size,
                                       drop_last=args.dataloader_drop_last)
        torch.manual_seed(self.local_seed)
        self._np_rng = np.random.RandomState(self.local_seed % (2 ** 31))
        self._select_gen = torch.Generator().manual_seed(self.local_seed)
        self.timer = PhaseTimer(self.device, nvtx=args.profile == "nvtx")
        self.heartbeat = Heartbeat(os.path.join(args.output_dir, "heartbeat"), self.comm.rank,
                                   args.watchdog_timeout_s, enable_watchdog=self.comm.world_size > 1)
        self.last_completions = None
        self.ckpt_writer = ckpt.AsyncCheckpointWriter(self.device)
        self._resumed = False
        self.io_bytes = {"h2d": 0, "d2h": 0}           # host<->device traffic of the public step API
        if self.comm.is_main:
            os.makedirs(args.output_dir, exist_ok=True)

    # ---- batch arithmetic (grpo_trainer.py:220-240) --------------------------------------------
    def _derive_batch_sizes(self):
        a = self.args
        a.world_size = self.comm.world_size
        a.local_batch_size = a.per_device_train_batch_size * a.gradient_accumulation_steps * a.num_mini_batches
        a.micro_batch_size = a.per_device_train_batch_size * a.world_size
        a.batch_size = a.local_batch_size * a.world_size
        a.mini_batch_size = exact_div(a.batch_size, a.num_mini_batches,
                                      "`batch_size` must be a multiple of `num_mini_batches`")
        a.local_mini_batch_size = exact_div(a.local_batch_size, a.num_mini_batches,
                                            "`local_batch_size` must be a multiple of `num_mini_batches`")
        a.num_total_batches = math.ceil(a.total_episodes / a.batch_size)

    @property
    def samples_per_prompt(self) -> int:
        return int(getattr(self.args, self.samples_per_prompt_field)) if self.samples_per_prompt_field else 1

    # ---- optimizer -----------------------------------------------------------------------------
    def create_optimizer(self):
        a = self.args
        groups = build_param_groups(self.model.named_parameters(), a.weight_decay, a.learning_rate)
        return self._make_optimizer(groups)

    def _make_optimizer(self, groups):
        a = self.args
        sd = torch.bfloat16 if a.optimizer_state_dtype == "bf16" else torch.float32
        mode = a.comm if self.device.type == "cuda" else "nccl"
        opt = FusedAdamW(groups, lr=a.learning_rate, betas=(a.adam_beta1, a.adam_beta2), eps=a.adam_epsilon,
                         weight_decay=a.weight_decay, state_dtype=sd, comm=self.comm, comm_mode=mode)
        if opt.comm_mode == "nccl" and a.ddp_bucket_mb > 0:
            opt.enable_bucket_overlap(int(a.ddp_bucket_mb) << 20)       # DDP-style: buckets reduce while the backward still runs
        return opt

    def get_train_dataloader(self):
        return self.dataloader

    # ---- hooks overridden per algorithm --------------------------------------------------------
    def rollout(self, queries: torch.Tensor) -> Dict[str, torch.Tensor]:
        a = self.args
        seed = sampler_engine.next_rollout_seed() if a.changing_seed else a.seed
        responses = sampler_engine.generate(self.samples_per_prompt, self.model, self.tokenizer, queries,
                                            a.temperature, a.response_length, top_p=a.top_p,
                                            seed=seed + self.comm.rank * 7919, backend=a.sampler,
                                            rollout_dtype=a.rollout_dtype, kv_block_size=a.kv_block_size,
                                            kv_cache_dtype=a.kv_cache_dtype, **self._weight_sync_kw())
        return {"responses": responses}

    def _weight_sync_kw(self) -> dict:
        """``weight_sync="sharded"`` (the default; needs fused DP on CUDA): the sampler arena refresh is K-BC across ranks."""
        a = self.args
        if (self.comm.world_size > 1 and self.device.type == "cuda" and a.comm == "fused" and a.weight_sync == "sharded"
                and a.sampler != "torch" and hasattr(self.policy, "peft_config")):
            return {"weight_sync_comm": self.comm}
        return {}

    def score(self, queries: torch.Tensor, responses: torch.Tensor) -> torch.Tensor:
        """Call the user reward callback exactly as the reference does (strings in, FloatTensor out)."""
        n = responses.shape[0] // queries.shape[0]
        if getattr(self.reward_func, "accepts_ids", False):
            return self.reward_func(queries.repeat_interleave(n, 0), responses, self.tokenizer).to(self.device).float()
        q_str = [q.replace(self.tokenizer.pad_token, "") for q in self.tokenizer.batch_decode(queries)]
        r_str = self.tokenizer.batch_decode(responses)
        self._last_strings = (q_str, r_str)
        texts = [q_str[i // n] + r for i, r in enumerate(r_str)]
        return self.reward_func(texts, self.tokenizer.eos_token).to(self.device).float()

    def select_samples(self, queries, rollout, scores):
        """Return dict(queries, responses, scores, log_scores[, seq_adv]) of the rows that are trained on."""
        return {"queries": queries, "responses": rollout["responses"], "scores": scores, "log_scores": scores}

    def token_rewards(self, sel, logprobs, ref_logprobs, seq_len, padding_mask, padding_mask_p1):
        a = self.args
        if self.kl_in_reward:
            kl = logprobs - ref_logprobs
            non_score = -a.kl_coef * kl
        else:
            non_score = torch.zeros_like(logprobs)
        rewards = scatter_terminal_reward(non_score, sel["scores"], seq_len)
        if a.whiten_rewards:
            rewards = masked_whiten(rewards, mask=~padding_mask_p1, shift_mean=True)
            rewards = torch.masked_fill(rewards, padding_mask_p1, 0)
        return rewards, non_score

    def after_rewards(self, R: Dict[str, torch.Tensor], roll: Dict[str, torch.Tensor]):
        return R

    def advantages(self, R):
        """Default: discounted suffix-sum of the per-token rewards (REINFORCE / ReMax / GRPO)."""
        a = self.args
        adv = ops.discounted_suffix_sum(R["rewards"], a.gamma)
        if a.advantage_whiten:
            adv = masked_whiten(adv, ~R["padding_mask"])
        return torch.masked_fill(adv, R["padding_mask"], 0), None

    @staticmethod
    def take_rows(R: Dict[str, torch.Tensor], idx: torch.Tensor) -> Dict[str, torch.Tensor]:
        n = R["responses"].shape[0]
        return {k: (v[idx] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n else v)
                for k, v in R.items()}

    def micro_loss(self, mb) -> (torch.Tensor, Dict[str, torch.Tensor]):
        a = self.args
        loss, st = ops.policy_loss_token(mb["new_logprobs"], mb["logprobs"], mb["advantages"], ~mb["padding_mask"],
                                         a.cliprange)
        st["pg_loss"] = loss.detach()
        return loss, st

    # ---- phase 4: no-grad logprob pass in token-budget chunks ----------------------------------
    @torch.no_grad()
    def logprob_pass(self, queries, responses):
        a = self.args
        ctx = queries.shape[1]
        pad = self.tokenizer.pad_token_id
        qr = torch.cat([queries, responses], 1)
        lens = (qr != pad).sum(1).tolist()
        # packed layout => the budget is on real tokens, not on padded rectangles
        chunks = create_batches(lens, a.token_budget_fwd, mode="packed")
        B, T_r = responses.shape
        lp = torch.empty(B, T_r, dtype=torch.float32, device=self.device)
        rlp = torch.empty_like(lp)
        vals = torch.zeros_like(lp) if self.uses_value_model else None
        self.policy.eval()
        for idx in chunks:
            ii = torch.as_tensor(idx, device=self.device)
            out = response_logprobs(self.policy, qr[ii], ctx, pad, a.temperature, want_entropy=False,
                                    value_model=self.model.value_model if self.uses_value_model else None)
            lp[ii] = out[0]
            if self.uses_value_model:
                vals[ii] = out[2]
            rlp[ii] = response_logprobs(self.ref_policy, qr[ii], ctx, pad, a.temperature, want_entropy=False)[0]
        return qr, lp, rlp, vals

    # ---- the update --------------------------------------------------------------------------
    def train(self):
        a = self.args
        self.policy.train()
        if not self._resumed:
            self.state.global_step = 0
            self.state.episode = 0
            self._try_resume()
        self.state.max_steps = a.num_total_batches * a.num_mini_batches          # (:444)
        self.state.num_train_epochs = a.total_episodes / self.train_dataset_len
        self.state.logging_steps, self.state.eval_steps, self.state.save_steps = a.logging_steps, a.eval_steps, a.save_steps
        self.control = self.callback_handler.on_train_begin(a, self.state, self.control)
        self.before_training()
        it = iter(self.dataloader)
        metrics = {}
        start_update = self.state.global_step + 1
        for update in range(start_update, a.num_total_batches + 1):
            if a.profile == "torch" and update == start_update + max(a.profile_update - 1, 0) and self.comm.is_main:
                metrics = self._profiled_update(update, next(it))
            else:
                metrics = self.train_one_update(update, next(it))
            self._write_memory_log(update, metrics)
            self.lr_scheduler.step()
            self.control = self.callback_handler.on_step_end(a, self.state, self.control)
            # state.max_steps = batches * minibatches is never reached by global_step (SURVEY.md 3.5 "Schedules"),
            # so the flow callback cannot stop the loop early; a stop raised by any *other* callback (early stopping,
            # a user callback) is honoured after this update's checkpoint.
            if self.control.should_save:
                with self.timer.phase("ckpt"):
                    self._save_checkpoint(self.model, trial=None, metrics=metrics)
                self.control = self.callback_handler.on_save(a, self.state, self.control)
            self.after_update(update, metrics)
            if self.control.should_training_stop:
                break
        self.control = self.callback_handler.on_train_end(a, self.state, self.control)
        if self.control.should_save:
            self._save_checkpoint(self.model, trial=None, metrics=metrics)
            self.control = self.callback_handler.on_save(a, self.state, self.control)
        self.ckpt_writer.wait()
        self.comm.barrier()
        self._maybe_load_best()
        self.heartbeat.close()
        return metrics

    def _profiled_update(self, update, data):
        """``profile="torch"``: one update under torch.profiler (CPU + CUDA activities); chrome trace and a kernel table under
        ``<scratch_dir>/profile``.  Numbers of this update are not benchmark numbers (profiler overhead)."""
        from torch.profiler import ProfilerActivity, profile
        acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if self.device.type == "cuda" else [])
        out_dir = os.path.join(self.args.scratch_dir, "profile")
        os.makedirs(out_dir, exist_ok=True)
        with profile(activities=acts) as prof:
            metrics = self.train_one_update(update, data)
        prof.export_chrome_trace(os.path.join(out_dir, f"update{update}_rank{self.comm.rank}.json"))
        with open(os.path.join(out_dir, f"update{update}_rank{self.comm.rank}.txt"), "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total" if self.device.type == "cuda" else "cpu_time_total", row_limit=60))
        return metrics

    def _write_memory_log(self, update, metrics):
        path = getattr(self.args, "memory_log", None)
        if not path or not self.comm.is_main:
            return
        import json
        row = {"update": update, **{k: v for k, v in metrics.items() if k.startswith(("mem/", "time/"))}}
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "a") as f:
            f.write(json.dumps(row) + "\n")

    def _maybe_load_best(self):
        """``load_best_model_at_end``: finish with the weights of ``state.best_model_checkpoint`` (HF semantics; the reference
        sets the flag only to satisfy EarlyStoppingCallback and its overridden ``train`` never acts on it -- SURVEY.md 2.2)."""
        a, best = self.args, self.state.best_model_checkpoint
        if not a.load_best_model_at_end or a.save_strategy == "no" or not best or not os.path.isdir(best):
            return
        last = os.path.join(a.output_dir, f"{ckpt.PREFIX}-{self.state.global_step}")
        if os.path.abspath(best) == os.path.abspath(last):
            return
        ckpt._load_model_into(self.policy, best)
        self._bump_policy_version()
        if self.comm.is_main:
            print(f"[train] loaded the best checkpoint ({best}, {a.metric_for_best_model} = {self.state.best_metric})")

    def before_training(self):
        pass

    def after_update(self, update: int, metrics: Dict[str, float]):
        pass

    def _phase(self, name: str, update: int):
        self.heartbeat.beat(name, update)
        maybe_inject(self.comm.rank, name, update)
        return self.timer.phase(name)

    def train_one_update(self, update: int, data) -> Dict[str, float]:
        a = self.args
        dev = self.device
        pad = self.tokenizer.pad_token_id
        t_start = time.time()
        self.state.episode += a.batch_size
        q_host = data["input_ids"]
        if dev.type == "cuda":
            q_host = q_host.pin_memory()                 # per-step inputs go pinned-host -> device
        queries = q_host.to(dev, non_blocking=True)
        self.io_bytes["h2d"] += q_host.numel() * q_host.element_size()

        with torch.no_grad():
            with self._phase("offload", update):
                self.tiering.evict("optimizer")                       # phase 1 (no-op when resident)
            with self._phase("rollout", update):
                rollout = self.rollout(queries)
            with self._phase("reward", update):
                scores_all = self.score(queries, rollout["responses"])
                rollout = self.post_score(queries, rollout, scores_all)
            rlhf_reward_mean = scores_all.mean()
            sel = self.select_samples(queries, rollout, scores_all)
            q, responses = sel["queries"], sel["responses"]
            with self._phase("logprob", update):
                self.tiering.fetch("ref")
                qr, logprobs, ref_logprobs, values = self.logprob_pass(q, responses)
                self.tiering.evict("ref")
            with self._phase("advantage", update):
                post = responses
                if a.stop_token_id is not None:
                    post = truncate_response(a.stop_token_id, pad, responses)
                seq_len, padding_mask, padding_mask_p1 = response_masks(post, pad)
                contain_eos = (post == self.tokenizer.eos_token_id).any(-1)
                if a.missing_eos_penalty is not None:
                    sel["scores"] = torch.where(contain_eos, sel["scores"], sel["scores"] - a.missing_eos_penalty)
                logprobs = torch.masked_fill(logprobs, padding_mask, INVALID_LOGPROB)
                ref_logprobs = torch.masked_fill(ref_logprobs, padding_mask, INVALID_LOGPROB)
                if values is not None:
                    values = torch.masked_fill(values, padding_mask_p1, 0)
                rewards, non_score = self.token_rewards(sel, logprobs, ref_logprobs, seq_len, padding_mask, padding_mask_p1)
                kl = logprobs - ref_logprobs
                roll = dict(mean_kl=kl.sum(1).mean(), mean_entropy=(-logprobs).sum(1).mean(),
                            non_score=non_score.sum(1).mean(),
                            # PPO/REINFORCE/ReMax log non-score + score (ppo_trainer.py:800); GRPO the raw mean
                            rlhf_reward=(non_score.sum(1).mean() + sel["log_scores"].mean()) if self.kl_in_reward
                            else rlhf_reward_mean,
                            scores=sel["log_scores"].mean(),
                            num_eos=(responses == self.tokenizer.eos_token_id).sum())
                R = dict(queries=q, responses=responses, query_responses=qr, logprobs=logprobs,
                         ref_logprobs=ref_logprobs, values=values, padding_mask=padding_mask,
                         padding_mask_p1=padding_mask_p1, seq_len=seq_len, rewards=rewards,
                         scores=sel["scores"], log_scores=sel["log_scores"])
                R = self.after_rewards(R, roll)          # RLOO / RAFT sub-select here
                adv, returns = self.advantages(R)
                check_finite("advantage", adv)
            self.tiering.fetch("optimizer")

        batch = {k: R[k] for k in ("query_responses", "responses", "logprobs", "ref_logprobs", "padding_mask",
                                   "padding_mask_p1", "values")}
        batch.update(advantages=adv, returns=returns, context_length=q.shape[1])
        with self._phase("train", update):
            stats = self.optimise(batch)
        self._bump_policy_version()

        with torch.no_grad():
            metrics = self.assemble_metrics(stats, roll)
            self.log_completions(R)
        secs = time.time() - t_start
        metrics["time/s_per_episode"] = secs / a.batch_size          # the reference's printed number (:726)
        metrics["throughput/episodes_per_s"] = a.batch_size / secs
        metrics.update(self.timer.collect())
        metrics["lr"] = self.lr_scheduler.get_last_lr()[0] if hasattr(self.lr_scheduler, "get_last_lr") else \
            self.optimizer.param_groups[0]["lr"]
        metrics["episode"] = self.state.episode
        self.state.epoch = self.state.episode / self.train_dataset_len
        self.state.global_step += 1
        self.log(metrics)
        return metrics

    def post_score(self, queries, rollout, scores):
        return rollout

    def _bump_policy_version(self):
        """Tell the resident sampler that the policy weights changed (it re-merges LoRA lazily)."""
        m = self.policy
        m = getattr(m, "base_model", m) if hasattr(m, "peft_config") else m
        m._nrl_version = getattr(m, "_nrl_version", 0) + 1

    # ---- phase 6 ------------------------------------------------------------------------------
    def _graph_micro_step(self):
        """CUDA-graph replay of the micro-step (trainer/graphed.py) when the step is graph-safe."""
        a = self.args
        mode = getattr(a, "train_cuda_graph", "auto")
        ok = (self.device.type == "cuda" and mode != "off" and a.lora_dropout == 0.0 and not a.logprob_top_p_consistent
              and ops.use_native(torch.empty(0, device=self.device)))
        if mode == "auto":
            ok = ok and not a.gradient_checkpointing           # torch.utils.checkpoint is not captured
        if not ok:
            return None
        if getattr(self, "_graphed", None) is None:
            from .graphed import GraphedMicroStep
            self._graphed = GraphedMicroStep(self)
        return self._graphed

    def optimise(self, batch) -> Dict[str, torch.Tensor]:
        a = self.args
        ctx = batch["context_length"]
        pad = self.tokenizer.pad_token_id
        n_local = batch["responses"].shape[0]
        # The number of optimizer steps per update is ``num_mini_batches`` whatever the number of trained rows (the LR
        # schedule and ``max_steps`` assume it): with ``train_samples_per_prompt = N`` the rows -- hence the local
        # mini-batch and its gradient-accumulation depth -- grow N-fold instead of the step count.
        lmb = max(a.per_device_train_batch_size, -(-n_local // a.num_mini_batches))
        accum = -(-lmb // a.per_device_train_batch_size)
        self._accum_steps = accum
        shape = (a.num_ppo_epochs, -(-n_local // lmb), accum)
        graphed = self._graph_micro_step()
        keys = graphed.stat_keys if graphed is not None else None
        rows = {}
        self.ckpt_writer.wait_snapshot()          # an in-flight checkpoint has finished READING the parameters / moments
        self.policy.train()
        for ep in range(a.num_ppo_epochs):
            b_inds = self._np_rng.permutation(n_local)
            for mi, mb_start in enumerate(range(0, n_local, lmb)):
                mini = b_inds[mb_start:mb_start + lmb]
                self.optimizer.zero_grad()
                for gi, mc_start in enumerate(range(0, len(mini), a.per_device_train_batch_size)):
                    inds = torch.as_tensor(mini[mc_start:mc_start + a.per_device_train_batch_size], device=self.device)
                    mb = {k: (v[inds] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n_local else v)
                          for k, v in batch.items()}
                    if graphed is not None:
                        rows[(ep, mi, gi)] = graphed(mb, ctx, pad)
                        keys = graphed.stat_keys
                        continue
                    if mc_start + a.per_device_train_batch_size >= len(mini):
                        self.optimizer.arm_overlap()        # last micro-step of the window (comm="nccl": bucketed all-reduce
                                                            # from backward hooks; earlier micro-steps do not reduce = no_sync)
                    out = response_logprobs(self.policy, mb["query_responses"], ctx, pad, a.temperature,
                                            want_entropy=True,
                                            value_model=self.model.value_model if self.uses_value_model else None)
                    new_lp = torch.masked_fill(out[0], mb["padding_mask"], INVALID_LOGPROB)
                    mb["new_logprobs"] = new_lp
                    if self.uses_value_model:
                        mb["vpred"] = out[2]
                    loss, st = self.micro_loss(mb)
                    (loss / accum).backward()
                    with torch.no_grad():
                        ent = out[1]
                        if a.stats_include_padding:
                            st["entropy"] = ent.mean()
                        else:
                            m = (~mb["padding_mask"]).float()
                            st["entropy"] = (ent * m).sum() / m.sum().clamp_min(1)
                        if keys is None:
                            keys = sorted(st)
                        rows[(ep, mi, gi)] = torch.stack([st[k].detach().float().reshape(()) for k in keys])
                self.optimizer.step()
        self.optimizer.zero_grad()
        # one [E, M, G, n_stats] tensor, sliced per statistic (slots that never ran stay zero)
        stats = defaultdict(lambda: torch.zeros(shape, device=self.device))
        if rows:
            table = torch.zeros(shape + (len(keys),), device=self.device)
            idx = torch.as_tensor(list(rows.keys()), device=self.device)
            table[idx[:, 0], idx[:, 1], idx[:, 2]] = torch.stack(list(rows.values()))
            for i, k in enumerate(keys):
                stats[k] = table[..., i]
        return stats

    # ---- metrics ------------------------------------------------------------------------------
    def assemble_metrics(self, stats, roll) -> Dict[str, float]:
        a = self.args
        inc = a.stats_include_padding
        dev_vals = {
            "objective/kl_old": self.kl_metric(stats, roll),
            "objective/entropy_old": roll["mean_entropy"],
            "objective/non_score_reward_old": roll["non_score"] if self.kl_in_reward else torch.zeros((), device=self.device),
            "eval_objective/rlhf_reward_old": roll["rlhf_reward"],
            "eval_objective/scores_old": roll["scores"],
            "loss/policy_avg_new": stats["pg_loss"].mean(),
            "policy/entropy_avg_new": stats["entropy"].mean(),
            "_num_eos": roll["num_eos"].float(),
        }
        if self.logs_policy_ratio_stats:
            dev_vals["policy/approxkl_avg_new"] = stats["approxkl_all" if inc else "approxkl_masked"].mean()
            dev_vals["policy/clipfrac_avg_new"] = stats["clipfrac"].mean()
            dev_vals["val/ratio_new"] = stats["ratio_mean_all" if inc else "ratio_mean_masked"].mean()
        if self.uses_value_model:
            dev_vals["loss/value_avg_new"] = stats["vf_loss"].mean()
            dev_vals["val/clipfrac_avg_new"] = stats["vf_clipfrac"].mean()
        keys = list(dev_vals)
        # ONE device -> host read for the whole update's metrics
        host = torch.stack([dev_vals[k].float().reshape(()) for k in keys]).cpu()
        self.io_bytes["d2h"] += host.numel() * host.element_size()
        local = {k: float(v) for k, v in zip(keys, host.tolist())}
        num_eos = int(local.pop("_num_eos"))
        if self.uses_value_model:
            local["eval_accuracy_new"] = 0.0
        out = self.comm.reduce_scalars(local, "mean")                     # ONE packed all-reduce (K23)
        if self.logs_policy_ratio_stats:
            r = stats["ratio_mean_all" if inc else "ratio_mean_masked"].reshape(-1)
            allr = self.comm.all_gather_cat(r)
            out["val/ratio_var_new"] = float(allr.var()) if allr.numel() > 1 else 0.0
        out["val/num_eos_tokens_old"] = num_eos
        return out

    def kl_metric(self, stats, roll):
        return roll["mean_kl"]

    def log(self, metrics: Dict[str, float]):
        logs = dict(metrics)
        logs["epoch"] = round(self.state.epoch, 6)
        logs["step"] = self.state.global_step
        self.state.log_history.append(logs)
        self.control = self.callback_handler.on_log(self.args, self.state, self.control, logs)

    def log_completions(self, sel):
        """5-row completions table (reference: grpo_trainer.py:712-724)."""
        a = self.args
        if not self.comm.is_main or a.num_sample_generations == 0:
            return
        k = min(5, sel["responses"].shape[0])
        q = [s.replace(self.tokenizer.pad_token, "") for s in self.tokenizer.batch_decode(sel["queries"][:k])]
        r = self.tokenizer.batch_decode(sel["responses"][:k])
        sc = sel["log_scores"][:k].tolist()
        rows = [[q[i], r[i], sc[i], sc[i]] for i in range(k)]
        self.last_completions = rows
        print_rich_table(["query", "response", "score", "rlhf_score"], rows)
        for cb in self.callback_handler.callbacks:
            if hasattr(cb, "log_table"):
                cb.log_table("completions", ["query", "response", "score", "rlhf_score"], rows)

    # ---- checkpointing (runtime/checkpoint.py holds the format) ---------------------------------
    def save_model(self, output_dir: Optional[str] = None, _internal_call: bool = False):
        ckpt.save_model(self, output_dir or self.args.output_dir)

    def _save_checkpoint(self, model, trial=None, metrics=None):
        ckpt.save_checkpoint(self, metrics)

    def _try_resume(self):
        a = self.args
        if a.resume == "never":
            return
        path = ckpt.find_resume_checkpoint(a) if a.resume == "auto" else a.resume
        if path:
            ckpt.load_checkpoint(self, path)
            self._resumed = True
            if self.comm.is_main:
                print(f"[resume] restored {path} at global_step={self.state.global_step}")
