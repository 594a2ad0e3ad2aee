"""The six algorithms (+ sparse GRPO) as thin subclasses of ``RLTrainer``.

Semantics table: SURVEY.md section 3.5.  Reference trainers:
GRPO/grpo_trainer.py, PPO/ppo_trainer.py, RLOO/rloo_trainer.py, ReMax/remax_trainer.py,
RAFT/raft_trainer.py, REINFORCE/reinforce_trainer.py, examples/r1-v0/grpo_r1_trainer.py.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict

import torch

from .. import ops
from ..models.qwen2 import response_logprobs
from ..parallel.optimizer import build_param_groups
from ..sampler import engine as sampler_engine
from ..utils import INVALID_LOGPROB, masked_whiten
from ..utils.batching import create_batches, strip_common_padding
from .base import RLTrainer


def _pick_one_of_n(B: int, n: int, gen: torch.Generator, device) -> torch.Tensor:
    """Flat row index of one uniformly random sample per prompt (reference: torch.randint(0,N,(B,)))."""
    r = torch.randint(0, n, (B,), generator=gen)
    return (torch.arange(B) * n + r).to(device)


# ================================================================================================
class ReinforceTrainer(RLTrainer):
    """REINFORCE with the PPO-clip surrogate; advantage whitening ON by default
    (REINFORCE/reinforce_trainer.py:568-591,634-640; REINFORCE/reinforce.py:103)."""
    algo_name = "reinforce"


# ================================================================================================
class GRPOTrainer(RLTrainer):
    """Group-normalised score, no KL in the reward, k3-KL inside the loss
    (GRPO/grpo_trainer.py:500-520,598-603,662-671)."""
    algo_name = "grpo"
    samples_per_prompt_field = "grpo_sample_N"
    kl_in_reward = False

    def group_normalise(self, scores: torch.Tensor) -> torch.Tensor:
        n = self.samples_per_prompt
        g = scores.view(-1, n)
        std = g.std(dim=1, keepdim=True)                 # unbiased, no eps (:508)
        return (g - g.mean(dim=1, keepdim=True)) / (std + self.args.grpo_std_eps)

    def select_samples(self, queries, rollout, scores):
        a, n, B = self.args, self.samples_per_prompt, queries.shape[0]
        z = self.group_normalise(scores)
        zflat = z.reshape(-1)
        zflat = torch.where(torch.isnan(zflat), torch.zeros_like(zflat), zflat)   # 0/0 -> 0 (:512)
        if a.train_samples_per_prompt >= n:
            idx = torch.arange(B * n, device=self.device)
            q = queries.repeat_interleave(n, 0)
        else:
            idx = _pick_one_of_n(B, n, self._select_gen, self.device)              # keep 1 of N (:504-514)
            q = queries
        return {"queries": q, "responses": rollout["responses"][idx], "scores": zflat[idx],
                "log_scores": scores[idx]}

    def micro_loss(self, mb):
        a = self.args
        loss, st = ops.policy_loss_token(mb["new_logprobs"], mb["logprobs"], mb["advantages"], ~mb["padding_mask"],
                                         a.cliprange, ref_logp=mb["ref_logprobs"], kl_coef=a.kl_coef)
        st["pg_loss"] = loss.detach()
        return loss, st

    def kl_metric(self, stats, roll):
        # the reference logs the mean of the *training-time* k = logp_new - logp_ref (:729)
        key = "refkl_all" if self.args.stats_include_padding else "refkl_masked"
        return stats[key].mean()


# ================================================================================================
class SparseGRPOTrainer(GRPOTrainer):
    """"Sparse GRPO" + dynamic mini-batching for 8000-token responses
    (examples/r1-v0/grpo_r1_trainer.py:562-591,691-791): drop rows whose normalised advantage is 0,
    strip common padding, pack the survivors into token-budget micro-buckets.

    Data-parallel safety (the reference would deadlock): gradients are
    accumulated locally and reduced once per optimizer step, and all ranks agree on the number of
    optimizer steps with one tiny all-reduce -- ranks that run out of rows contribute zero grads.
    """
    algo_name = "sparse_grpo"
    loss_agg = "sample_share"      # r1 per-bucket weighting (:787); "token" = global token mean

    def select_samples(self, queries, rollout, scores):
        sel = super().select_samples(queries, rollout, scores)
        keep = (sel["scores"] != 0).nonzero(as_tuple=False).squeeze(1)             # SPARSE (:565-568)
        self.last_kept = int(keep.numel())
        if keep.numel() == 0:
            keep = torch.zeros(1, dtype=torch.long, device=self.device)            # keep one row; its adv is 0
        sel = {k: v[keep] for k, v in sel.items()}
        q, r = strip_common_padding(sel["queries"], sel["responses"], self.tokenizer.pad_token_id)
        sel["queries"], sel["responses"] = q, r
        return sel

    def optimise(self, batch):
        a = self.args
        ctx, pad = batch["context_length"], self.tokenizer.pad_token_id
        n_local = batch["responses"].shape[0]
        mini_size = a.per_device_train_batch_size * a.gradient_accumulation_steps
        n_steps = self.comm.max_int((n_local + mini_size - 1) // mini_size)        # rank-invariant
        stats = defaultdict(list)
        lens = (batch["query_responses"] != pad).sum(1).tolist()
        self.policy.train()
        order = self._np_rng.permutation(n_local)
        for step in range(n_steps):
            mini = order[step * mini_size:(step + 1) * mini_size]
            self.optimizer.zero_grad()
            if len(mini) > 0:
                buckets = create_batches([lens[i] for i in mini], a.token_budget_train, mode="packed")
                n_tok_total = float(sum(int((~batch["padding_mask"][mini[j]]).sum()) for b in buckets for j in b))
                for b in buckets:
                    inds = torch.as_tensor([mini[j] for j in b], device=self.device)
                    mb = {k: (v[inds] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n_local else v)
                          for k, v in batch.items()}
                    out = response_logprobs(self.policy, mb["query_responses"], ctx, pad, a.temperature, want_entropy=True)
                    mb["new_logprobs"] = torch.masked_fill(out[0], mb["padding_mask"], INVALID_LOGPROB)
                    loss, st = self.micro_loss(mb)
                    if self.loss_agg == "token":
                        w = float((~mb["padding_mask"]).sum()) / max(n_tok_total, 1.0)
                    else:
                        w = len(b) / len(mini)
                    if b is buckets[-1]:
                        self.optimizer.arm_overlap()      # comm="nccl": bucket all-reduces ride the last backward; a rank without
                                                          # rows issues the same sequence inside step() (parallel/optimizer.py)
                    (loss * w).backward()
                    with torch.no_grad():
                        m = (~mb["padding_mask"]).float()
                        st["entropy"] = (out[1] * m).sum() / m.sum().clamp_min(1)
                        for k, v in st.items():
                            stats[k].append(v)
            self.optimizer.step()
        self.optimizer.zero_grad()
        zero = torch.zeros((), device=self.device)
        keys = ["pg_loss", "entropy", "approxkl_all", "approxkl_masked", "clipfrac", "ratio_mean_all",
                "ratio_mean_masked", "refkl_all", "refkl_masked"]
        return {k: (torch.stack(stats[k]) if stats[k] else zero.reshape(1)) for k in keys}


# ================================================================================================
class RLOOTrainer(RLTrainer):
    """Leave-one-out baseline over the summed per-token reward, sequence-level ratio
    (RLOO/rloo_trainer.py:571-613,660-669).  The log-prob pass runs on all B*N samples."""
    algo_name = "rloo"
    samples_per_prompt_field = "rloo_sample_N"

    def select_samples(self, queries, rollout, scores):
        n = self.samples_per_prompt
        return {"queries": queries.repeat_interleave(n, 0), "responses": rollout["responses"], "scores": scores,
                "log_scores": scores}

    def after_rewards(self, R, roll):
        a, n = self.args, self.samples_per_prompt
        rlhf = R["rewards"].sum(1)
        roll["rlhf_reward"] = rlhf.mean()                                   # log_rlhf_reward (:594,709)
        g = rlhf.view(-1, n)
        adv = g - (g.sum(1, keepdim=True) - g) / (n - 1)
        B = g.shape[0]
        if a.train_samples_per_prompt >= n:
            idx = torch.arange(B * n, device=self.device)
        else:
            idx = _pick_one_of_n(B, n, self._select_gen, self.device)
        out = self.take_rows(R, idx)
        out["seq_adv"] = adv.reshape(-1)[idx]
        return out

    def advantages(self, R):
        adv = R["seq_adv"]
        if self.args.advantage_whiten:
            adv = masked_whiten(adv, torch.ones_like(adv, dtype=torch.bool))
        return adv, None

    def micro_loss(self, mb):
        loss, st = ops.policy_loss_sequence(mb["new_logprobs"], mb["logprobs"], mb["advantages"], self.args.cliprange)
        st["pg_loss"] = loss.detach()
        return loss, st


# ================================================================================================
class RemaxTrainer(RLTrainer):
    """Score minus the score of a greedy decode of the same prompt
    (ReMax/remax_trainer.py:124-185,511-513,602-608)."""
    algo_name = "remax"

    def rollout(self, queries):
        a = self.args
        out = super().rollout(queries)
        # second, greedy (T=0) pass on the same resident engine -- no re-boot (remax_trainer.py:166-179)
        out["baseline_responses"] = sampler_engine.generate(
            1, self.model, self.tokenizer, queries, 0.0, a.response_length, top_p=1.0, seed=0,
            backend=a.sampler, rollout_dtype=a.rollout_dtype, kv_block_size=a.kv_block_size,
                                            kv_cache_dtype=a.kv_cache_dtype)
        return out

    def post_score(self, queries, rollout, scores):
        rollout["baseline_scores"] = self.score(queries, rollout["baseline_responses"])
        return rollout

    def select_samples(self, queries, rollout, scores):
        return {"queries": queries, "responses": rollout["responses"],
                "scores": scores - rollout["baseline_scores"], "log_scores": scores}


# ================================================================================================
class RAFTTrainer(RLTrainer):
    """Best-of-K -> SFT (NLL) (RAFT/raft_trainer.py:564-588,636).

    ``raft_select="best"`` (default) trains on the arg-max sample; ``"random"`` reproduces the shipped
    reference, whose arg-max index is overwritten by a random one (raft_trainer.py:586-588).
    """
    algo_name = "raft"
    samples_per_prompt_field = "raft_sample_K"
    logs_policy_ratio_stats = False

    def select_samples(self, queries, rollout, scores):
        n = self.samples_per_prompt
        return {"queries": queries.repeat_interleave(n, 0), "responses": rollout["responses"], "scores": scores,
                "log_scores": scores}

    def after_rewards(self, R, roll):
        n = self.samples_per_prompt
        rlhf = R["rewards"].sum(1)
        roll["rlhf_reward"] = rlhf.mean()
        g = rlhf.view(-1, n)
        B = g.shape[0]
        if getattr(self.args, "raft_select", "best") == "random":
            idx = _pick_one_of_n(B, n, self._select_gen, self.device)
        else:
            idx = torch.arange(B, device=self.device) * n + g.argmax(1)
        return self.take_rows(R, idx)

    def advantages(self, R):
        return torch.zeros_like(R["rewards"]), None

    def micro_loss(self, mb):
        loss = ops.nll_loss(mb["new_logprobs"])
        return loss, {"pg_loss": loss.detach()}


# ================================================================================================
class PPOTrainer(RLTrainer):
    """PPO-clip + GAE + clipped value loss, separate policy/value learning rates
    (PPO/ppo_trainer.py:341-402,630-634,668-697,732-756)."""
    algo_name = "ppo"
    uses_value_model = True

    def __init__(self, config, processing_class, policy, ref_policy, train_dataset, value_model=None, **kw):
        if value_model is None:
            raise ValueError("PPOTrainer needs a value_model")
        super().__init__(config, processing_class, policy, ref_policy, train_dataset, value_model=value_model, **kw)

    def create_optimizer(self):
        a = self.args
        plr = getattr(a, "policy_learning_rate", a.learning_rate)
        vlr = getattr(a, "value_learning_rate", a.learning_rate)
        groups = build_param_groups(self.model.policy.named_parameters(), a.weight_decay, plr)
        groups += build_param_groups(self.model.value_model.named_parameters(), a.weight_decay, vlr)
        return self._make_optimizer(groups)

    def advantages(self, R):
        a = self.args
        adv, returns = ops.gae(R["rewards"], R["values"], a.gamma, a.lam)
        if a.advantage_whiten:
            adv = masked_whiten(adv, ~R["padding_mask"])
        return torch.masked_fill(adv, R["padding_mask"], 0), returns

    def micro_loss(self, mb):
        a = self.args
        vpred = torch.masked_fill(mb["vpred"], mb["padding_mask_p1"], 0)
        vf_loss, vf_clipfrac = ops.value_loss(vpred, mb["values"], mb["returns"], ~mb["padding_mask_p1"],
                                              a.cliprange_value)
        pg_loss, st = ops.policy_loss_token(mb["new_logprobs"], mb["logprobs"], mb["advantages"],
                                            ~mb["padding_mask"], a.cliprange)
        st.update(pg_loss=pg_loss.detach(), vf_loss=vf_loss.detach(), vf_clipfrac=vf_clipfrac)
        return pg_loss + a.vf_coef * vf_loss, st
