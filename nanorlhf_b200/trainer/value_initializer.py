"""Value-model initialisation for PPO: pre-fit the critic to Monte-Carlo returns of the initial policy.

Reference: ``finetuned_value_model`` (/root/reference/PPO/value_initializer.py:69-388; call site
PPO/ppo.py:371-380; "about 15 minutes" on an A100, PPO/ppo.py:370).  Behaviour reproduced:

  1. one batch of ``min(train_data_size, len(dataset))`` prompts, n = 1 rollout with the policy;
  2. reward callback; policy / ref log-probs over the responses;
  3. per-token reward = -kl_coef * (logp - ref_logp) + score at ``actual_end``;
     return_t = r_t + gamma * return_{t+1}  (pure Monte-Carlo, no lambda, :353-359);
  4. 80/20 train/eval split; loss = 0.5 * masked_mean((V[:, ctx-1:-1] - return)^2, ~padding_mask_p1)
     (:199-212); AdamW lr 1e-3, ``reduce_lr_on_plateau(factor 0.5, patience 0)``, eval after every
     optimizer step, early stopping (patience 3), best weights restored at the end (PPO/ppo.py:84-110).

Differences by design: everything runs in-process on the resident sampler / fused log-prob kernels
(no vLLM boot, no /data/cache_value_model checkpoint churn -- the best state is kept in memory), and
``stop_token_id`` is honoured before truncation (the reference runs this step with it unset; parity
switch ``finetune_args.value_init_truncate=False``).
"""
from __future__ import annotations

import copy
import math
from typing import Callable, Dict, List

import torch

from .. import ops
from ..models.qwen2 import pack_padded, response_logprobs
from ..parallel.optimizer import FusedAdamW, build_param_groups
from ..sampler import engine as sampler_engine
from ..utils import INVALID_LOGPROB, disable_dropout_in_model, response_masks, scatter_terminal_reward, truncate_response
from ..utils.batching import create_batches
from ..utils.data import DataCollatorWithPadding
from ..utils.schedules import get_scheduler


def _values_for(value_model, query_responses, ctx, pad):
    """V aligned like the reference's ``score(h_last)[:, ctx-1:-1]`` -> [B, T_r]."""
    B, L = query_responses.shape
    ids, cu, pos, mx, flat = pack_padded(query_responses, pad)
    hidden = value_model.hidden_states(ids, cu, pos, mx)
    col, row = flat % L, flat // L
    sel = ((col >= ctx - 1) & (col <= L - 2)).nonzero(as_tuple=False).squeeze(1)
    vals = value_model.values(hidden[sel])
    out = torch.zeros((B, L - ctx), dtype=torch.float32, device=query_responses.device)
    return out.index_put((row[sel], col[sel] + 1 - ctx), vals)


def _value_mse(value_model, batch, pad):
    v = _values_for(value_model, batch["query_responses"], batch["context_length"], pad)
    m = (~batch["padding_mask_p1"]).float()
    return 0.5 * (((v - batch["returns"]) ** 2) * m).sum() / m.sum().clamp_min(1.0)


def finetuned_value_model(value_model, policy, ref_policy, reward_func: Callable, ppo_dataset, tokenizer, ppo_args,
                          finetune_args, device=None, verbose: bool = True):
    dev = torch.device(device) if device is not None else next(policy.parameters()).device
    pad = tokenizer.pad_token_id
    for m in (value_model, policy, ref_policy):
        disable_dropout_in_model(m)
        m.to(dev)
    n = min(finetune_args.train_data_size, len(ppo_dataset))
    collate = DataCollatorWithPadding(tokenizer)
    queries = collate([ppo_dataset[i] for i in range(n)])["input_ids"].to(dev)
    ctx = queries.shape[1]

    with torch.no_grad():
        responses = sampler_engine.generate(1, policy, tokenizer, queries, ppo_args.temperature, ppo_args.response_length,
                                            top_p=ppo_args.top_p, backend=ppo_args.sampler,
                                            rollout_dtype=ppo_args.rollout_dtype)
        q_str = [s.replace(tokenizer.pad_token, "") for s in tokenizer.batch_decode(queries)]
        r_str = tokenizer.batch_decode(responses)
        if getattr(reward_func, "accepts_ids", False):
            scores = reward_func(queries, responses, tokenizer).to(dev).float()
        else:
            scores = reward_func([a + b for a, b in zip(q_str, r_str)], tokenizer.eos_token).to(dev).float()
        qr = torch.cat([queries, responses], 1)
        lens = (qr != pad).sum(1).tolist()
        T_r = responses.shape[1]
        lp = torch.empty(n, T_r, device=dev)
        rlp = torch.empty_like(lp)
        policy.eval()
        for idx in create_batches(lens, finetune_args.token_budget_fwd, mode="packed"):
            ii = torch.as_tensor(idx, device=dev)
            lp[ii] = response_logprobs(policy, qr[ii], ctx, pad, ppo_args.temperature)[0]
            rlp[ii] = response_logprobs(ref_policy, qr[ii], ctx, pad, ppo_args.temperature)[0]
        post = responses
        stop_id = tokenizer.eos_token_id if (finetune_args.value_init_truncate and ppo_args.stop_token == "eos") else ppo_args.stop_token_id
        if stop_id is not None:
            post = truncate_response(stop_id, pad, responses)
        seq_len, padding_mask, padding_mask_p1 = response_masks(post, pad)
        lp = torch.masked_fill(lp, padding_mask, INVALID_LOGPROB)
        rlp = torch.masked_fill(rlp, padding_mask, INVALID_LOGPROB)
        rewards = scatter_terminal_reward(-ppo_args.kl_coef * (lp - rlp), scores, seq_len)
        returns = ops.discounted_suffix_sum(rewards, ppo_args.gamma)             # Monte-Carlo returns

    # ---- 80/20 split ------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(finetune_args.seed)
    perm = torch.randperm(n, generator=g).to(dev)
    n_train = max(1, int(n * finetune_args.train_split_rate))
    tr_idx, ev_idx = perm[:n_train], perm[n_train:]
    if ev_idx.numel() == 0:
        ev_idx = tr_idx

    def take(idx):
        return {"query_responses": qr[idx], "padding_mask_p1": padding_mask_p1[idx], "returns": returns[idx],
                "context_length": ctx}

    trainable = [(k, p) for k, p in value_model.named_parameters() if p.requires_grad]
    if not trainable:
        raise ValueError("value model has no trainable parameters")
    opt = FusedAdamW(build_param_groups(trainable, finetune_args.weight_decay, finetune_args.learning_rate),
                     lr=finetune_args.learning_rate, betas=(finetune_args.adam_beta1, finetune_args.adam_beta2),
                     eps=finetune_args.adam_epsilon, weight_decay=finetune_args.weight_decay)
    steps_per_epoch = max(1, math.ceil(n_train / (finetune_args.per_device_train_batch_size * finetune_args.gradient_accumulation_steps)))
    sched = get_scheduler(finetune_args.lr_scheduler_type, opt, 0, steps_per_epoch * finetune_args.num_train_epochs,
                          finetune_args.lr_scheduler_kwargs)

    @torch.no_grad()
    def evaluate():
        value_model.eval()
        tot, cnt = 0.0, 0
        for s in range(0, ev_idx.numel(), finetune_args.per_device_eval_batch_size):
            b = take(ev_idx[s:s + finetune_args.per_device_eval_batch_size])
            tot += float(_value_mse(value_model, b, pad)) * b["returns"].shape[0]
            cnt += b["returns"].shape[0]
        value_model.train()
        return tot / max(cnt, 1)

    best_loss, best_state, bad = float("inf"), None, 0
    history: List[Dict[str, float]] = []
    value_model.train()
    stop = False
    mb = finetune_args.per_device_train_batch_size
    for epoch in range(finetune_args.num_train_epochs):
        order = tr_idx[torch.randperm(n_train, generator=g).to(dev)]
        for s in range(0, n_train, mb * finetune_args.gradient_accumulation_steps):
            chunk = order[s:s + mb * finetune_args.gradient_accumulation_steps]
            n_micro = math.ceil(chunk.numel() / mb)
            opt.zero_grad()
            tr_loss = 0.0
            for k in range(n_micro):
                b = take(chunk[k * mb:(k + 1) * mb])
                loss = _value_mse(value_model, b, pad)
                (loss / n_micro).backward()
                tr_loss += float(loss) / n_micro
            opt.step()
            ev = evaluate()
            if isinstance(sched, torch.optim.lr_scheduler.ReduceLROnPlateau):
                sched.step(ev)
            else:
                sched.step()
            history.append({"epoch": epoch, "train_loss": tr_loss, "eval_loss": ev, "lr": opt.param_groups[0]["lr"]})
            if verbose:
                print(f"[value-init] epoch {epoch} train {tr_loss:.5f} eval {ev:.5f} lr {opt.param_groups[0]['lr']:.2e}")
            if ev < best_loss - 1e-12:
                best_loss, bad = ev, 0
                best_state = {k: p.detach().clone() for k, p in trainable}
            else:
                bad += 1
                if bad >= finetune_args.early_stopping_patience:
                    stop = True
                    break
        if stop:
            break
    if best_state is not None:                     # load_best_model_at_end
        with torch.no_grad():
            for k, p in trainable:
                p.copy_(best_state[k])
    opt.zero_grad()
    for _, p in trainable:                         # the PPO optimizer re-flattens these parameters
        p.grad = None
    value_model.value_init_history = history
    return value_model
