"""Small tensor helpers whose semantics the RLHF algorithms depend on.

The reference borrows these from trl (import sites: /root/reference/GRPO/grpo_trainer.py:54-68);
trl is not a dependency here, so the behaviour is owned and unit-tested by this package
(tests/test_helpers.py).
"""
from __future__ import annotations

import torch

INVALID_LOGPROB = 1.0  # ref: GRPO/grpo_trainer.py:81 -- masked log-probs give ratio 1 and zero KL


def exact_div(a: int, b: int, custom_error_message: str = "") -> int:
    """Integer division that refuses to round (ref use: GRPO/grpo_trainer.py:226,229)."""
    q = a // b
    if a != q * b:
        raise ValueError(f"{custom_error_message}, inexact division: {a} / {b} = {a / b}")
    return q


def masked_mean(values: torch.Tensor, mask: torch.Tensor, axis=None) -> torch.Tensor:
    """Global (or per-axis) mean over the True entries of ``mask``."""
    mask = mask.to(values.dtype)
    if axis is not None:
        return (values * mask).sum(axis=axis) / mask.sum(axis=axis)
    return (values * mask).sum() / mask.sum()


def masked_var(values: torch.Tensor, mask: torch.Tensor, unbiased: bool = True) -> torch.Tensor:
    mean = masked_mean(values, mask)
    centered = values - mean
    var = masked_mean(centered * centered, mask)
    if unbiased:
        n = mask.sum()
        if n == 0:
            raise ValueError("masked_var: mask has no True entry")
        # Bessel correction n/(n-1); n == 1 yields inf/nan exactly like the trl helper would.
        var = var * (n / (n - 1))
    return var


def masked_whiten(values: torch.Tensor, mask: torch.Tensor, shift_mean: bool = True) -> torch.Tensor:
    """(x-mu)*rsqrt(var+1e-8) with masked mean / masked unbiased variance.

    ``shift_mean=False`` adds the mean back (variance-only normalisation).
    Used at GRPO/grpo_trainer.py:607,619 (reward whitening, advantage whitening).
    """
    mean, var = masked_mean(values, mask), masked_var(values, mask)
    whitened = (values - mean) * torch.rsqrt(var + 1e-8)
    if not shift_mean:
        whitened = whitened + mean
    return whitened


def first_true_indices(bools: torch.Tensor, dtype=torch.long) -> torch.Tensor:
    """Index of the first True along the last dim; ``row_len`` when the row has none."""
    row_len = bools.size(-1)
    idx = torch.arange(row_len, device=bools.device, dtype=dtype)
    filled = torch.where(bools, idx, torch.full_like(idx, row_len))
    return filled.min(dim=-1).values


def truncate_response(stop_token_id: int, pad_token_id: int, responses: torch.Tensor) -> torch.Tensor:
    """Everything *after* the first ``stop_token_id`` becomes ``pad_token_id`` (stop token kept)."""
    first_stop = first_true_indices(responses == stop_token_id).unsqueeze(-1)
    idx = torch.arange(responses.size(-1), device=responses.device).expand_as(responses)
    return torch.where(idx > first_stop, torch.full_like(responses, pad_token_id), responses)


def disable_dropout_in_model(model: torch.nn.Module) -> None:
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0


def response_masks(postprocessed_responses: torch.Tensor, pad_token_id: int):
    """sequence_lengths / padding_mask / padding_mask_p1 exactly as GRPO/grpo_trainer.py:566,589-594.

    ``sequence_lengths`` is the index of the last real token (the EOS when present).
    """
    seq_len = first_true_indices(postprocessed_responses == pad_token_id) - 1
    idx = torch.arange(postprocessed_responses.size(1), device=postprocessed_responses.device)
    idx = idx.unsqueeze(0).expand_as(postprocessed_responses)
    padding_mask = idx > seq_len.unsqueeze(1)
    padding_mask_p1 = idx > (seq_len + 1).unsqueeze(1)
    return seq_len, padding_mask, padding_mask_p1


def scatter_terminal_reward(rewards: torch.Tensor, scores: torch.Tensor, seq_len: torch.Tensor) -> torch.Tensor:
    """Add the scalar score at ``actual_end`` = seq_len+1 when it fits else seq_len (ref :600-603)."""
    T = rewards.size(1)
    p1 = seq_len + 1
    actual_end = torch.where(p1 < T, p1, seq_len)
    rows = torch.arange(rewards.size(0), device=rewards.device)
    rewards = rewards.clone()
    rewards[rows, actual_end] += scores.to(rewards.dtype)
    return rewards


def state_to_device(state, device, non_blocking: bool = False) -> None:
    """Move every tensor of an optimizer ``state`` dict (ref: GRPO/grpo_trainer.py:168-172).

    Kept for API parity; the tiering engine (runtime/offload.py) is the fast path.
    """
    for per_param in state.values():
        for k, v in per_param.items():
            if isinstance(v, torch.Tensor):
                per_param[k] = v.to(device, non_blocking=non_blocking)
