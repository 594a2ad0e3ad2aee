"""Reward callback contract + simple rule rewards.

Contract (reference: /root/reference/GRPO/grpo.py:162): ``reward_func(pmt_and_responses: list[str],
eos_token: str) -> FloatTensor[len]`` -- the trainer hands over decoded prompt+response strings.
A callback may set ``accepts_ids = True`` to receive ``(queries, responses, tokenizer)`` id tensors
instead and skip the decode -> parse round trip (tokenizers row).
"""
from __future__ import annotations

from typing import List

import torch


def split_prompt_response(text: str, eos_token: str):
    """Recover (question, response) from the chat-templated string the way the reference's
    reward_func does (GRPO/grpo.py:174-186): question between ``user\\n`` and ``<|im_end|>``,
    response after ``<|im_start|>assistant\\n`` up to the first eos."""
    qs = text.find("user\n") + len("user\n")
    qe = text.find("<|im_end|>", qs)
    question = text[qs:qe]
    rs = text.find("<|im_start|>assistant\n") + len("<|im_start|>assistant\n")
    re_ = text.find(eos_token, rs)
    response = text[rs:] if re_ == -1 else text[rs:re_]
    return question, response


class ConstantReward:
    """Plumbing reward: every sample scores ``value`` (BASELINE.json config 1)."""

    def __init__(self, value: float = 1.0):
        self.value = value

    def __call__(self, pmt_and_responses: List[str], eos_token: str) -> torch.Tensor:
        return torch.full((len(pmt_and_responses),), float(self.value))


class LengthReward:
    """Scores -|len(response) - target| / target: a cheap, learnable rule reward for tests."""

    def __init__(self, target_chars: int = 40):
        self.target = target_chars

    def __call__(self, pmt_and_responses: List[str], eos_token: str) -> torch.Tensor:
        out = []
        for t in pmt_and_responses:
            _, r = split_prompt_response(t, eos_token)
            out.append(-abs(len(r) - self.target) / self.target)
        return torch.tensor(out, dtype=torch.float32)


class TokenIdReward:
    """Id-level rule reward (fraction of response tokens equal to ``token``); ``accepts_ids`` path."""
    accepts_ids = True

    def __init__(self, token: int):
        self.token = token

    def __call__(self, queries, responses, tokenizer) -> torch.Tensor:
        real = responses != tokenizer.pad_token_id
        return ((responses == self.token) & real).float().sum(1) / real.float().sum(1).clamp_min(1)
