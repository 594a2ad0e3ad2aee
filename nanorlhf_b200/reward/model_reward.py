"""Model-based reward callback (DeBERTa-v3 reward model).

Reference behaviour (/root/reference/GRPO/grpo.py:159-198): parse (question, response) back out of the
decoded chat-templated string, tokenize the *pair* with the reward model's own tokenizer, score in
batches of ``reward_batch_size`` taking ``logits.squeeze()``, moving the RM GPU<->CPU around the pass.
Here the RM is resident in HBM in bf16 (tiering via ``TieringEngine`` when ``offload_reward=host``),
batches are length-bucketed, and there is an id-level fast path for synthetic / offline runs where no
RM tokenizer exists (``accepts_ids``: policy ids are folded into the RM vocab as
``[CLS] question [SEP] response [SEP]``).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .api import split_prompt_response


class ModelReward:
    def __init__(self, reward_model, rm_tokenizer=None, reward_batch_size: int = 16, device=None,
                 tiering=None, max_length: Optional[int] = None, token_budget: int = 32768):
        self.rm = reward_model.eval()
        self.tok = rm_tokenizer
        self.bs = reward_batch_size
        self.device = device or next(reward_model.parameters()).device
        self.tiering = tiering
        self.max_length = max_length
        self.token_budget = token_budget
        self.accepts_ids = rm_tokenizer is None
        for p in self.rm.parameters():
            p.requires_grad_(False)

    # ---- scoring of pre-built id batches (length-bucketed) ---------------------------------------
    @torch.no_grad()
    def _score_ids(self, rows: List[List[int]]) -> torch.Tensor:
        cfg = self.rm.config
        out = torch.zeros(len(rows), dtype=torch.float32)
        order = sorted(range(len(rows)), key=lambda i: len(rows[i]))
        if self.tiering is not None:
            self.tiering.fetch("reward")
        i = 0
        while i < len(order):
            # grow the batch up to reward_batch_size rows while the padded rectangle fits the token budget
            j, L = i, 0
            while j < len(order) and j - i < self.bs:
                L2 = max(L, len(rows[order[j]]))
                if j > i and L2 * (j - i + 1) > self.token_budget:
                    break
                L, j = L2, j + 1
            idx = order[i:j]
            ids = torch.full((len(idx), L), cfg.pad_token_id, dtype=torch.long)
            for r, k in enumerate(idx):
                ids[r, :len(rows[k])] = torch.tensor(rows[k], dtype=torch.long)
            ids = ids.to(self.device, non_blocking=True)
            logits = self.rm(ids, ids != cfg.pad_token_id)
            out[idx] = logits.reshape(len(idx), -1)[:, 0].float().cpu()
            i = j
        if self.tiering is not None:
            self.tiering.evict("reward")
        return out

    def _pair(self, q_ids: List[int], r_ids: List[int]) -> List[int]:
        cfg = self.rm.config
        row = [cfg.cls_token_id] + q_ids + [cfg.sep_token_id] + r_ids + [cfg.sep_token_id]
        if self.max_length is not None and len(row) > self.max_length:
            row = row[:self.max_length - 1] + [cfg.sep_token_id]
        return row

    # ---- the two call shapes ------------------------------------------------------------------------
    def __call__(self, *args):
        if len(args) == 3 and isinstance(args[0], torch.Tensor):
            return self.score_ids(*args)
        return self.score_strings(*args)

    def score_strings(self, pmt_and_responses: List[str], eos_token: str) -> torch.Tensor:
        if self.tok is None:
            raise RuntimeError("string scoring needs the reward model's tokenizer")
        rows = []
        for text in pmt_and_responses:
            q, r = split_prompt_response(text, eos_token)
            enc = self.tok(q, r)
            rows.append(list(enc["input_ids"]))
        return self._score_ids(rows)

    def score_ids(self, queries: torch.Tensor, responses: torch.Tensor, tokenizer) -> torch.Tensor:
        cfg = self.rm.config
        pad, eos = tokenizer.pad_token_id, tokenizer.eos_token_id
        V = cfg.vocab_size
        q_rows, r_rows = queries.tolist(), responses.tolist()
        rows = []
        for q, r in zip(q_rows, r_rows):
            q = [3 + (t % (V - 3)) for t in q if t != pad]
            rr = []
            for t in r:
                if t == pad or t == eos:
                    break
                rr.append(3 + (t % (V - 3)))
            rows.append(self._pair(q, rr))
        return self._score_ids(rows)
