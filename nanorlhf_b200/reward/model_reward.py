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

    # ---- scoring of a padded id matrix in length-sorted batches ---------------------------------------
    @torch.no_grad()
    def _score_padded(self, full: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
        """``full``: [n, Lmax] RM ids on the RM's device, right-padded; ``lens``: [n] row lengths.  Rows are scored in
        ascending-length batches of ``reward_batch_size`` (bounded by ``token_budget`` padded tokens); the scores stay on
        the device -- the only host read is the [n] length vector that fixes the batch boundaries."""
        cfg = self.rm.config
        n = full.shape[0]
        out = torch.zeros(n, dtype=torch.float32, device=full.device)
        order = torch.argsort(lens, stable=True)
        sorted_lens = lens[order].tolist()                      # one small D2H; batch shapes are host decisions
        if self.tiering is not None:
            self.tiering.fetch("reward")
        i = 0
        while i < n:
            j, L = i, 0
            while j < n and j - i < self.bs:
                L2 = max(L, sorted_lens[j])
                if j > i and L2 * (j - i + 1) > self.token_budget:
                    break
                L, j = L2, j + 1
            idx = order[i:j]
            ids = full[idx, :max(L, 1)]
            logits = self.rm(ids, ids != cfg.pad_token_id)
            out[idx] = logits.reshape(len(idx), -1)[:, 0].float()
            i = j
        if self.tiering is not None:
            self.tiering.evict("reward")
        return out

    def _score_ids(self, rows: List[List[int]]) -> torch.Tensor:
        """Python id lists (the tokenizer path) -> one padded matrix, one H2D copy."""
        cfg = self.rm.config
        lens = torch.tensor([len(r) for r in rows], dtype=torch.long)
        full = torch.full((len(rows), max(int(lens.max()), 1) if len(rows) else 1), cfg.pad_token_id, dtype=torch.long)
        for i, r in enumerate(rows):
            full[i, :len(r)] = torch.as_tensor(r, dtype=torch.long)
        return self._score_padded(full.to(self.device, non_blocking=True), lens.to(self.device))

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

    @torch.no_grad()
    def score_ids(self, queries: torch.Tensor, responses: torch.Tensor, tokenizer) -> torch.Tensor:
        """Id-level fast path (no RM tokenizer: synthetic / offline runs).  Policy ids are folded into the RM vocabulary and
        every row ``[CLS] question [SEP] response [SEP]`` is assembled ON THE DEVICE with index arithmetic -- no Python loop
        over the 2048 x 1500 response ids, no per-row tensors, no per-batch host read (the reference decodes to strings and
        re-tokenizes on the CPU: /root/reference/GRPO/grpo.py:166-188).  ``queries`` are left-padded, ``responses``
        right-padded; a response ends before its first pad / EOS."""
        cfg = self.rm.config
        dev = self.device
        pad, eos = tokenizer.pad_token_id, tokenizer.eos_token_id
        V = cfg.vocab_size
        q = queries.to(dev)
        r = responses.to(dev)
        n, ctx = q.shape
        T_r = r.shape[1]
        lq = (q != pad).sum(1)                                                  # real tokens sit at the right end
        stop = (r == pad) | (r == eos) if eos is not None else (r == pad)
        lr = torch.where(stop.any(1), stop.int().argmax(1), torch.full((n,), T_r, device=dev))
        if self.max_length is not None:
            lq = lq.clamp(max=max(self.max_length - 3, 0))
            lr = torch.minimum(lr, (self.max_length - 3 - lq).clamp(min=0))
        lens = lq + lr + 3
        Lmax = int(ctx + T_r + 3)
        c = torch.arange(Lmax, device=dev)[None, :]
        lq_, lr_ = lq[:, None], lr[:, None]
        qm = 3 + q % (V - 3)
        rm_ = 3 + r % (V - 3)
        q_src = (ctx - lq_ + (c - 1)).clamp(0, ctx - 1)                         # column of the (c-1)-th real query token
        r_src = (c - lq_ - 2).clamp(0, T_r - 1)
        full = torch.full((n, Lmax), cfg.pad_token_id, dtype=torch.long, device=dev)
        full = torch.where((c >= 1) & (c <= lq_), qm.gather(1, q_src.expand(n, Lmax)), full)
        full = torch.where((c >= lq_ + 2) & (c < lq_ + 2 + lr_), rm_.gather(1, r_src.expand(n, Lmax)), full)
        full = torch.where((c == lq_ + 1) | (c == lq_ + 2 + lr_), torch.full_like(full, cfg.sep_token_id), full)
        full[:, 0] = cfg.cls_token_id
        return self._score_padded(full[:, :int(lens.max())], lens)
