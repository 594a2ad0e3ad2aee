"""Token-budget bucket batching (dynamic mini-batching of the r1-v0 example).

Behaviour of the reference's ``_create_batches`` (/root/reference/examples/r1-v0/grpo_r1_trainer.py:
410-435): sort by length ascending and greedily grow a batch while
``max_len_in_batch * (count + 1) <= budget`` (the padded-token footprint).  Because this engine
packs tokens (no padding reaches a kernel) the same routine also offers ``mode="packed"`` where
the budget is on the *sum* of lengths -- the knapsack the survey (section 5.7) asks for.
"""
from __future__ import annotations

from typing import List, Sequence


def create_batches(lengths: Sequence[int], max_batch_memory_size: int, mode: str = "padded") -> List[List[int]]:
    order = sorted(range(len(lengths)), key=lambda i: (int(lengths[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    cur_max = 0
    cur_sum = 0
    for i in order:
        L = int(lengths[i])
        if mode == "padded":
            fits = max(cur_max, L) * (len(cur) + 1) <= max_batch_memory_size
        else:
            fits = cur_sum + L <= max_batch_memory_size
        if cur and not fits:
            batches.append(cur)
            cur, cur_max, cur_sum = [], 0, 0
        cur.append(i)
        cur_max = max(cur_max, L)
        cur_sum += L
    if cur:
        batches.append(cur)
    return batches


def strip_common_padding(queries, responses, pad_token_id: int):
    """Drop all-pad leading query columns and all-pad trailing response columns
    (sparse-GRPO de-padding, grpo_r1_trainer.py:572-584)."""
    q_real = (queries != pad_token_id).any(dim=0)
    first = int(q_real.float().argmax()) if bool(q_real.any()) else queries.shape[1] - 1
    r_real = (responses != pad_token_id).any(dim=0)
    if bool(r_real.any()):
        last = responses.shape[1] - int(r_real.flip(0).float().argmax())
    else:
        last = 1
    return queries[:, first:], responses[:, :last]
