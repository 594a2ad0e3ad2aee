"""Datasets: synthetic hh-rlhf-shaped prompts, the reference's prompt templates, collation.

Reference behaviour reproduced (not its code): hh-rlhf prompts are the text between the first
``Human: `` and the following ``Assistant: `` of the ``chosen`` field, wrapped in a fixed Qwen chat
template and tokenized without padding into ``{"input_ids"}`` (/root/reference/GRPO/grpo.py:247-270);
the trainer shuffles, batches ``local_batch_size`` prompts with ``drop_last`` and left-pads
(/root/reference/GRPO/grpo_trainer.py:300-310).  There is no network on the GPU box, so
``synthetic_hh_prompts`` generates prompts with an hh-rlhf-like length distribution.
"""
from __future__ import annotations

import random
from typing import Callable, Dict, Iterator, List, Optional, Sequence

import torch

QWEN_CHAT_TEMPLATE = ("<|im_start|>system\nYou are Qwen, created by Alibaba Cloud. You are a helpful assistant."
                      "<|im_end|>\n<|im_start|>user\nQUESTION<|im_end|>\n<|im_start|>assistant\n")

R1_TEMPLATE = ("# Question:\nQUESTION\nPlease reason step by step, and put your final answer within \\boxed{}."
               "\n# Answer:\n")

_WORDS = ("how what why can you tell me about the best way to make find help with my a is it do does should "
          "people think when where who cook travel learn program fix repair write explain history science "
          "music movie recipe computer phone money health exercise language country city dog cat garden car "
          "book school work friend family weather food game sport idea story advice problem question").split()


def extract_hh_question(chosen: str) -> str:
    """Text between the first 'Human: ' and the next 'Assistant: ' (GRPO/grpo.py:256-258)."""
    s = chosen.find("Human: ") + len("Human: ")
    e = chosen.find("Assistant: ", s)
    return chosen[s:e]


def synthetic_hh_questions(n: int, seed: int = 0, min_words: int = 6, max_words: int = 60) -> List[str]:
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        k = int(min(max_words, max(min_words, rng.lognormvariate(2.9, 0.6))))
        out.append(" ".join(rng.choice(_WORDS) for _ in range(k)).capitalize() + "?\n\n")
    return out


class ListDataset(torch.utils.data.Dataset):
    def __init__(self, rows: List[Dict]):
        self.rows = rows

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        return self.rows[i]


def prepare_prompt_dataset(questions: Sequence[str], tokenizer, template: str = QWEN_CHAT_TEMPLATE,
                           max_prompt_tokens: Optional[int] = None) -> ListDataset:
    rows = []
    for q in questions:
        ids = tokenizer(template.replace("QUESTION", q), padding=False)["input_ids"]
        if max_prompt_tokens is not None:
            ids = ids[-max_prompt_tokens:]
        rows.append({"input_ids": ids})
    return ListDataset(rows)


def synthetic_hh_dataset(tokenizer, n: int, seed: int = 0, **kw) -> ListDataset:
    return prepare_prompt_dataset(synthetic_hh_questions(n, seed), tokenizer, **kw)


def synthetic_token_dataset(n: int, vocab_size: int, min_len: int = 24, max_len: int = 160, seed: int = 0,
                            reserved_ids: Sequence[int] = ()) -> ListDataset:
    """Prompts as raw token ids (used by bench.py with random-init full-vocab models)."""
    g = torch.Generator().manual_seed(seed)
    reserved = set(int(r) for r in reserved_ids)
    rows = []
    for _ in range(n):
        L = int(torch.randint(min_len, max_len + 1, (1,), generator=g))
        ids = torch.randint(0, vocab_size, (L,), generator=g).tolist()
        ids = [t if t not in reserved else (t + 1) % vocab_size for t in ids]
        rows.append({"input_ids": ids})
    return ListDataset(rows)


class DataCollatorWithPadding:
    def __init__(self, tokenizer):
        self.tokenizer = tokenizer

    def __call__(self, features):
        return self.tokenizer.pad([{"input_ids": f["input_ids"]} for f in features], return_tensors="pt")


class PromptLoader:
    """Shuffled, drop-last, rank-sharded, resumable batch iterator.

    All ranks draw the same permutation (seeded) and take interleaved shards -- the behaviour of an
    accelerate-prepared DataLoader.  ``drop_last=False`` keeps the tail of every epoch: the last
    global batch is completed with the first prompts of the same permutation (the update needs full batches).  ``state_dict`` makes the dataloader
    position part of checkpoints (the reference cannot resume).
    """

    def __init__(self, dataset, batch_size: int, collate_fn: Callable, seed: int, rank: int = 0,
                 world_size: int = 1, shuffle: bool = True, drop_last: bool = True):
        self.dataset, self.batch_size, self.collate_fn = dataset, batch_size, collate_fn
        self.seed, self.rank, self.world_size = seed, rank, world_size
        self.shuffle, self.drop_last = shuffle, drop_last
        self.epoch, self.cursor = 0, 0
        if len(dataset) < batch_size * world_size:
            raise ValueError(f"dataset of {len(dataset)} prompts is smaller than one global batch "
                             f"({batch_size} x {world_size})")

    def __len__(self):
        per_step = self.batch_size * self.world_size
        return len(self.dataset) // per_step if self.drop_last else -(-len(self.dataset) // per_step)

    def _order(self) -> List[int]:
        idx = list(range(len(self.dataset)))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            idx = torch.randperm(len(idx), generator=g).tolist()
        return idx

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        while True:   # infinite "repeat_generator" (grpo_trainer.py:416-420)
            order = self._order()
            per_step = self.batch_size * self.world_size
            n_steps = len(order) // per_step
            if not self.drop_last and len(order) % per_step:
                n_steps += 1
                order = order + order[:per_step - len(order) % per_step]
            while self.cursor < n_steps:
                s = self.cursor * per_step
                chunk = order[s:s + per_step][self.rank::self.world_size]
                self.cursor += 1
                yield self.collate_fn([self.dataset[i] for i in chunk])
            self.epoch += 1
            self.cursor = 0

    def state_dict(self):
        return {"epoch": self.epoch, "cursor": self.cursor, "seed": self.seed}

    def load_state_dict(self, sd):
        self.epoch, self.cursor = sd["epoch"], sd["cursor"]
