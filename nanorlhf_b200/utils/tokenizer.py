"""Offline tokenizers.

The GPU box has no network, so ``AutoTokenizer.from_pretrained("Qwen/...")`` (reference:
GRPO/grpo.py:209-216) cannot be assumed.  ``ByteTokenizer`` is a self-contained UTF-8 byte-level
tokenizer with the Qwen chat special tokens as single ids; ``HFTokenizerAdapter`` wraps an installed
``tokenizer.json`` when one exists on disk.  Both expose the small surface the trainers use:
``pad_token_id / eos_token_id / eos_token / pad_token``, ``__call__``, ``batch_decode``, ``pad`` and
left padding (the reference pads queries on the left, GRPO/grpo.py:211).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Union

import torch

SPECIALS = ["<|endoftext|>", "<|im_start|>", "<|im_end|>", "[PAD]"]


class ByteTokenizer:
    padding_side = "left"

    def __init__(self, vocab_size: Optional[int] = None):
        self.special_tokens: Dict[str, int] = {s: 256 + i for i, s in enumerate(SPECIALS)}
        self.id_to_special = {v: k for k, v in self.special_tokens.items()}
        self.eos_token = "<|im_end|>"
        self.pad_token = "[PAD]"
        self.bos_token = None
        self.eos_token_id = self.special_tokens[self.eos_token]
        self.pad_token_id = self.special_tokens[self.pad_token]
        self.vocab_size = max(vocab_size or 0, 256 + len(SPECIALS))
        self.chat_template = None

    def __len__(self):
        return self.vocab_size

    def add_special_tokens(self, mapping: Dict[str, str]) -> int:
        added = 0
        for _, tok in mapping.items():
            if tok not in self.special_tokens:
                self.special_tokens[tok] = self.vocab_size
                self.id_to_special[self.vocab_size] = tok
                self.vocab_size += 1
                added += 1
        if "pad_token" in mapping:
            self.pad_token = mapping["pad_token"]
            self.pad_token_id = self.special_tokens[self.pad_token]
        if "eos_token" in mapping:
            self.eos_token = mapping["eos_token"]
            self.eos_token_id = self.special_tokens[self.eos_token]
        return added

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        i = 0
        specials = sorted(self.special_tokens, key=len, reverse=True)
        while i < len(text):
            for s in specials:
                if text.startswith(s, i):
                    ids.append(self.special_tokens[s])
                    i += len(s)
                    break
            else:
                j = i + 1
                while j < len(text) and not any(text.startswith(s, j) for s in specials):
                    j += 1
                ids.extend(text[i:j].encode("utf-8"))
                i = j
        return ids

    def decode(self, ids: Sequence[int], skip_special_tokens: bool = False) -> str:
        out, buf = [], bytearray()
        for t in (int(x) for x in ids):
            if t < 256:
                buf.append(t)
                continue
            if buf:
                out.append(buf.decode("utf-8", errors="replace"))
                buf = bytearray()
            if not skip_special_tokens:
                out.append(self.id_to_special.get(t, f"<|unk{t}|>"))
        if buf:
            out.append(buf.decode("utf-8", errors="replace"))
        return "".join(out)

    def batch_decode(self, batch, skip_special_tokens: bool = False) -> List[str]:
        if isinstance(batch, torch.Tensor):
            batch = batch.tolist()
        return [self.decode(x, skip_special_tokens) for x in batch]

    def __call__(self, text: Union[str, List[str]], padding: bool = False, return_tensors: Optional[str] = None, **_):
        single = isinstance(text, str)
        enc = [self.encode(t) for t in ([text] if single else text)]
        if padding or return_tensors == "pt":
            padded = self.pad([{"input_ids": e} for e in enc], return_tensors=return_tensors)
            return padded
        return {"input_ids": enc[0] if single else enc}

    def pad(self, features: List[Dict[str, List[int]]], return_tensors: Optional[str] = "pt", **_):
        L = max(len(f["input_ids"]) for f in features)
        ids, mask = [], []
        for f in features:
            x = list(f["input_ids"])
            n = L - len(x)
            if self.padding_side == "left":
                ids.append([self.pad_token_id] * n + x)
                mask.append([0] * n + [1] * len(x))
            else:
                ids.append(x + [self.pad_token_id] * n)
                mask.append([1] * len(x) + [0] * n)
        if return_tensors == "pt":
            return {"input_ids": torch.tensor(ids, dtype=torch.long), "attention_mask": torch.tensor(mask, dtype=torch.long)}
        return {"input_ids": ids, "attention_mask": mask}

    def save_pretrained(self, path: str):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
            json.dump({"tokenizer_class": "ByteTokenizer", "padding_side": self.padding_side,
                       "eos_token": self.eos_token, "pad_token": self.pad_token,
                       "vocab_size": self.vocab_size}, f, indent=2)
        with open(os.path.join(path, "special_tokens_map.json"), "w") as f:
            json.dump({"eos_token": self.eos_token, "pad_token": self.pad_token}, f, indent=2)
        with open(os.path.join(path, "added_tokens.json"), "w") as f:
            json.dump(self.special_tokens, f, indent=2)

    @classmethod
    def from_pretrained(cls, path: str, **_):
        tok = cls()
        p = os.path.join(path, "added_tokens.json")
        if os.path.exists(p):
            with open(p) as f:
                tok.special_tokens = {k: int(v) for k, v in json.load(f).items()}
            tok.id_to_special = {v: k for k, v in tok.special_tokens.items()}
        p = os.path.join(path, "tokenizer_config.json")
        if os.path.exists(p):
            with open(p) as f:
                c = json.load(f)
            tok.vocab_size = c.get("vocab_size", tok.vocab_size)
            tok.eos_token, tok.pad_token = c.get("eos_token", tok.eos_token), c.get("pad_token", tok.pad_token)
            tok.eos_token_id = tok.special_tokens[tok.eos_token]
            tok.pad_token_id = tok.special_tokens[tok.pad_token]
        return tok


class HFTokenizerAdapter:
    """Use a real HF tokenizer from a local directory when one is available (format compatibility)."""

    def __new__(cls, path: str, **kw):
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(path, padding_side=kw.pop("padding_side", "left"), **kw)
        if tok.pad_token is None:
            tok.add_special_tokens({"pad_token": "[PAD]"})
        return tok


def load_tokenizer(path_or_name: str, vocab_size: Optional[int] = None):
    """Local HF tokenizer if the directory has one, else the byte-level fallback."""
    if path_or_name and os.path.isdir(path_or_name):
        if os.path.exists(os.path.join(path_or_name, "tokenizer.json")):
            return HFTokenizerAdapter(path_or_name)
        if os.path.exists(os.path.join(path_or_name, "tokenizer_config.json")):
            return ByteTokenizer.from_pretrained(path_or_name)
    return ByteTokenizer(vocab_size)
