"""Reference autoregressive sampler in plain PyTorch (CPU plumbing backend + oracle).

Plays the role vLLM plays in the reference's ``vllm_generate`` (/root/reference/GRPO/
grpo_trainer.py:122-166): temperature + nucleus (top-p 0.95) sampling, ``n`` samples per prompt,
stop at EOS (kept), right-pad with ``pad_token_id`` to ``max_tokens``, output ordered prompt-major /
sample-minor.  It runs *in process* on the live policy weights (LoRA un-merged: the adapter path is
just part of the module call), so there is no export, no engine boot and no weight copy.  The
native sm_100a sampler (sampler/native_sampler.py) must match this one token-for-token at
temperature 0.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from ..ops import reference as ref


def _unwrap(model):
    m = getattr(model, "policy", model)
    return getattr(m, "base_model", m) if hasattr(m, "peft_config") else m


@torch.no_grad()
def torch_generate(model, prompts: Sequence[Sequence[int]], n: int, temperature: float, top_p: float,
                   max_tokens: int, eos_token_id: Optional[int], pad_token_id: int, seed: int = 0,
                   batch_size: int = 256) -> torch.Tensor:
    """Returns LongTensor [len(prompts)*n, max_tokens] (right-padded with ``pad_token_id``)."""
    lm = _unwrap(model)
    device = next(lm.parameters()).device
    was_training = lm.training
    lm.eval()
    expanded: List[Sequence[int]] = [p for p in prompts for _ in range(n)]
    out = torch.full((len(expanded), max_tokens), pad_token_id, dtype=torch.long, device=device)
    gen = torch.Generator(device=device).manual_seed(int(seed))
    for s in range(0, len(expanded), batch_size):
        out[s:s + batch_size] = _generate_batch(lm, expanded[s:s + batch_size], temperature, top_p, max_tokens,
                                                eos_token_id, pad_token_id, gen, device)
    lm.train(was_training)
    return out


def _generate_batch(lm, prompts, temperature, top_p, max_tokens, eos_id, pad_id, gen, device):
    cfg = lm.config
    B = len(prompts)
    Lp = max(len(p) for p in prompts)
    ids = torch.full((B, Lp), pad_id, dtype=torch.long, device=device)
    key_ok = torch.zeros((B, Lp + max_tokens), dtype=torch.bool, device=device)
    for i, p in enumerate(prompts):
        ids[i, Lp - len(p):] = torch.as_tensor(list(p), dtype=torch.long, device=device)
        key_ok[i, Lp - len(p):Lp] = True
    plen = key_ok[:, :Lp].sum(1)
    pos = (key_ok[:, :Lp].long().cumsum(1) - 1).clamp_min(0)
    H, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    nl = cfg.num_hidden_layers
    dtype = next(lm.parameters()).dtype
    kc = [torch.zeros(B, Hkv, Lp + max_tokens, D, dtype=dtype, device=device) for _ in range(nl)]
    vc = [torch.zeros(B, Hkv, Lp + max_tokens, D, dtype=dtype, device=device) for _ in range(nl)]

    def step(tok, positions, start, length):
        """tok: [B, length]; writes KV at [start, start+length); returns hidden of the last position."""
        x = lm.model.embed_tokens(tok)
        cos, sin = ref.rope_cos_sin(positions.reshape(-1), D, cfg.rope_theta)
        end = start + length
        q_idx = torch.arange(start, end, device=device)
        k_idx = torch.arange(0, end, device=device)
        mask = (k_idx[None, :] <= q_idx[:, None])[None] & key_ok[:, None, :end]
        mask = mask[:, None]                                                   # [B,1,length,end]
        for li, layer in enumerate(lm.model.layers):
            h = layer.input_layernorm(x)
            at = layer.self_attn
            q = at.q_proj(h).view(B * length, H, D)
            k = at.k_proj(h).view(B * length, Hkv, D)
            v = at.v_proj(h).view(B, length, Hkv, D)
            q = ref.apply_rope(q, cos, sin).view(B, length, H, D).transpose(1, 2)
            k = ref.apply_rope(k, cos, sin).view(B, length, Hkv, D)
            kc[li][:, :, start:end] = k.transpose(1, 2)
            vc[li][:, :, start:end] = v.transpose(1, 2)
            kk = kc[li][:, :, :end].repeat_interleave(H // Hkv, dim=1)
            vv = vc[li][:, :, :end].repeat_interleave(H // Hkv, dim=1)
            att = torch.matmul(q.float(), kk.float().transpose(-1, -2)) / math.sqrt(D)
            att = att.masked_fill(~mask, float("-inf"))
            att = torch.nan_to_num(att.softmax(-1), nan=0.0)
            o = torch.matmul(att, vv.float()).to(dtype).transpose(1, 2).reshape(B, length, H * D)
            x = x + at.o_proj(o)
            x = x + layer.mlp(layer.post_attention_layernorm(x).reshape(B * length, -1)).view(B, length, -1)
        return lm.model.norm(x[:, -1])

    hidden = step(ids, pos, 0, Lp)
    out = torch.full((B, max_tokens), pad_id, dtype=torch.long, device=device)
    finished = torch.zeros(B, dtype=torch.bool, device=device)
    for t in range(max_tokens):
        logits = lm.lm_head(hidden).float()
        nxt = ref.top_p_sample(logits, temperature, top_p, gen)
        nxt = torch.where(finished, torch.full_like(nxt, pad_id), nxt)
        out[:, t] = nxt
        if eos_id is not None:
            finished = finished | (nxt == eos_id)
        if bool(finished.all()) or t == max_tokens - 1:
            break
        key_ok[:, Lp + t] = True
        hidden = step(nxt[:, None], (plen + t)[:, None], Lp + t, 1)
    return out
