"""Qwen2 / Qwen2.5 decoder, packed-varlen, built on the op layer.

The reference reaches this model through ``transformers.AutoModelForCausalLM`` +
flash-attn-2 (/root/reference/GRPO/grpo.py:218-224) and, for the PPO critic,
``AutoModelForSequenceClassification(num_labels=1)`` (/root/reference/PPO/ppo.py:280-287).
This is a from-scratch module with the same parameter names as the HF checkpoint format
(``model.layers.N.self_attn.q_proj.weight`` ...) so safetensors checkpoints round-trip, but with a
B200-first execution model:

* tokens are packed ``[total_tokens]`` with ``cu_seqlens`` -- no padding ever reaches a kernel;
* there is no ``[tokens, vocab]`` logits tensor on the hot path: ``token_logprobs`` calls the
  fused lm-head log-prob op (K-LP);
* every elementwise/normalisation op is a single fused kernel from ``ops``.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import asdict, dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.utils.checkpoint

from .. import ops


@dataclass
class Qwen2Config:
    vocab_size: int = 151936
    hidden_size: int = 1536
    intermediate_size: int = 8960
    num_hidden_layers: int = 28
    num_attention_heads: int = 12
    num_key_value_heads: int = 2
    head_dim: Optional[int] = None
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    tie_word_embeddings: bool = True
    max_position_embeddings: int = 32768
    attention_bias: bool = True          # Qwen2: bias on q/k/v only
    pad_token_id: Optional[int] = None
    eos_token_id: Optional[int] = None
    bos_token_id: Optional[int] = None
    model_type: str = "qwen2"
    name_or_path: str = ""
    num_labels: int = 1

    def __post_init__(self):
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads

    # ---- named shapes ---------------------------------------------------
    @classmethod
    def qwen2_5_1_5b(cls, **kw):
        return cls(vocab_size=151936, hidden_size=1536, intermediate_size=8960, num_hidden_layers=28,
                   num_attention_heads=12, num_key_value_heads=2, tie_word_embeddings=True,
                   name_or_path="Qwen/Qwen2.5-1.5B-Instruct", **kw)

    @classmethod
    def qwen2_5_7b(cls, **kw):
        return cls(vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_hidden_layers=28,
                   num_attention_heads=28, num_key_value_heads=4, tie_word_embeddings=False,
                   name_or_path="Qwen/Qwen2.5-7B", **kw)

    @classmethod
    def plumbing_125m(cls, vocab_size=512, **kw):
        return cls(vocab_size=vocab_size, hidden_size=768, intermediate_size=2048, num_hidden_layers=12,
                   num_attention_heads=12, num_key_value_heads=4, tie_word_embeddings=True,
                   name_or_path="synthetic/qwen2-125m", **kw)

    @classmethod
    def tiny(cls, vocab_size=300, **kw):
        return cls(vocab_size=vocab_size, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                   num_attention_heads=4, num_key_value_heads=2, tie_word_embeddings=True,
                   name_or_path="synthetic/qwen2-tiny", **kw)

    def to_dict(self):
        d = asdict(self)
        d["architectures"] = ["Qwen2ForCausalLM"]
        d["torch_dtype"] = "bfloat16"
        return d

    @classmethod
    def from_dict(cls, d):
        keys = {f for f in cls.__dataclass_fields__}
        return cls(**{k: v for k, v in d.items() if k in keys})

    @classmethod
    def from_pretrained(cls, path):
        with open(os.path.join(path, "config.json")) as f:
            return cls.from_dict(json.load(f))


class Linear(nn.Linear):
    """``nn.Linear`` (same parameters, same checkpoint names) whose three contractions -- forward, dgrad, wgrad -- run on
    the general tcgen05 GEMM (ops/gemm.py) instead of cuBLAS."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class RMSNorm(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x):
        return ops.rmsnorm(x, self.weight, self.eps)


class Qwen2Attention(nn.Module):
    def __init__(self, cfg: Qwen2Config):
        super().__init__()
        self.cfg = cfg
        d, hd = cfg.hidden_size, cfg.head_dim
        self.q_proj = Linear(d, cfg.num_attention_heads * hd, bias=cfg.attention_bias)
        self.k_proj = Linear(d, cfg.num_key_value_heads * hd, bias=cfg.attention_bias)
        self.v_proj = Linear(d, cfg.num_key_value_heads * hd, bias=cfg.attention_bias)
        self.o_proj = Linear(cfg.num_attention_heads * hd, d, bias=False)

    def forward(self, x, cos, sin, cu_seqlens, max_seqlen):
        cfg = self.cfg
        T = x.shape[0]
        from .lora import lora_group_forward
        q, k, v = lora_group_forward(x, [self.q_proj, self.k_proj, self.v_proj])        # one rank-r launch when all three carry LoRA
        q = q.view(T, cfg.num_attention_heads, cfg.head_dim)
        k = k.view(T, cfg.num_key_value_heads, cfg.head_dim)
        v = v.view(T, cfg.num_key_value_heads, cfg.head_dim)
        q = ops.apply_rope(q, cos, sin)
        k = ops.apply_rope(k, cos, sin)
        o = ops.attention_varlen(q, k, v, cu_seqlens, max_seqlen, causal=True)
        return self.o_proj(o.reshape(T, -1))


class Qwen2MLP(nn.Module):
    def __init__(self, cfg: Qwen2Config):
        super().__init__()
        self.gate_proj = Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.up_proj = Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.down_proj = Linear(cfg.intermediate_size, cfg.hidden_size, bias=False)

    def forward(self, x):
        from .lora import lora_group_forward
        gate, up = lora_group_forward(x, [self.gate_proj, self.up_proj])
        return self.down_proj(ops.swiglu_pair(gate, up))


class Qwen2DecoderLayer(nn.Module):
    def __init__(self, cfg: Qwen2Config):
        super().__init__()
        self.self_attn = Qwen2Attention(cfg)
        self.mlp = Qwen2MLP(cfg)
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, h, cos, sin, cu_seqlens, max_seqlen):
        h = h + self.self_attn(self.input_layernorm(h), cos, sin, cu_seqlens, max_seqlen)
        h = h + self.mlp(self.post_attention_layernorm(h))
        return h


class Qwen2Model(nn.Module):
    def __init__(self, cfg: Qwen2Config):
        super().__init__()
        self.cfg = cfg
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList([Qwen2DecoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.gradient_checkpointing = False

    def forward(self, input_ids, cu_seqlens, position_ids, max_seqlen=None):
        """input_ids/position_ids: [T] packed; cu_seqlens: [S+1] int32.  Returns hidden [T, d]."""
        h = self.embed_tokens(input_ids)
        cos, sin = ops.ref.rope_cos_sin(position_ids, self.cfg.head_dim, self.cfg.rope_theta)
        ckpt = self.gradient_checkpointing and self.training and torch.is_grad_enabled()
        if ckpt and not h.requires_grad:
            # LoRA-only training: the embedding output has no grad; checkpoint needs one input that
            # does (the reference calls enable_input_require_grads(), GRPO/grpo.py:240-241).
            h = h.detach().requires_grad_(True)
        for layer in self.layers:
            if ckpt:
                h = torch.utils.checkpoint.checkpoint(layer, h, cos, sin, cu_seqlens, max_seqlen,
                                                      use_reentrant=False)
            else:
                h = layer(h, cos, sin, cu_seqlens, max_seqlen)
        return self.norm(h)


def _init_weights(module: nn.Module, std: float = 0.02):
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, 0.0, std)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, 0.0, std)


class Qwen2PreTrained(nn.Module):
    """Shared plumbing: packing helpers, HF-format save/load, gradient checkpointing switch."""

    def __init__(self, cfg: Qwen2Config):
        super().__init__()
        self.config = cfg
        self.name_or_path = cfg.name_or_path

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None):
        self.model.gradient_checkpointing = True

    def gradient_checkpointing_disable(self):
        self.model.gradient_checkpointing = False

    def enable_input_require_grads(self):  # API parity; handled inside Qwen2Model.forward
        return None

    # ---- HF-format checkpoint I/O ------------------------------------------------------------
    def save_pretrained(self, path: str, safe_serialization: bool = True):
        from .hf_io import save_state_dict
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.config.to_dict(), f, indent=2)
        sd = {k: v for k, v in self.state_dict().items()}
        if self.config.tie_word_embeddings and "lm_head.weight" in sd:
            sd.pop("lm_head.weight")
        save_state_dict(sd, os.path.join(path, "model.safetensors"))

    @classmethod
    def from_pretrained(cls, path: str, torch_dtype=torch.bfloat16, device="cpu", **_):
        from .hf_io import load_state_dict
        cfg = Qwen2Config.from_pretrained(path)
        model = cls(cfg)
        sd = load_state_dict(os.path.join(path, "model.safetensors"))
        if cfg.tie_word_embeddings and "lm_head.weight" not in sd and hasattr(model, "lm_head"):
            sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
        missing, unexpected = model.load_state_dict(sd, strict=False)
        missing = [m for m in missing if not m.endswith("score.weight")]
        if missing or unexpected:
            raise RuntimeError(f"checkpoint mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
        model = model.to(device=device, dtype=torch_dtype)
        if hasattr(model, "tie_weights"):
            model.tie_weights()
        return model

    @classmethod
    def from_config(cls, cfg: Qwen2Config, torch_dtype=torch.bfloat16, device="cpu", seed: Optional[int] = None):
        """Random-init (there is no network on the GPU box: SURVEY.md environment facts)."""
        dev = torch.device(device)
        # Billion-parameter models are initialised ON the GPU (seconds instead of ~half a minute of CPU normal_() per
        # model -- start-up time counts against the benchmark driver's per-run limit); small models keep the CPU generator
        # so CPU- and GPU-built test models with one seed stay identical.
        big = dev.type == "cuda" and cfg.num_hidden_layers * cfg.hidden_size * cfg.intermediate_size > 5e7
        if seed is not None:
            gen_state = torch.random.get_rng_state()
            # torch.manual_seed also seeds every CUDA generator -- deliberately NOT undone: default CUDA generators start from a
            # non-deterministic per-process seed, and modules created right after (LoRA adapters) must match across ranks
            torch.manual_seed(seed)
        if big:
            with torch.device(dev):
                model = cls(cfg)
        else:
            model = cls(cfg)
        _init_weights(model)
        if seed is not None:
            torch.random.set_rng_state(gen_state)
        model = model.to(device=device, dtype=torch_dtype)
        if hasattr(model, "tie_weights"):
            model.tie_weights()
        return model


class Qwen2ForCausalLM(Qwen2PreTrained):
    def __init__(self, cfg: Qwen2Config):
        super().__init__(cfg)
        self.model = Qwen2Model(cfg)
        self.lm_head = Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        self.tie_weights()

    def tie_weights(self):
        if self.config.tie_word_embeddings:
            self.lm_head.weight = self.model.embed_tokens.weight

    def untie_weights(self):
        """Give lm_head its own storage (peft ``modules_to_save`` semantics)."""
        if self.lm_head.weight is self.model.embed_tokens.weight:
            self.lm_head.weight = nn.Parameter(self.model.embed_tokens.weight.detach().clone())

    def hidden_states(self, input_ids, cu_seqlens, position_ids, max_seqlen=None):
        return self.model(input_ids, cu_seqlens, position_ids, max_seqlen)

    def logits(self, hidden):
        """Materialised logits -- API-edge / tests / CPU sampler only; never on the hot path."""
        return self.lm_head(hidden)

    def token_logprobs(self, hidden, targets, temperature=1.0, want_entropy=True):
        """(logp, entropy) of ``targets`` under softmax(lm_head(hidden)/temperature); fused K-LP."""
        top_p = getattr(self, "logprob_top_p", None)          # set by the trainer for ``logprob_top_p_consistent=True``
        if top_p is not None and top_p < 1.0:
            from ..ops import reference as _ref
            return _ref.lmhead_logprob_top_p(hidden, self.lm_head.weight, targets, temperature, top_p)
        return ops.lmhead_logprob(hidden, self.lm_head.weight, targets, temperature, want_entropy)

    def forward(self, input_ids, cu_seqlens, position_ids, max_seqlen=None):
        return self.logits(self.hidden_states(input_ids, cu_seqlens, position_ids, max_seqlen))


class Qwen2ForSequenceClassification(Qwen2PreTrained):
    """Backbone + ``score: Linear(d, num_labels, bias=False)`` -- the PPO critic (PPO/ppo.py:280-287)."""

    def __init__(self, cfg: Qwen2Config):
        super().__init__(cfg)
        self.model = Qwen2Model(cfg)
        self.score = nn.Linear(cfg.hidden_size, cfg.num_labels, bias=False)

    def hidden_states(self, input_ids, cu_seqlens, position_ids, max_seqlen=None):
        return self.model(input_ids, cu_seqlens, position_ids, max_seqlen)

    def values(self, hidden):
        return self.score(hidden).squeeze(-1).float()

    def forward(self, input_ids, cu_seqlens, position_ids, max_seqlen=None):
        return self.values(self.hidden_states(input_ids, cu_seqlens, position_ids, max_seqlen))

    @classmethod
    def from_causal_lm(cls, lm: Qwen2ForCausalLM):
        """Critic initialised from a policy backbone (what HF does when loading a CausalLM ckpt)."""
        m = cls(lm.config)
        m.model.load_state_dict(lm.model.state_dict())
        nn.init.normal_(m.score.weight, 0.0, 1.0 / math.sqrt(lm.config.hidden_size + 1))
        return m.to(device=lm.device, dtype=lm.dtype)


# --------------------------------------------------------------------------------------------
# packing helpers: padded [B, T] API edge  <->  packed varlen compute
# --------------------------------------------------------------------------------------------
def pack_padded(query_responses: torch.Tensor, pad_token_id: int) -> Tuple[torch.Tensor, ...]:
    """Turn a padded batch (left-padded queries, right-padded responses) into packed form.

    A token is real iff ``id != pad`` -- the reference's rule (GRPO/grpo_trainer.py:100-105:
    ``attention_mask = ids != pad``; ``position_ids = cumsum(mask) - mask``).
    Returns (input_ids [T], cu_seqlens [B+1] int32, position_ids [T], max_seqlen, flat_index [T])
    where ``flat_index`` are positions into ``query_responses.view(-1)``.
    """
    B, L = query_responses.shape
    mask = query_responses != pad_token_id
    lens = mask.sum(1)
    cu = torch.zeros(B + 1, dtype=torch.int32, device=query_responses.device)
    cu[1:] = lens.cumsum(0)
    flat_index = mask.view(-1).nonzero(as_tuple=False).squeeze(1)
    ids = query_responses.reshape(-1)[flat_index]
    pos = (mask.long().cumsum(1) - 1).reshape(-1)[flat_index]
    return ids, cu, pos, int(lens.max().item()) if B > 0 else 0, flat_index


def forward(model, query_responses: torch.Tensor, pad_token_id: int):
    """Reference-compatible helper (GRPO/grpo_trainer.py:90-120): padded ids in, padded logits out.

    Returns a 1-tuple ``(logits [B, L, V],)`` like HF's ``return_dict=False``.  Rows are computed
    packed; pad positions come back as zeros.  Kept for API parity and tests -- the trainers use
    ``response_logprobs`` which never builds the logits tensor.
    """
    lm = getattr(model, "policy", model)
    B, L = query_responses.shape
    ids, cu, pos, mx, flat = pack_padded(query_responses, pad_token_id)
    logits_packed = lm(ids, cu, pos, mx)
    out = logits_packed.new_zeros(B * L, logits_packed.shape[-1])
    out[flat] = logits_packed
    return (out.view(B, L, -1),)


def build_logprob_plan(query_responses: torch.Tensor, context_length: int, pad_token_id: int) -> dict:
    """Host-synchronising part of ``response_logprobs``: packing indices and the (source row -> response slot) map.

    ``src[i]`` is the packed position whose hidden state predicts ``targets[i]``; the result lands in
    slot ``(r[i], c[i])`` of the [B, T_r] output.
    """
    B, L = query_responses.shape
    ids, cu, pos, mx, flat = pack_padded(query_responses, pad_token_id)
    # packed token at flat position (b, l) predicts target (b, l+1); needed when l+1 >= ctx
    col = flat % L
    row = flat // L
    nxt_flat = torch.empty_like(flat)
    nxt_flat[:-1] = flat[1:]
    nxt_flat[-1] = -1
    # the next *packed* token must be the next column of the same row (responses are contiguous)
    pred = (nxt_flat == flat + 1) & (col + 1 < L) & (col + 1 >= context_length)
    src = pred.nonzero(as_tuple=False).squeeze(1)
    targets = query_responses.reshape(-1)[flat[src] + 1]
    return {"ids": ids, "cu": cu, "pos": pos, "max_seqlen": mx, "flat": flat, "col": col, "row": row, "src": src,
            "targets": targets, "r": row[src], "c": col[src] + 1 - context_length}


def planned_response_logprobs(lm, plan: dict, B: int, T_r: int, temperature: float, want_entropy: bool,
                              invalid_value: float = 1.0, max_seqlen=None, hidden=None):
    """Device-only part of ``response_logprobs`` (no host sync: CUDA-graph capturable).  ``plan`` may be padded:
    extra ``src`` entries must carry ``r == B`` (a dump row that is sliced off)."""
    dev = plan["ids"].device
    if hidden is None:
        hidden = lm.hidden_states(plan["ids"], plan["cu"], plan["pos"], max_seqlen or plan["max_seqlen"])
    logp, ent = lm.token_logprobs(hidden.index_select(0, plan["src"]), plan["targets"], temperature, want_entropy)
    out_lp = torch.full((B + 1, T_r), invalid_value, dtype=torch.float32, device=dev)
    out_lp = out_lp.index_put((plan["r"], plan["c"]), logp)[:B]
    out_ent = torch.zeros((B + 1, T_r), dtype=torch.float32, device=dev)
    if ent is not None:
        out_ent = out_ent.index_put((plan["r"], plan["c"]), ent)
    return out_lp, out_ent[:B], hidden


def response_logprobs(lm, query_responses: torch.Tensor, context_length: int, pad_token_id: int,
                      temperature: float, want_entropy: bool = False, invalid_value: float = 1.0,
                      value_model=None):
    """Per-token log-probs of the response part, [B, T_r], via the fused lm-head op.

    Position j of the result is log p(response[j] | everything before), i.e. the reference's
    ``logits[:, ctx-1:-1] / T -> log_softmax -> gather`` (GRPO/grpo_trainer.py:543-549) evaluated
    only where both the predicting token and the target are real; every other slot is
    ``invalid_value`` (the caller masks those with INVALID_LOGPROB anyway).
    If ``value_model`` is given also returns values [B, T_r] aligned the same way (PPO
    ``get_reward(...)[0][:, ctx-1:-1]``, PPO/ppo_trainer.py:630-634): values[j] is the critic's
    output at column ctx-1+j, so values[seq_len+1] is the value of the post-EOS state.
    Returns (logp, entropy[, values]).
    """
    B, L = query_responses.shape
    T_r = L - context_length
    dev = query_responses.device
    plan = build_logprob_plan(query_responses, context_length, pad_token_id)
    out_lp, out_ent, hidden = planned_response_logprobs(lm, plan, B, T_r, temperature, want_entropy, invalid_value)
    result = [out_lp, out_ent]
    if value_model is not None:
        col, row = plan["col"], plan["row"]
        ids, cu, pos, mx = plan["ids"], plan["cu"], plan["pos"], plan["max_seqlen"]
        vsel = ((col >= context_length - 1) & (col <= L - 2)).nonzero(as_tuple=False).squeeze(1)
        vhidden = hidden if value_model is lm else value_model.hidden_states(ids, cu, pos, mx)
        vals = value_model.values(vhidden[vsel])
        out_v = torch.zeros((B, T_r), dtype=torch.float32, device=dev)
        out_v = out_v.index_put((row[vsel], col[vsel] + 1 - context_length), vals)
        result.append(out_v)
    return tuple(result)
