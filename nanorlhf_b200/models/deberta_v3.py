"""DeBERTa-v3 sequence classifier (the reward model of the reference's default configs).

The reference loads ``OpenAssistant/reward-model-deberta-v3-large-v2`` through HF
``AutoModelForSequenceClassification`` in fp32 and swaps it GPU<->CPU around every scoring pass
(/root/reference/GRPO/grpo.py:159-198).  This is a from-scratch module with the HF
parameter names (``deberta.encoder.layer.N.attention.self.query_proj.weight`` ...) so the published
checkpoint loads, running in bf16 and staying resident in HBM.

Architecture facts reproduced (DeBERTa-v2/v3 "disentangled attention"):
  * no absolute position embeddings (``position_biased_input=False``), embeddings -> LayerNorm;
  * one shared table of ``2 * position_buckets`` relative-position embeddings, layer-normed, projected
    with the *content* key/query projections (``share_att_key=True``);
  * score = (Qc Kc^T + c2p + p2c) / sqrt(d_head * 3) where c2p gathers Qc Kr^T at the log-bucketed
    relative position and p2c gathers Kc Qr^T at the mirrored position;
  * post-LN transformer blocks, GELU FFN, ContextPooler(dense+GELU on token 0) + Linear(->num_labels).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import asdict, dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .qwen2 import Linear


@dataclass
class DebertaV3Config:
    vocab_size: int = 128100
    hidden_size: int = 1024
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    intermediate_size: int = 4096
    layer_norm_eps: float = 1e-7
    position_buckets: int = 256
    max_relative_positions: int = -1
    max_position_embeddings: int = 512
    pooler_hidden_size: int = 1024
    num_labels: int = 1
    pad_token_id: int = 0
    cls_token_id: int = 1
    sep_token_id: int = 2
    model_type: str = "deberta-v2"

    @classmethod
    def large(cls, **kw):
        return cls(**kw)

    @classmethod
    def tiny(cls, vocab_size=512, **kw):
        return cls(vocab_size=vocab_size, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                   intermediate_size=128, position_buckets=16, max_position_embeddings=64, pooler_hidden_size=64, **kw)

    @classmethod
    def from_pretrained(cls, path):
        with open(os.path.join(path, "config.json")) as f:
            d = json.load(f)
        return cls(**{k: v for k, v in d.items() if k in cls.__dataclass_fields__})


def make_log_bucket_position(rel: torch.Tensor, bucket_size: int, max_position: int) -> torch.Tensor:
    sign = torch.sign(rel)
    mid = bucket_size // 2
    abs_pos = torch.where((rel < mid) & (rel > -mid), torch.full_like(rel, mid - 1), rel.abs())
    log_pos = torch.ceil(torch.log(abs_pos.float() / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    return torch.where(abs_pos <= mid, rel, (log_pos * sign).to(rel.dtype))


def build_relative_position(q_len: int, k_len: int, bucket_size: int, max_position: int, device) -> torch.Tensor:
    q = torch.arange(q_len, device=device)
    k = torch.arange(k_len, device=device)
    rel = q[:, None] - k[None, :]
    if bucket_size > 0 and max_position > 0:
        rel = make_log_bucket_position(rel, bucket_size, max_position)
    return rel.long()


class DisentangledSelfAttention(nn.Module):
    def __init__(self, cfg: DebertaV3Config):
        super().__init__()
        self.h = cfg.num_attention_heads
        self.dh = cfg.hidden_size // cfg.num_attention_heads
        self.query_proj = Linear(cfg.hidden_size, cfg.hidden_size)
        self.key_proj = Linear(cfg.hidden_size, cfg.hidden_size)
        self.value_proj = Linear(cfg.hidden_size, cfg.hidden_size)
        self.span = cfg.position_buckets
        self.scale_factor = 3          # content + c2p + p2c

    def _heads(self, x):               # [B, T, H*dh] -> [B, H, T, dh]
        B, T, _ = x.shape
        return x.view(B, T, self.h, self.dh).transpose(1, 2)

    def forward(self, x, key_mask, rel_pos, rel_emb):
        B, T, _ = x.shape
        q, k, v = self._heads(self.query_proj(x)), self._heads(self.key_proj(x)), self._heads(self.value_proj(x))
        scale = 1.0 / math.sqrt(self.dh * self.scale_factor)
        scores = torch.matmul(q, k.transpose(-1, -2)) * scale
        # relative-position tables share the content projections
        pos_k = self.key_proj(rel_emb).view(-1, self.h, self.dh).transpose(0, 1)      # [H, 2s, dh]
        pos_q = self.query_proj(rel_emb).view(-1, self.h, self.dh).transpose(0, 1)
        idx_c2p = (rel_pos + self.span).clamp(0, 2 * self.span - 1)                    # [T, T]
        c2p = torch.matmul(q, pos_k.transpose(-1, -2)[None])                           # [B, H, T, 2s]
        scores = scores + torch.gather(c2p, -1, idx_c2p[None, None].expand(B, self.h, T, T)) * scale
        idx_p2c = (-rel_pos + self.span).clamp(0, 2 * self.span - 1)
        p2c = torch.matmul(k, pos_q.transpose(-1, -2)[None])                           # [B, H, Tk, 2s]
        scores = scores + torch.gather(p2c, -1, idx_p2c[None, None].expand(B, self.h, T, T)).transpose(-1, -2) * scale
        scores = scores.masked_fill(~key_mask[:, None, None, :], torch.finfo(scores.dtype).min)
        probs = torch.softmax(scores.float(), dim=-1).to(x.dtype)
        probs = probs.masked_fill(~key_mask[:, None, None, :], 0.0)
        out = torch.matmul(probs, v).transpose(1, 2).reshape(B, T, -1)
        return out


    # ---- packed / fused path (CUDA): one flash-style kernel for content + c2p + p2c ---------------------
    def forward_packed(self, x, cu_seqlens, max_len, lut, rel_emb):
        """x: [T_total, hidden] packed real tokens.  c2p and p2c share the index c = bucket(i-j)+span
        (bucketing is odd in its argument), so score = (Qc.Kc + A[i,c] + B[j,c]) * scale with
        A = Qc Kr^T and B = Kc Qr^T -- two small GEMMs feeding the fused attention kernel."""
        from ..ops import native
        T = x.shape[0]
        q = self.query_proj(x).view(T, self.h, self.dh)
        k = self.key_proj(x).view(T, self.h, self.dh)
        v = self.value_proj(x).view(T, self.h, self.dh)
        pos_k = self.key_proj(rel_emb).view(-1, self.h, self.dh).transpose(0, 1)          # [H, NB, dh]
        pos_q = self.query_proj(rel_emb).view(-1, self.h, self.dh).transpose(0, 1)
        # the two bias tables A = Qc Kr^T and B = Kc Qr^T ([H, T, NB]): per head a [T, 64] x [NB, 64]^T product on the general
        # tcgen05 GEMM; the per-head operands are column slices of the packed projections (no copies)
        rel_a, rel_b = _bias_tables(q, pos_k), _bias_tables(k, pos_q)
        scale = 1.0 / math.sqrt(self.dh * self.scale_factor)
        native._count()
        if os.environ.get("NANORLHF_DEBERTA_TMA", "1") != "0" and self.dh == 64:
            # every tile (Q, K, V and the two sliding bias-table windows) arrives by TMA
            out = native.ext().deberta_attn_fwd(q, k, v, cu_seqlens, int(max_len), scale, rel_a, rel_b, lut)
        else:
            out, _ = native.ext().attn_varlen_fwd(q, k, v, cu_seqlens, int(max_len), scale, False, rel_a, rel_b, lut)
        return out.reshape(T, -1)


def _bias_tables(x: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """x: [T, H, dh] (a view of a packed projection), pos: [H, NB, dh] (a transposed view) -> [H, T, NB] with
    out[h] = x[:, h] pos[h]^T."""
    T, H, dh = x.shape
    NB = pos.shape[1]
    if x.is_cuda and x.dtype == torch.bfloat16 and ops.use_native(x) and dh % 8 == 0 and NB % 8 == 0:
        from ..ops import native
        native._count()
        # one batched launch: 3D TMA maps address the per-head [T, dh] / [NB, dh] slices in place
        return native.ext().gemm_tc_batched(x.transpose(0, 1), pos, None)
    return torch.bmm(x.transpose(0, 1), pos.transpose(1, 2)).contiguous()


def build_bucket_lut(max_len: int, bucket_size: int, max_position: int, span: int, device) -> torch.Tensor:
    """int16 table: delta -> clamp(bucket(delta) + span, 0, 2*span-1) for |delta| <= max_len + 64.  The margin
    covers the padded tail rows / keys of the last attention tiles, so the kernel indexes it without clamping."""
    max_len = max_len + 64
    d = torch.arange(-max_len, max_len + 1, device=device)
    b = make_log_bucket_position(d, bucket_size, max_position) if (bucket_size > 0 and max_position > 0) else d
    return (b + span).clamp(0, 2 * span - 1).to(torch.int16).contiguous()


def _fused_inference(x: torch.Tensor) -> bool:
    """bf16 CUDA tensors outside autograd take the fused kernels (GEMM+bias+GELU epilogue, add+LayerNorm)."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and not torch.is_grad_enabled()
            and os.environ.get("NANORLHF_DEBERTA", "fused") != "eager")


def _add_layernorm(ln: nn.LayerNorm, y: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    if _fused_inference(y) and y.shape[-1] <= 2048 and y.shape[-1] % 8 == 0:
        from ..ops import native
        return native.add_layernorm(y.contiguous(), residual.contiguous(), ln.weight, ln.bias, ln.eps)
    return ln(y + residual)


class _SelfOutput(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = Linear(cfg.hidden_size, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, cfg.layer_norm_eps)

    def forward(self, h, residual):
        return _add_layernorm(self.LayerNorm, self.dense(h), residual)


class _Attention(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.self = DisentangledSelfAttention(cfg)
        self.output = _SelfOutput(cfg)

    def forward(self, x, key_mask, rel_pos, rel_emb):
        return self.output(self.self(x, key_mask, rel_pos, rel_emb), x)


class _Intermediate(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = Linear(cfg.hidden_size, cfg.intermediate_size)

    def forward(self, x):
        # GEMM + bias + GELU epilogue (tcgen05 kernel, act=1).  Measured on B200 at 26.5k x 4096 x 1024 the erf in the
        # epilogue makes the tile epilogue longer than its 16 k-block main loop (0.41 ms vs cuBLAS 0.16 + GELU 0.12),
        # so it is opt-in until the epilogue is split across more warps.
        if _fused_inference(x) and x.shape[-1] % 8 == 0 and os.environ.get("NANORLHF_DEBERTA_GELU_FUSED", "0") == "1":
            return ops.linear(x, self.dense.weight, self.dense.bias, act=1)
        return F.gelu(self.dense(x))


class _Output(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = Linear(cfg.intermediate_size, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, cfg.layer_norm_eps)

    def forward(self, h, residual):
        return _add_layernorm(self.LayerNorm, self.dense(h), residual)


class _Layer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.attention = _Attention(cfg)
        self.intermediate = _Intermediate(cfg)
        self.output = _Output(cfg)

    def forward(self, x, key_mask, rel_pos, rel_emb):
        a = self.attention(x, key_mask, rel_pos, rel_emb)
        return self.output(self.intermediate(a), a)


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=cfg.pad_token_id)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, cfg.layer_norm_eps)

    def forward(self, ids, mask):
        return self.LayerNorm(self.word_embeddings(ids)) * mask[..., None].to(self.LayerNorm.weight.dtype)


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.layer = nn.ModuleList([_Layer(cfg) for _ in range(cfg.num_hidden_layers)])
        self.rel_embeddings = nn.Embedding(cfg.position_buckets * 2, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, cfg.layer_norm_eps)

    def forward(self, x, mask):
        cfg = self.cfg
        T = x.shape[1]
        max_rel = cfg.max_relative_positions if cfg.max_relative_positions > 0 else cfg.max_position_embeddings
        rel_pos = build_relative_position(T, T, cfg.position_buckets, max_rel, x.device)
        rel_emb = self.LayerNorm(self.rel_embeddings.weight)
        for layer in self.layer:
            x = layer(x, mask, rel_pos, rel_emb)
        return x


class _Backbone(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)

    def forward(self, ids, mask):
        return self.encoder(self.embeddings(ids, mask), mask)


class _Pooler(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = Linear(cfg.hidden_size, cfg.pooler_hidden_size)

    def forward(self, h):
        return F.gelu(self.dense(h[:, 0]))


class DebertaV3ForSequenceClassification(nn.Module):
    def __init__(self, cfg: DebertaV3Config):
        super().__init__()
        self.config = cfg
        self.deberta = _Backbone(cfg)
        self.pooler = _Pooler(cfg)
        self.classifier = Linear(cfg.pooler_hidden_size, cfg.num_labels)

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Returns logits [B, num_labels]."""
        mask = (input_ids != self.config.pad_token_id) if attention_mask is None else attention_mask.bool()
        if self._can_fuse(input_ids):
            return self._forward_packed(input_ids, mask)
        h = self.deberta(input_ids, mask)
        return self.classifier(self.pooler(h)).float()

    def _can_fuse(self, input_ids) -> bool:
        if not input_ids.is_cuda or next(self.parameters()).dtype != torch.bfloat16:
            return False
        if os.environ.get("NANORLHF_BACKEND", "") == "torch" or os.environ.get("NANORLHF_DEBERTA", "") == "eager":
            return False
        cfg = self.config
        return cfg.hidden_size // cfg.num_attention_heads == 64 and (cfg.position_buckets * 2) % 8 == 0

    def _forward_packed(self, input_ids, mask):
        """Padding-free forward: all token-wise layers run on the packed real tokens, attention in the fused
        disentangled flash kernel (csrc/attention_varlen.cu, REL_BIAS)."""
        cfg = self.config
        lens = mask.sum(1)
        cu = torch.zeros(mask.shape[0] + 1, dtype=torch.int32, device=input_ids.device)
        cu[1:] = lens.cumsum(0)
        max_len = input_ids.shape[1]
        flat = mask.reshape(-1).nonzero(as_tuple=False).squeeze(1)
        ids = input_ids.reshape(-1)[flat]
        emb = self.deberta.embeddings
        x = emb.LayerNorm(emb.word_embeddings(ids))
        enc = self.deberta.encoder
        max_rel = cfg.max_relative_positions if cfg.max_relative_positions > 0 else cfg.max_position_embeddings
        key = (max_len, str(input_ids.device))
        if getattr(self, "_lut_key", None) != key:
            self._lut = build_bucket_lut(max_len, cfg.position_buckets, max_rel, cfg.position_buckets, input_ids.device)
            self._lut_key = key
        rel_emb = enc.LayerNorm(enc.rel_embeddings.weight)
        for layer in enc.layer:
            a = layer.attention.self.forward_packed(x, cu, max_len, self._lut, rel_emb)
            a = layer.attention.output(a, x)
            x = layer.output(layer.intermediate(a), a)
        first = x[cu[:-1].long()]
        return self.classifier(F.gelu(self.pooler.dense(first))).float()

    @classmethod
    def from_config(cls, cfg, torch_dtype=torch.bfloat16, device="cpu", seed: Optional[int] = None):
        st = torch.random.get_rng_state()
        dev = torch.device(device)
        big = dev.type == "cuda" and cfg.num_hidden_layers * cfg.hidden_size * cfg.intermediate_size > 5e7   # see qwen2.from_config
        if seed is not None:
            torch.manual_seed(seed)
        if big:
            with torch.device(dev):
                m = cls(cfg)
        else:
            m = cls(cfg)
        for mod in m.modules():
            if isinstance(mod, (nn.Linear, nn.Embedding)):
                nn.init.normal_(mod.weight, 0.0, 0.02)
                if getattr(mod, "bias", None) is not None:
                    nn.init.zeros_(mod.bias)
        torch.random.set_rng_state(st)
        return m.to(device=device, dtype=torch_dtype)

    @classmethod
    def from_pretrained(cls, path, torch_dtype=torch.bfloat16, device="cpu"):
        from .hf_io import load_state_dict
        cfg = DebertaV3Config.from_pretrained(path)
        m = cls(cfg)
        f = os.path.join(path, "model.safetensors")
        sd = load_state_dict(f) if os.path.exists(f) else torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        missing, unexpected = m.load_state_dict(sd, strict=False)
        if missing:
            raise RuntimeError(f"DeBERTa checkpoint is missing {missing[:5]}")
        return m.to(device=device, dtype=torch_dtype)

    def save_pretrained(self, path):
        from .hf_io import save_state_dict
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(asdict(self.config), f, indent=2)
        save_state_dict(self.state_dict(), os.path.join(path, "model.safetensors"))
