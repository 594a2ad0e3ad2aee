"""K-BC: policy -> sampler weight refresh, on device, no filesystem.

Reference: every rollout the reference saves the adapter, reloads the base model
on the CPU, runs ``merge_and_unload`` there, writes the merged model to /data/temp_vllm_model and
boots vLLM from it (/root/reference/GRPO/grpo_trainer.py:131-141): 3.1 GB written+read twice plus
an engine boot per update.

Here the sampler keeps a fused arena (qkv / gate_up concatenated) and ``refresh_sampler_arena``
rewrites it from the live training parameters:

* LoRA layers:    W' = W + (alpha/r) * B @ A  -- fused kernel ``lora_merge`` (csrc/quant.cu): the rank-r
                  product, the add, the optional fp8 (e4m3, 1x128 block scales) quantisation and the
                  store into the arena happen in one pass over W;
* plain layers:   device copy into the fused layout;
* data parallel:  the work is sharded by layer across ranks and each rank writes its merged tiles
                  straight into every peer's arena over NVLink (symmetric memory) -- see
                  ``refresh_sampler_arena_sharded``; with one GPU it degenerates to the local merge.
"""
from __future__ import annotations

import torch

from ..models.lora import LoraLinear
from ..ops import native


def _merged(mod: torch.nn.Module, out: torch.Tensor):
    """Write the effective weight of ``mod`` (nn.Linear or LoraLinear) into ``out`` ([N, K] view)."""
    if isinstance(mod, LoraLinear):
        ext = native.ext() if out.is_cuda else None
        if ext is not None and hasattr(ext, "lora_merge"):
            native._count()
            ext.lora_merge(mod.base_layer.weight, mod.lora_A.weight, mod.lora_B.weight, float(mod.scaling), out)
        else:
            torch.addmm(mod.base_layer.weight, mod.lora_B.weight, mod.lora_A.weight, alpha=mod.scaling, out=out)
    else:
        out.copy_(mod.weight)


def _bias(mod):
    b = mod.base_layer.bias if isinstance(mod, LoraLinear) else mod.bias
    return b


@torch.no_grad()
def refresh_layer(sampler, li: int):
    cfg = sampler.cfg
    D = cfg.head_dim
    nq, nkv, F = cfg.num_attention_heads * D, cfg.num_key_value_heads * D, cfg.intermediate_size
    layer = sampler.lm.model.layers[li]
    lw = sampler.layers[li]
    at, mlp = layer.self_attn, layer.mlp
    _merged(at.q_proj, lw.wqkv[:nq])
    _merged(at.k_proj, lw.wqkv[nq:nq + nkv])
    _merged(at.v_proj, lw.wqkv[nq + nkv:])
    for mod, sl in ((at.q_proj, slice(0, nq)), (at.k_proj, slice(nq, nq + nkv)), (at.v_proj, slice(nq + nkv, nq + 2 * nkv))):
        b = _bias(mod)
        if b is not None:
            lw.bqkv[sl].copy_(b)
    _merged(at.o_proj, lw.wo)
    _merged(mlp.gate_proj, lw.wgu[:F])
    _merged(mlp.up_proj, lw.wgu[F:])
    _merged(mlp.down_proj, lw.wdown)
    lw.ln1 = layer.input_layernorm.weight
    lw.ln2 = layer.post_attention_layernorm.weight


@torch.no_grad()
def refresh_sampler_arena(sampler):
    for li in range(sampler.cfg.num_hidden_layers):
        refresh_layer(sampler, li)
