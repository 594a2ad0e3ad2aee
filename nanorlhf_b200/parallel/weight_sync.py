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


def interleave_gate_up(lw, F: int):
    """[gate | up] -> per 32 features [32 gate rows | 32 up rows]: the layout the fused SwiGLU GEMM epilogue
    (EPI_SWIGLU) expects, so silu(gate)*up happens in registers and the [tokens, 2F] tensor never exists."""
    if getattr(lw, "wgu_i", None) is None or F % 32 != 0:
        return
    d = lw.wgu.shape[1]
    lw.wgu_i.view(F // 32, 2, 32, d).copy_(lw.wgu.view(2, F // 32, 32, d).transpose(0, 1))


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
    interleave_gate_up(lw, F)
    _merged(mlp.down_proj, lw.wdown)
    lw.ln1 = layer.input_layernorm.weight
    lw.ln2 = layer.post_attention_layernorm.weight


@torch.no_grad()
def refresh_sampler_arena(sampler):
    for li in range(sampler.cfg.num_hidden_layers):
        refresh_layer(sampler, li)


# ------------------------------------------------------------------------------------------------
# data-parallel variant: layer-sharded merge + peer stores over NVLink
# ------------------------------------------------------------------------------------------------
class ShardedWeightSync:
    """K-BC across ranks: rank r merges (W + (alpha/r) B A) only for layers with ``layer % world == r`` and the
    GEMM epilogue's TMA stores land directly in *every* rank's sampler arena (the arenas live in symmetric
    memory, so a peer's arena is an ordinary global address over NVLink).  Per rank: 1/world of the merge
    math and HBM reads, (world-1)/world of the arena arrives over NVLink instead of being recomputed.

    Measured trade-off: when every rank already holds the full training weights a
    *local* merge moves fewer bytes than receiving merged weights over NVLink, so ``refresh_sampler_arena``
    stays the default; this path is for configurations where the merge inputs are sharded.
    """

    def __init__(self, sampler, comm):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.sampler, self.comm = sampler, comm
        self.symm_mem, self.group = symm_mem, dist.group.WORLD
        cfg = sampler.cfg
        D = cfg.head_dim
        nq, nkv, F, d = cfg.num_attention_heads * D, cfg.num_key_value_heads * D, cfg.intermediate_size, cfg.hidden_size
        self.shapes = [("wqkv", (nq + 2 * nkv, d)), ("wo", (d, nq)), ("wgu", (2 * F, d)), ("wdown", (d, F))]
        per_layer = sum(a * b for _, (a, b) in self.shapes)
        self.per_layer = per_layer
        self.flat = symm_mem.empty(per_layer * cfg.num_hidden_layers, dtype=torch.bfloat16, device=sampler.device)
        self.hdl = symm_mem.rendezvous(self.flat, self.group.group_name)
        # re-point the sampler's arena at the symmetric buffer
        for li, lw in enumerate(sampler.layers):
            off = li * per_layer
            for name, (a, b) in self.shapes:
                getattr(lw, name)  # must exist
                setattr(lw, name, self.flat[off:off + a * b].view(a, b))
                off += a * b

    def _peer_view(self, peer: int, li: int, name: str):
        off = li * self.per_layer
        for n, (a, b) in self.shapes:
            if n == name:
                return self.hdl.get_buffer(peer, (a, b), torch.bfloat16, off)
            off += a * b
        raise KeyError(name)

    @torch.no_grad()
    def refresh(self):
        s, W, R = self.sampler, self.comm.world_size, self.comm.rank
        cfg = s.cfg
        D = cfg.head_dim
        nq, nkv, F = cfg.num_attention_heads * D, cfg.num_key_value_heads * D, cfg.intermediate_size
        self.hdl.barrier(channel=0)                         # nobody is still sampling from the old arena
        for li in range(cfg.num_hidden_layers):
            layer, lw = s.lm.model.layers[li], s.layers[li]
            at, mlp = layer.self_attn, layer.mlp
            if li % W == R:
                for peer in range(W):
                    wqkv = self._peer_view(peer, li, "wqkv")
                    _merged(at.q_proj, wqkv[:nq])
                    _merged(at.k_proj, wqkv[nq:nq + nkv])
                    _merged(at.v_proj, wqkv[nq + nkv:])
                    _merged(at.o_proj, self._peer_view(peer, li, "wo"))
                    wgu = self._peer_view(peer, li, "wgu")
                    _merged(mlp.gate_proj, wgu[:F])
                    _merged(mlp.up_proj, wgu[F:])
                    _merged(mlp.down_proj, self._peer_view(peer, li, "wdown"))
            # biases / norms are tiny and replicated: local copies
            for mod, sl in ((at.q_proj, slice(0, nq)), (at.k_proj, slice(nq, nq + nkv)), (at.v_proj, slice(nq + nkv, nq + 2 * nkv))):
                b = _bias(mod)
                if b is not None:
                    lw.bqkv[sl].copy_(b)
            lw.ln1, lw.ln2 = layer.input_layernorm.weight, layer.post_attention_layernorm.weight
        self.hdl.barrier(channel=1)                         # every peer's stores have landed
        for lw in s.layers:
            interleave_gate_up(lw, F)
