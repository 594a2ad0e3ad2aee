"""K-BC: policy -> sampler weight refresh, on device, no filesystem.

Reference: every rollout the reference saves the adapter, reloads the base model
on the CPU, runs ``merge_and_unload`` there, writes the merged model to /data/temp_vllm_model and
boots vLLM from it (/root/reference/GRPO/grpo_trainer.py:131-141): 3.1 GB written+read twice plus
an engine boot per update.

Here the sampler keeps a fused arena (qkv / gate_up concatenated) and ``refresh_sampler_arena``
rewrites it from the live training parameters:

* LoRA layers:    W' = W + (alpha/r) * B @ A  -- the ``EPI_MERGE`` epilogue of the tcgen05 GEMM
                  (csrc/gemm_sm100.cu): the rank-r product runs on the tensor cores (contraction = r), W is streamed
                  into the epilogue, the merged tile goes straight into the arena -- one pass over W;
* plain layers:   device copy into the fused layout;
* data parallel:  ``ShardedWeightSync`` -- the arena lives in symmetric memory bound to an NVLS multicast object; rank r
                  merges only the layers with ``layer % world == r`` and its epilogue writes every merged tile with
                  ``multimem.st`` to the multicast address, i.e. into ALL ranks' arenas at once (the NVSwitch replicates
                  the store).  Per rank: 1/world of the merge math and of the HBM reads, no NCCL broadcast, no per-peer
                  recompute.  Without a multicast object (or on one GPU) it degenerates to the local merge.
* fp8 rollout:    ``rollout_dtype="fp8"`` re-quantises the refreshed arena (e4m3, one scale per output channel,
                  csrc/quant.cu) in a second pass -- the per-channel amax needs the whole merged row, which spans
                  several epilogue tiles.
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
    """K-BC across ranks (see the module docstring).  ``refresh`` = barrier, sharded merge with multicast stores,
    barrier; plain (non-LoRA) layers, biases and norms are replicated locally."""

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
        self.mc_base = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        # re-point the sampler's arena at the symmetric buffer
        for li, lw in enumerate(sampler.layers):
            off = li * per_layer
            for name, (a, b) in self.shapes:
                getattr(lw, name)  # must exist
                setattr(lw, name, self.flat[off:off + a * b].view(a, b))
                off += a * b
        self.drain_before_barrier = False
        import os
        self.barrier_timeout_ms = int(float(os.environ.get("NANORLHF_BARRIER_TIMEOUT_S", "900")) * 1000)
        self.stats = {"refreshes": 0, "multicast": bool(self.mc_base)}

    def _mc_addr(self, view: torch.Tensor) -> int:
        return self.mc_base + (view.data_ptr() - self.flat.data_ptr()) if self.mc_base else 0

    def _peer_view(self, peer: int, li: int, name: str):
        off = li * self.per_layer
        for n, (a, b) in self.shapes:
            if n == name:
                return self.hdl.get_buffer(peer, (a, b), torch.bfloat16, off)
            off += a * b
        raise KeyError(name)

    def _write(self, mod, li: int, name: str, rows: slice):
        """Merged weight of ``mod`` into rows ``rows`` of arena matrix ``name`` of layer ``li`` on EVERY rank."""
        local = getattr(self.sampler.layers[li], name)[rows]
        if isinstance(mod, LoraLinear) and self.mc_base:
            native._count()
            native.ext().lora_merge(mod.base_layer.weight, mod.lora_A.weight, mod.lora_B.weight, float(mod.scaling), local,
                                    self._mc_addr(local))
            return
        # no multicast object (or a plain layer that only this rank was asked to write): peer stores over NVLink
        for peer in range(self.comm.world_size):
            _merged(mod, self._peer_view(peer, li, name)[rows])

    @torch.no_grad()
    def refresh(self):
        s, W, R = self.sampler, self.comm.world_size, self.comm.rank
        cfg = s.cfg
        D = cfg.head_dim
        nq, nkv, F = cfg.num_attention_heads * D, cfg.num_key_value_heads * D, cfg.intermediate_size
        self.hdl.barrier(channel=0, timeout_ms=self.barrier_timeout_ms)                         # nobody is still sampling from the old arena
        for li in range(cfg.num_hidden_layers):
            layer, lw = s.lm.model.layers[li], s.layers[li]
            at, mlp = layer.self_attn, layer.mlp
            parts = ((at.q_proj, "wqkv", slice(0, nq)), (at.k_proj, "wqkv", slice(nq, nq + nkv)), (at.v_proj, "wqkv", slice(nq + nkv, nq + 2 * nkv)),
                     (at.o_proj, "wo", slice(None)), (mlp.gate_proj, "wgu", slice(0, F)), (mlp.up_proj, "wgu", slice(F, 2 * F)),
                     (mlp.down_proj, "wdown", slice(None)))
            for mod, name, rows in parts:
                if isinstance(mod, LoraLinear):
                    if li % W == R:                          # this rank's share of the merge work
                        self._write(mod, li, name, rows)
                else:
                    getattr(lw, name)[rows].copy_(mod.weight)     # frozen everywhere: every rank already holds it
            # biases / norms are tiny and replicated: local copies
            for mod, sl in ((at.q_proj, slice(0, nq)), (at.k_proj, slice(nq, nq + nkv)), (at.v_proj, slice(nq + nkv, nq + 2 * nkv))):
                b = _bias(mod)
                if b is not None:
                    lw.bqkv[sl].copy_(b)
            lw.ln1, lw.ln2 = layer.input_layernorm.weight, layer.post_attention_layernorm.weight
        if self.drain_before_barrier:
            # belt and braces for the multicast path: make sure this rank's kernels (and their posted NVLink / NVLS writes)
            # have fully retired before it signals the closing barrier
            torch.cuda.current_stream(s.device).synchronize()
        self.hdl.barrier(channel=1, timeout_ms=self.barrier_timeout_ms)                         # every rank's (multicast) stores have landed
        for lw in s.layers:
            interleave_gate_up(lw, F)
        self.stats["refreshes"] += 1
