"""In-process rollout engine: the replacement for the reference's ``vllm_generate``.

Reference (/root/reference/GRPO/grpo_trainer.py:122-166): move the policy to the CPU, merge LoRA on
the CPU, write the model to disk twice, boot a fresh vLLM engine from disk, generate, tear the
engine down, move the policy back.  Here the sampler lives in the trainer's process and CUDA
context and shares HBM with it; "weight sync" is a fused on-device kernel (parallel/weight_sync.py)
and nothing touches the host or the filesystem.

``generate`` keeps the reference's call shape -- ``generate([n,] model, tokenizer, prompts,
temperature, max_tokens)`` returning a right-padded ``LongTensor[len(prompts)*n, max_tokens]`` in
prompt-major / sample-minor order -- but takes prompt *ids* (the reference round-trips ids ->
strings -> ids) and exposes ``top_p`` / ``seed``.
"""
from __future__ import annotations

import random
from typing import List, Optional, Sequence, Union

import torch

from .torch_sampler import torch_generate

_SEED_RNG = random.Random(42)   # the reference seeds python's RNG with 42 and draws randint(1,5000)


def next_rollout_seed() -> int:
    """The "changing seed" feature (grpo_trainer.py:127): a deterministic stream of per-rollout seeds."""
    return _SEED_RNG.randint(1, 5000)


def seed_stream_state():
    return _SEED_RNG.getstate()


def set_seed_stream_state(state):
    _SEED_RNG.setstate(state)


def reseed_stream(seed: int):
    _SEED_RNG.seed(seed)


def prompts_from_padded(queries: torch.Tensor, pad_token_id: int) -> List[List[int]]:
    """Left-padded [B, ctx] tensor -> list of id lists without pads."""
    rows = queries.tolist()
    return [[t for t in r if t != pad_token_id] for r in rows]


def _pick_backend(model, requested: str) -> str:
    if requested in ("torch", "native"):
        return requested
    dev = next(model.parameters()).device
    return "native" if dev.type == "cuda" else "torch"


def get_native_engine(model, **kw):
    """One persistent native engine per policy object (no boot/teardown per rollout).  The engine hangs off the
    policy module itself, so its weight arena and KV pages are released with the model (no process-global cache);
    ``release_native_engine`` frees them earlier."""
    from .native_sampler import NativeSampler
    pol = getattr(model, "policy", model)
    eng = pol.__dict__.get("_nrl_sampler")
    comm = kw.pop("weight_sync_comm", None)
    if eng is None or eng.rollout_dtype != kw.get("rollout_dtype", eng.rollout_dtype):
        eng = NativeSampler(model, **kw)
        pol.__dict__["_nrl_sampler"] = eng        # plain attribute: not a sub-module, not in state_dict
        if comm is not None and comm.world_size > 1:
            from ..parallel.weight_sync import ShardedWeightSync
            eng.sharded_sync = ShardedWeightSync(eng, comm)       # collective: every rank builds its engine at the same point
    return eng


def release_native_engine(model):
    """Drop the resident sampler of ``model`` (weight arena, KV pages, CUDA graphs)."""
    pol = getattr(model, "policy", model)
    eng = pol.__dict__.pop("_nrl_sampler", None)
    if eng is not None:
        eng.release()


def generate(n: int, model, tokenizer, prompts: Union[torch.Tensor, Sequence[Sequence[int]]], temperature: float,
             max_tokens: int, top_p: float = 0.95, seed: Optional[int] = None, backend: str = "auto",
             eos_token_id: Optional[int] = None, rollout_dtype: str = "bf16", **engine_kw) -> torch.Tensor:
    pad_id = tokenizer.pad_token_id
    eos_id = tokenizer.eos_token_id if eos_token_id is None else eos_token_id
    if isinstance(prompts, torch.Tensor):
        prompts = prompts_from_padded(prompts, pad_id)
    if seed is None:
        seed = next_rollout_seed()
    which = _pick_backend(model, backend)
    if which == "native":
        eng = get_native_engine(model, rollout_dtype=rollout_dtype, **engine_kw)
        eng.sync_weights()
        return eng.generate(prompts, n, temperature, top_p, max_tokens, eos_id, pad_id, seed)
    return torch_generate(model, prompts, n, temperature, top_p, max_tokens, eos_id, pad_id, seed)


def vllm_generate(*args, **kwargs) -> torch.Tensor:
    """Drop-in name of the reference's rollout helper.  Both call shapes are accepted:
    ``vllm_generate(n, model, tokenizer, prompts, temperature, max_tokens)`` (GRPO / RLOO / RAFT,
    /root/reference/GRPO/grpo_trainer.py:122) and ``vllm_generate(model, tokenizer, prompts, temperature, max_tokens)``
    (PPO / REINFORCE / ReMax, /root/reference/PPO/ppo_trainer.py:132).  There is no vLLM behind it: the resident
    in-process sampler is used, nothing is written to disk and no engine is booted."""
    if args and isinstance(args[0], int) and not isinstance(args[0], bool):
        return generate(*args, **kwargs)
    return generate(1, *args, **kwargs)
