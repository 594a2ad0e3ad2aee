"""Op layer: every hot op of the engine behind one name, two backends.

* ``torch``  -- pure PyTorch (ops/reference.py): CPU plumbing backend and test oracle.
* ``cuda``   -- hand-written sm_100a kernels from ``nanorlhf_b200/csrc`` (ops/native.py).

On a CUDA tensor the native extension is *mandatory* unless ``NANORLHF_BACKEND=torch`` is set
explicitly: a missing ``.so`` raises instead of silently falling back (the round-end driver
records which ``.so`` files were loaded).
"""
from __future__ import annotations

import os

import torch

from . import reference as ref

_FORCED = os.environ.get("NANORLHF_BACKEND", "").lower()
_native = None
_native_err = None


def _nat():
    """Return the loaded native extension module, raising a clear error if it is unavailable."""
    global _native, _native_err
    if _native is None and _native_err is None:
        try:
            from . import native as _n
            _n.load()
            _native = _n
        except Exception as e:  # pragma: no cover - exercised on boxes without the build
            _native_err = e
    if _native is None:
        raise RuntimeError(
            "nanorlhf_b200 native extension is not available "
            f"({_native_err!r}). Build it with `python -m nanorlhf_b200.csrc.build` "
            "or set NANORLHF_BACKEND=torch to run the PyTorch reference ops.")
    return _native


def use_native(t: torch.Tensor) -> bool:
    if _FORCED == "torch":
        return False
    return t.is_cuda


def backend_name(t: torch.Tensor) -> str:
    return "cuda" if use_native(t) else "torch"


# ---- thin dispatchers (autograd-aware wrappers live in ops/native.py) -----------------------
def rmsnorm(x, weight, eps):
    if use_native(x):
        return _nat().rmsnorm(x, weight, eps)
    return ref.rmsnorm(x, weight, eps)


def add_rmsnorm(x, residual, weight, eps):
    if use_native(x):
        return _nat().add_rmsnorm(x, residual, weight, eps)
    return ref.add_rmsnorm(x, residual, weight, eps)


def apply_rope(x, cos, sin):
    if use_native(x):
        return _nat().apply_rope(x, cos, sin)
    return ref.apply_rope(x, cos, sin)


def swiglu(gate_up):
    if use_native(gate_up):
        return _nat().swiglu(gate_up)
    return ref.swiglu(gate_up)


def swiglu_pair(gate, up):
    """silu(gate) * up without concatenating the two projections."""
    if use_native(gate):
        return _nat().swiglu_pair(gate, up)
    return (torch.nn.functional.silu(gate.float()) * up.float()).to(gate.dtype)


def linear(x, weight, bias=None, act: int = 0):
    """``F.linear`` (+ optional fused exact GELU, ``act=1``) -- the general tcgen05 GEMM on CUDA bf16 (ops/gemm.py)."""
    if use_native(x):
        _nat()
        from . import gemm as _lin
        return _lin.linear(x, weight, bias, act)
    y = torch.nn.functional.linear(x, weight, bias)
    return torch.nn.functional.gelu(y) if act == 1 else y


def lora_linear(x, weight, bias, lora_a, lora_b, scaling: float):
    """``x W^T + b + scaling * (x A^T) B^T`` with a frozen ``W``; one dual-source-K GEMM per direction on CUDA bf16."""
    if use_native(x):
        _nat()
        from . import gemm as _lin
        if _lin.lora_supported(x, weight, lora_a, lora_b):
            return _lin.lora_linear(x, weight, bias, lora_a, lora_b, scaling)
    return ref.lora_linear(x, weight, bias, lora_a, lora_b, scaling)


def lora_linear_group(x, projs, scaling: float):
    """LoRA projections that share their input (q/k/v, gate/up): ``projs`` = [(W, bias, A, B), ...] -> list of outputs.
    One rank-r launch for all adapters on CUDA bf16 (ops/gemm.py); independent ``lora_linear`` calls elsewhere."""
    if use_native(x):
        _nat()
        from . import gemm as _lin
        if all(_lin.lora_supported(x, w, a, b) for w, _bias, a, b in projs) and len({a.shape[0] for _w, _b, a, _B in projs}) == 1:
            return _lin.lora_linear_group(x, projs, scaling)
    return [lora_linear(x, w, bias, a, b, scaling) for w, bias, a, b in projs]


def attention_varlen(q, k, v, cu_seqlens, max_seqlen=None, causal=True, scale=None):
    if use_native(q):
        return _nat().attention_varlen(q, k, v, cu_seqlens, max_seqlen, causal, scale)
    return ref.attention_varlen(q, k, v, cu_seqlens, causal=causal, scale=scale)


def lmhead_logprob(hidden, weight, targets, temperature=1.0, want_entropy=True):
    """Autograd-capable fused lm-head log-prob.  Returns (logp, entropy) both fp32 [T]."""
    if use_native(hidden):
        return _nat().lmhead_logprob(hidden, weight, targets, temperature, want_entropy)
    return _LmHeadLogprobTorch.apply(hidden, weight, targets, float(temperature), bool(want_entropy))


class _LmHeadLogprobTorch(torch.autograd.Function):
    """Chunked, recompute-in-backward log-prob head on the PyTorch backend."""

    @staticmethod
    def forward(ctx, hidden, weight, targets, temperature, want_entropy):
        logp, ent, lse = ref.lmhead_logprob(hidden, weight, targets, temperature, want_entropy=want_entropy)
        ctx.save_for_backward(hidden, weight, targets, lse)
        ctx.temperature = temperature
        ctx.mark_non_differentiable(ent)
        return logp, ent

    @staticmethod
    def backward(ctx, g_logp, _g_ent):
        hidden, weight, targets, lse = ctx.saved_tensors
        dh, dw = ref.lmhead_logprob_backward(hidden, weight, targets, lse, g_logp, ctx.temperature,
                                             need_weight_grad=ctx.needs_input_grad[1])
        return dh if ctx.needs_input_grad[0] else None, dw, None, None, None


def discounted_suffix_sum(rewards, gamma=1.0):
    if use_native(rewards):
        return _nat().gae_scan(rewards, None, gamma, 1.0)[0]
    return ref.discounted_suffix_sum(rewards, gamma)


def gae(rewards, values, gamma, lam):
    if use_native(rewards):
        return _nat().gae_scan(rewards, values, gamma, lam)
    return ref.gae(rewards, values, gamma, lam)


def policy_loss_token(new_logp, old_logp, adv, mask, cliprange, ref_logp=None, kl_coef=0.0):
    if use_native(new_logp):
        return _nat().policy_loss_token(new_logp, old_logp, adv, mask, cliprange, ref_logp, kl_coef)
    return ref.policy_loss_token(new_logp, old_logp, adv, mask, cliprange, ref_logp, kl_coef)


policy_loss_sequence = ref.policy_loss_sequence   # [B]-sized math; not a hot path
nll_loss = ref.nll_loss


def value_loss(vpred, values_old, returns, mask, cliprange_value):
    if use_native(vpred):
        return _nat().value_loss(vpred, values_old, returns, mask, cliprange_value)
    return ref.value_loss(vpred, values_old, returns, mask, cliprange_value)
