"""CUDA backend of the op layer: thin autograd wrappers over ``nanorlhf_b200/_C.so``.

``load()`` imports the in-tree extension (built by ``python -m nanorlhf_b200.csrc.build``).  It is
not JIT-compiled at import: the ``.so`` must travel with the source tree (gpurun snapshot), and a
missing build on a GPU box is a hard error (ops/__init__.py), never a silent PyTorch fallback.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import reference as ref

_C = None
LAUNCHES = 0          # number of native kernel launches issued through this module (bench "gpu_launches")


def load():
    global _C
    if _C is None:
        from .. import _C as ext          # noqa: F401  (ImportError propagates with the real reason)
        _C = ext
    return _C


def ext():
    return load()


def _count(n: int = 1):
    global LAUNCHES
    LAUNCHES += n


def launches() -> int:
    return LAUNCHES


# --------------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------------
def add_layernorm(x, residual, weight, bias, eps):
    """LayerNorm(x + residual) in one pass (inference; reward model)."""
    _count()
    return ext().add_layernorm(x, residual, weight, bias, float(eps))


def gemm_bf16(a, b, bias=None, out=None, block_n: int = 0, act: int = 0):
    """a[M,K] @ b[N,K]^T (+bias) on the tcgen05 kernel; inputs must be bf16, K-contiguous."""
    _count()
    return ext().gemm_bf16(a, b, bias, out, block_n, act)


def linear(x, w, b=None, act: int = 0):
    """y = x W^T + b with forward / dgrad / wgrad on the general tcgen05 GEMM (ops/gemm.py)."""
    from . import gemm as _lin
    return _lin.linear(x, w, b, act)


# --------------------------------------------------------------------------------------------
# norms / rope / activation
# --------------------------------------------------------------------------------------------
class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        need = x.requires_grad or w.requires_grad
        _count()
        y, _, rstd = ext().rmsnorm(x2, w, eps, None, need)
        if need:
            ctx.save_for_backward(x2, w, rstd)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        x2, w, rstd = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1]).contiguous()
        _count()
        gx = ext().rmsnorm_bwd(x2, w, g2, rstd).view(gy.shape)
        gw = None
        if ctx.needs_input_grad[1]:
            gw = (g2.float() * (x2.float() * rstd[:, None])).sum(0).to(w.dtype)
        return gx, gw, None


def rmsnorm(x, weight, eps):
    if x.dtype != torch.bfloat16:
        return ref.rmsnorm(x, weight, eps)
    return _RMSNormFn.apply(x, weight, eps)


def add_rmsnorm(x, residual, weight, eps):
    """Inference-only fused residual add + RMSNorm: returns (y, residual + x)."""
    _count()
    y, res, _ = ext().rmsnorm(x.contiguous(), weight, eps, residual.contiguous(), False)
    return y, res


class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos, sin):
        ctx.save_for_backward(cos, sin)
        _count()
        return ext().rope(x, cos, sin, 1.0, False)

    @staticmethod
    def backward(ctx, gy):
        cos, sin = ctx.saved_tensors
        _count()
        return ext().rope(gy.contiguous(), cos, sin, -1.0, False), None, None


def apply_rope(x, cos, sin):
    if x.dtype != torch.bfloat16:
        return ref.apply_rope(x, cos, sin)
    if x.stride(-1) != 1 or x.stride(1) != x.shape[2]:
        x = x.contiguous()
    return _RopeFn.apply(x, cos.float().contiguous(), sin.float().contiguous())


class _SwigluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gu):
        g2 = gu.reshape(-1, gu.shape[-1]).contiguous()
        ctx.save_for_backward(g2)
        _count()
        return ext().swiglu(g2).view(*gu.shape[:-1], gu.shape[-1] // 2)

    @staticmethod
    def backward(ctx, go):
        (g2,) = ctx.saved_tensors
        _count()
        return ext().swiglu_bwd(g2, go.reshape(-1, go.shape[-1]).contiguous()).view(*go.shape[:-1], go.shape[-1] * 2)


def swiglu(gate_up):
    if gate_up.dtype != torch.bfloat16:
        return ref.swiglu(gate_up)
    return _SwigluFn.apply(gate_up)


class _SwigluPairFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up):
        g2 = gate.reshape(-1, gate.shape[-1]).contiguous()
        u2 = up.reshape(-1, up.shape[-1]).contiguous()
        ctx.save_for_backward(g2, u2)
        _count()
        return ext().swiglu_pair(g2, u2).view(gate.shape)

    @staticmethod
    def backward(ctx, go):
        g2, u2 = ctx.saved_tensors
        _count()
        dg, du = ext().swiglu_pair_bwd(g2, u2, go.reshape(-1, go.shape[-1]).contiguous())
        return dg.view(go.shape), du.view(go.shape)


def swiglu_pair(gate, up):
    if gate.dtype != torch.bfloat16:
        return (F.silu(gate.float()) * up.float()).to(gate.dtype)
    return _SwigluPairFn.apply(gate, up)


# --------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------
def attention_varlen(q, k, v, cu_seqlens, max_seqlen=None, causal=True, scale=None):
    from . import attention
    return attention.attention_varlen(q, k, v, cu_seqlens, max_seqlen, causal, scale)


# --------------------------------------------------------------------------------------------
# fused lm-head log-prob (K-LP)
# --------------------------------------------------------------------------------------------
class _LmHeadLogprobFn(torch.autograd.Function):
    """Forward: tcgen05 GEMM with an online-softmax epilogue (no [T,V] tensor).
    Backward: the same mainloop recomputes the logits and its epilogue emits dZ in bf16 (one 8192-row chunk at a
    time); dH = dZ W (K-major x MN-major) and dW += dZ^T H (MN-major x MN-major, fp32 accumulate across chunks in the
    epilogue) run on the general tcgen05 GEMM -- W and H are consumed as stored, nothing is transposed."""

    CHUNK = 8192      # rows of dZ materialised at a time in backward (8192 x 152k bf16 = 2.5 GB)

    @staticmethod
    def forward(ctx, hidden, weight, targets, temperature, want_entropy):
        h = hidden.contiguous()
        t32 = targets.to(torch.int32).contiguous()
        _count(2)
        logp, ent, lse = ext().lmhead_logprob_fwd(h, weight, t32, 1.0 / temperature, 0)
        ctx.save_for_backward(h, weight, t32, lse)
        ctx.temperature = temperature
        ctx.mark_non_differentiable(ent)
        return logp, ent

    @staticmethod
    def backward(ctx, g_logp, _g_ent):
        h, weight, t32, lse = ctx.saved_tensors
        inv_t = 1.0 / ctx.temperature
        g = g_logp.float().contiguous()
        need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dh = torch.empty_like(h) if need_h else None
        dw = torch.zeros(weight.shape, dtype=torch.float32, device=weight.device) if need_w else None
        T = h.shape[0]
        from . import gemm as _lin
        for s in range(0, T, _LmHeadLogprobFn.CHUNK):
            e = min(T, s + _LmHeadLogprobFn.CHUNK)
            _count()
            dz = ext().lmhead_dlogits(h[s:e], weight, t32[s:e], lse[s:e], g[s:e], inv_t)
            if need_h:
                _lin.gemm(dz, weight, b_mn=True, out=dh[s:e])
            if need_w:
                _lin.gemm(dz, h[s:e], a_mn=True, b_mn=True, out_f32=dw, accumulate=True)
        return dh, (dw.to(weight.dtype) if need_w else None), None, None, None


def lmhead_logprob(hidden, weight, targets, temperature=1.0, want_entropy=True):
    if hidden.dtype != torch.bfloat16 or hidden.shape[0] == 0:
        logp, ent, _ = ref.lmhead_logprob(hidden, weight, targets, temperature, want_entropy=want_entropy)
        return logp, ent
    return _LmHeadLogprobFn.apply(hidden, weight, targets, float(temperature), bool(want_entropy))


# --------------------------------------------------------------------------------------------
# RL kernels
# --------------------------------------------------------------------------------------------
def gae_scan(rewards, values, gamma, lam):
    _count()
    adv, ret = ext().gae_scan(rewards.float().contiguous(), None if values is None else values.float().contiguous(),
                              float(gamma), float(lam))
    return adv, ret


class _PolicyLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, new_lp, old_lp, adv, mask, ref_lp, cliprange, kl_coef):
        _count()
        grad, acc = ext().policy_loss(new_lp.contiguous(), old_lp.contiguous(), adv.contiguous(), mask.contiguous(),
                                      None if ref_lp is None else ref_lp.contiguous(), cliprange, kl_coef)
        denom = acc[1]
        ctx.save_for_backward(grad, denom)
        ctx.shape = new_lp.shape
        ctx.mark_non_differentiable(acc)
        return acc[0] / denom, acc

    @staticmethod
    def backward(ctx, g_loss, _g_acc):
        grad, denom = ctx.saved_tensors
        return (grad * (g_loss / denom)).view(ctx.shape), None, None, None, None, None, None


def policy_loss_token(new_logp, old_logp, adv, mask, cliprange, ref_logp=None, kl_coef=0.0):
    loss, acc = _PolicyLossFn.apply(new_logp.float(), old_logp.float(), adv.float(), mask, None if ref_logp is None
                                    else ref_logp.float(), float(cliprange), float(kl_coef))
    n = float(new_logp.numel())
    with torch.no_grad():
        d = acc[1]
        stats = {"clipfrac": acc[2] / d, "approxkl_masked": 0.5 * acc[3] / d, "approxkl_all": 0.5 * acc[4] / n,
                 "ratio_mean_all": acc[5] / n, "ratio_mean_masked": acc[6] / d}
        if ref_logp is not None:
            stats["refkl_all"] = acc[7] / n
            stats["refkl_masked"] = acc[8] / d
    return loss, stats


class _ValueLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vpred, vold, ret, mask, clip):
        _count()
        grad, acc = ext().value_loss(vpred.contiguous(), vold.contiguous(), ret.contiguous(), mask.contiguous(), clip)
        ctx.save_for_backward(grad, acc[1])
        ctx.shape = vpred.shape
        ctx.mark_non_differentiable(acc)
        return 0.5 * acc[0] / acc[1], acc

    @staticmethod
    def backward(ctx, g_loss, _g):
        grad, denom = ctx.saved_tensors
        return (grad * (g_loss / denom)).view(ctx.shape), None, None, None, None


def value_loss(vpred, values_old, returns, mask, cliprange_value):
    loss, acc = _ValueLossFn.apply(vpred.float(), values_old.float(), returns.float(), mask, float(cliprange_value))
    with torch.no_grad():
        clipfrac = acc[2] / acc[1]
    return loss, clipfrac


def adamw_flat(param, grad, m, v, lr, beta1, beta2, eps, wd, step, scale=1.0, master=None):
    _count()
    ext().adamw_flat(param, grad, m, v, lr, beta1, beta2, eps, wd, step, scale, master)


# --------------------------------------------------------------------------------------------
# sampler kernels
# --------------------------------------------------------------------------------------------
def sample(logits, temperature, top_p, seed, step, row_ids=None, row_steps=None, out=None, impl=0):
    _count()
    return ext().sample(logits, float(temperature), float(top_p), int(seed), int(step), row_ids, row_steps, out, int(impl))


def kv_cache_write(k, v, k_cache, v_cache, slot_mapping, src_index=None):
    _count()
    ext().kv_cache_write(k, v, k_cache, v_cache, slot_mapping, src_index)


def paged_decode(q, k_cache, v_cache, block_tables, context_lens, scale=None, splits=1, out=None):
    _count()
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    return ext().paged_decode(q, k_cache, v_cache, block_tables, context_lens, float(scale), int(splits), out)
