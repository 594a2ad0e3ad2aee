"""Linear layers of the training / log-prob / reward paths on the general tcgen05 GEMM (csrc/gemm_tc.cu).

What the reference runs here is cuBLAS behind ``nn.Linear`` plus peft's LoRA wrapper -- per projection and direction a base
GEMM, two adapter GEMMs and an add over the ``[tokens, out]`` result (/root/reference/GRPO/grpo.py:228-243; call sites
/root/reference/GRPO/grpo_trainer.py:543-556,652-660).  Here every contraction of the decoder is one launch of the same
sm_100a kernel, with the operand "major-ness" chosen so that no transposed copy is ever made:

=================  =========================================  =====================================================
 product            operands as stored                         kernel form
=================  =========================================  =====================================================
 y  = x W^T (+b)    x [T,K], W [N,K]                           K-major x K-major
 dx = dy W          dy [T,N], W [N,K] = [contraction, out]     K-major x MN-major
 dW = dy^T x        dy [T,N], x [T,K]  (tokens = contraction)  MN-major x MN-major
 LoRA forward       t = s x A^T;  y = [x | t] [W | B]^T        dual-source K: ONE GEMM with K + r columns
 LoRA dgrad         t' = s dy B;  dx = [dy | t'] [W ; A]       dual-source K, MN-major B operands
 LoRA wgrad         dA = t'^T x,  dB = dy^T t                  MN-major x MN-major (rank-r outputs)
=================  =========================================  =====================================================

Tensors whose shape the kernel cannot address (inner extents not multiples of 8, mis-aligned rows) and non-bf16 /
non-CUDA tensors take the PyTorch path; a bf16 CUDA tensor of a supported shape never falls back silently.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import native as _native_mod


def _ok2d(t: torch.Tensor) -> bool:
    return (t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0
            and t.shape[0] > 0 and t.shape[1] % 8 == 0)


def supported(x: torch.Tensor, w: torch.Tensor) -> bool:
    from . import use_native
    return (use_native(x) and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.dim() == 2
            and w.shape[0] % 8 == 0 and w.shape[1] % 8 == 0 and w.is_contiguous() and x.shape[-1] == w.shape[1])


def _rows(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    return x2 if _ok2d(x2) else x2.contiguous()


def gemm(a, b, a_mn=False, b_mn=False, a2=None, b2=None, bias=None, act=0, alpha=1.0, out=None, out_f32=None,
         accumulate=False, cg=0, block_n=0, split_k=-1):
    """Thin wrapper over ``_C.gemm_tc`` (see csrc/bindings.cpp) that counts the launch."""
    _native_mod._count()
    return _native_mod.ext().gemm_tc(a, b, a_mn, b_mn, a2, b2, bias, act, alpha, out, out_f32, accumulate, cg, block_n,
                                     split_k)


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b; forward, dgrad and wgrad on the tcgen05 kernel, no transposed copies."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        x2 = _rows(x)
        y = gemm(x2, w, bias=b, act=act)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        if act != 0:
            ctx.mark_non_differentiable(y)           # the fused activation is inference-only (reward model)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        g2 = _rows(gy)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm(g2, w, b_mn=True).view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            gw = gemm(g2, x2, a_mn=True, b_mn=True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.float().sum(0).to(g2.dtype)
        return gx, gw, gb, None


def linear(x, w, b=None, act: int = 0):
    """``F.linear`` on the native GEMM (``act=1``: exact GELU fused into the epilogue, no-grad use only)."""
    if supported(x, w):
        return _LinearFn.apply(x, w, b, act)
    y = F.linear(x, w, b)
    return F.gelu(y) if act == 1 else y


class _LoraLinearFn(torch.autograd.Function):
    """y = x W^T + b + s (x A^T) B^T with a frozen W: the adapter rides in the base GEMM as extra contraction columns."""

    @staticmethod
    def forward(ctx, x, w, bias, a, b, scaling):
        x2 = _rows(x)
        t = gemm(x2, a, alpha=scaling)                                   # [T, r] = s x A^T
        y = gemm(x2, w, a2=t, b2=b, bias=bias)                           # [x | t] [W | B]^T
        ctx.save_for_backward(x2, w, a, b, t)
        ctx.scaling = scaling
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w, a, b, t = ctx.saved_tensors
        s = ctx.scaling
        g2 = _rows(gy)
        tp = gemm(g2, b, b_mn=True, alpha=s)                             # [T, r] = s dy B
        gx = ga = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm(g2, w, b_mn=True, a2=tp, b2=a).view(ctx.xshape)    # [dy | t'] [W ; A]
        if ctx.needs_input_grad[3]:
            ga = gemm(tp, x2, a_mn=True, b_mn=True)                      # [r, K] = t'^T x      (s is inside t')
        if ctx.needs_input_grad[4]:
            gb = gemm(g2, t, a_mn=True, b_mn=True)                       # [N, r] = dy^T t      (s is inside t)
        return gx, None, None, ga, gb, None


def lora_supported(x, w, a, b) -> bool:
    return (supported(x, w) and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_contiguous()
            and b.is_contiguous() and a.shape[0] % 8 == 0 and a.shape[0] == b.shape[1])


def lora_linear(x, w, bias, a, b, scaling: float):
    return _LoraLinearFn.apply(x, w, bias, a, b, float(scaling))


class _LoraGroupFn(torch.autograd.Function):
    """Several LoRA projections of ONE input (q / k / v, gate / up): the rank-r down-projections share a launch
    (``t = s x [A_1; ...; A_n]^T``), each base GEMM takes its slice of ``t`` as dual-source-K operand, and in the backward the
    adapter-input gradients ``dA_i`` come out of one GEMM over the concatenated ``t'`` while the whole LoRA term of ``dx`` rides
    in the first dgrad GEMM (K2 = n r).  Compared with n independent ``_LoraLinearFn``: n - 1 fewer skinny launches per
    direction, ``x`` is re-read once instead of n times."""

    @staticmethod
    def forward(ctx, x, scaling, n, *flat):
        x2 = _rows(x)
        ws, bs, As, Bs = flat[0::4], flat[1::4], flat[2::4], flat[3::4]
        r = As[0].shape[0]
        a_cat = torch.cat(As, 0)                                             # [n r, K]
        t_cat = gemm(x2, a_cat, alpha=scaling)                               # [T, n r]
        ys = [gemm(x2, ws[i], a2=t_cat[:, i * r:(i + 1) * r], b2=Bs[i], bias=bs[i]) for i in range(n)]
        ctx.save_for_backward(x2, t_cat, a_cat, *ws, *Bs)
        ctx.meta = (scaling, n, r, x.shape)
        return tuple(y.view(*x.shape[:-1], y.shape[-1]) for y in ys)

    @staticmethod
    def backward(ctx, *gys):
        s, n, r, xshape = ctx.meta
        saved = ctx.saved_tensors
        x2, t_cat, a_cat = saved[:3]
        ws, Bs = saved[3:3 + n], saved[3 + n:3 + 2 * n]
        gs = [_rows(g) for g in gys]
        tp_cat = torch.empty(x2.shape[0], n * r, dtype=x2.dtype, device=x2.device)
        for i in range(n):
            gemm(gs[i], Bs[i], b_mn=True, alpha=s, out=tp_cat[:, i * r:(i + 1) * r])       # t'_i = s dy_i B_i
        gx = None
        if ctx.needs_input_grad[0]:
            gx = gemm(gs[0], ws[0], b_mn=True, a2=tp_cat, b2=a_cat)                          # dy_0 W_0 + [t'_1 .. t'_n] [A_1; ..; A_n]
            for i in range(1, n):
                gx = gx + gemm(gs[i], ws[i], b_mn=True)
            gx = gx.view(xshape)
        ga_cat = gemm(tp_cat, x2, a_mn=True, b_mn=True)                                       # [n r, K]
        grads = [gx, None, None]
        for i in range(n):
            gb = gemm(gs[i], t_cat[:, i * r:(i + 1) * r], a_mn=True, b_mn=True)             # [N_i, r]
            grads += [None, None, ga_cat[i * r:(i + 1) * r], gb]
        return tuple(grads)


def lora_linear_group(x, projs, scaling: float):
    """``projs``: [(W, bias | None, A, B), ...] sharing the input ``x`` -> list of outputs."""
    flat = []
    for w, bias, a, b in projs:
        flat += [w, bias, a, b]
    return list(_LoraGroupFn.apply(x, float(scaling), len(projs), *flat))
