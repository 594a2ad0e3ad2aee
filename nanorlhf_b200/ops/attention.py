"""Packed-varlen causal GQA attention on CUDA (training / log-prob / prefill path).

Native kernels: ``csrc/attention_fwd_tc.cu`` / ``csrc/attention_bwd_tc.cu`` (tcgen05 MMAs with TMEM accumulators,
TMA-fed, warp-specialised; head_dim 128); other head sizes and ``NANORLHF_ATTN_TC=0`` use
``csrc/attention_varlen.cu`` (mma.sync flash attention, also the oracle of the tcgen05 kernels).  ``NANORLHF_ATTN=flash_attn`` runs the flash-attn
library instead; that path is the *baseline* this framework replaces
(reference: attn_implementation="flash_attention_2", /root/reference/GRPO/grpo.py:219).
"""
from __future__ import annotations

import math
import os

import torch

_IMPL = os.environ.get("NANORLHF_ATTN", "auto")
_USE_TC = os.environ.get("NANORLHF_ATTN_TC", "1") != "0"


def _native_available() -> bool:
    from . import native
    return hasattr(native.ext(), "attn_varlen_fwd")


class _NativeAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens, max_seqlen, scale):
        from . import native
        native._count()
        if _USE_TC and q.shape[-1] == 128:
            o, lse = native.ext().attn_fwd_tc(q, k, v, cu_seqlens, float(scale))
        else:
            o, lse = native.ext().attn_varlen_fwd(q, k, v, cu_seqlens, int(max_seqlen), float(scale))
        ctx.save_for_backward(q, k, v, o, lse, cu_seqlens)
        ctx.max_seqlen, ctx.scale = int(max_seqlen), float(scale)
        return o

    @staticmethod
    def backward(ctx, do):
        from . import native
        q, k, v, o, lse, cu = ctx.saved_tensors
        native._count(3)
        if _USE_TC and q.shape[-1] == 128 and q.is_contiguous() and k.is_contiguous() and v.is_contiguous():
            dq, dk, dv = native.ext().attn_bwd_tc(do.contiguous(), q, k, v, o, lse, cu, ctx.scale)
        else:
            dq, dk, dv = native.ext().attn_varlen_bwd(do.contiguous(), q, k, v, o, lse, cu, ctx.max_seqlen, ctx.scale)
        return dq, dk, dv, None, None, None


def attention_varlen(q, k, v, cu_seqlens, max_seqlen=None, causal=True, scale=None):
    assert causal, "only causal attention is used by the decoder"
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    if max_seqlen is None:
        max_seqlen = int((cu_seqlens[1:] - cu_seqlens[:-1]).max().item())
    impl = _IMPL
    if impl == "auto":
        impl = "native" if (_native_available() and q.shape[-1] == 128 and q.dtype == torch.bfloat16) else "flash_attn"
    if impl == "native":
        cu = cu_seqlens.to(torch.int32)
        needs_grad = torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad)
        if not needs_grad and _USE_TC and q.shape[-1] == 128:
            # inference (rollout prefill, log-prob passes): the TMA maps take strided q/k/v views as they are
            from . import native
            native._count()
            return native.ext().attn_fwd_tc(q, k, v, cu, float(scale))[0]
        return _NativeAttn.apply(q.contiguous(), k.contiguous(), v.contiguous(), cu, max_seqlen, scale)
    if impl == "flash_attn":
        from flash_attn import flash_attn_varlen_func
        cu = cu_seqlens.to(torch.int32)
        return flash_attn_varlen_func(q, k, v, cu, cu, max_seqlen, max_seqlen, softmax_scale=scale, causal=True)
    from . import reference as ref
    return ref.attention_varlen(q, k, v, cu_seqlens, causal=True, scale=scale)
