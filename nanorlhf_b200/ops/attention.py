"""Packed-varlen causal GQA attention on CUDA (training / log-prob / prefill path).

Native kernels: ``csrc/attention_varlen.cu`` (forward + backward, mma.sync tensor-core flash
attention).  Until those are validated on hardware the op can run on the flash-attn library
(``NANORLHF_ATTN=flash_attn``); that path is the *baseline* this framework replaces
(reference: attn_implementation="flash_attention_2", /root/reference/GRPO/grpo.py:219).
"""
from __future__ import annotations

import math
import os

import torch

_IMPL = os.environ.get("NANORLHF_ATTN", "auto")


def _native_available() -> bool:
    from . import native
    return hasattr(native.ext(), "attn_varlen_fwd")


class _NativeAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens, max_seqlen, scale):
        from . import native
        native._count()
        o, lse = native.ext().attn_varlen_fwd(q, k, v, cu_seqlens, int(max_seqlen), float(scale))
        ctx.save_for_backward(q, k, v, o, lse, cu_seqlens)
        ctx.max_seqlen, ctx.scale = int(max_seqlen), float(scale)
        return o

    @staticmethod
    def backward(ctx, do):
        from . import native
        q, k, v, o, lse, cu = ctx.saved_tensors
        native._count(3)
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
        return _NativeAttn.apply(q.contiguous(), k.contiguous(), v.contiguous(), cu_seqlens.to(torch.int32), max_seqlen, scale)
    if impl == "flash_attn":
        from flash_attn import flash_attn_varlen_func
        cu = cu_seqlens.to(torch.int32)
        return flash_attn_varlen_func(q, k, v, cu, cu, max_seqlen, max_seqlen, softmax_scale=scale, causal=True)
    from . import reference as ref
    return ref.attention_varlen(q, k, v, cu_seqlens, causal=True, scale=scale)
