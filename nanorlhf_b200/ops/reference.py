"""Pure-PyTorch reference implementations of every native op.

These are (a) the CPU / plumbing backend (``NANORLHF_BACKEND=torch``) and (b) the oracle the
CUDA kernels in ``csrc/`` are tested against (tests/test_kernels_gpu.py).  Everything here is
written from the behaviour described in SURVEY.md section 3.5; citations point at the reference
call sites the semantics come from.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# normalisation / rotary / activation
# --------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (y * weight.float()).to(x.dtype)


def add_rmsnorm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float):
    """residual' = residual + x ; y = rmsnorm(residual').  Returns (y, residual')."""
    r = (residual.float() + x.float()).to(x.dtype)
    return rmsnorm(r, weight, eps), r


def rope_cos_sin(positions: torch.Tensor, head_dim: int, theta: float, dtype=torch.float32):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=positions.device, dtype=torch.float32) / head_dim))
    ang = positions.float()[:, None] * inv[None, :]
    return ang.cos().to(dtype), ang.sin().to(dtype)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x: [tokens, heads, head_dim]; half-rotation (HF "rotate_half") convention."""
    d2 = x.shape[-1] // 2
    xf = x.float()
    x1, x2 = xf[..., :d2], xf[..., d2:]
    c, s = cos[:, None, :].float(), sin[:, None, :].float()
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(x.dtype)


def swiglu(gate_up: torch.Tensor) -> torch.Tensor:
    """gate_up: [..., 2*ffn] laid out [gate | up] -> silu(gate) * up."""
    g, u = gate_up.chunk(2, dim=-1)
    return (F.silu(g.float()) * u.float()).to(gate_up.dtype)


# --------------------------------------------------------------------------------------------
# attention (packed varlen, causal, GQA)
# --------------------------------------------------------------------------------------------
def attention_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens: torch.Tensor,
                     causal: bool = True, scale: Optional[float] = None) -> torch.Tensor:
    """q: [T, Hq, D], k/v: [T, Hkv, D], cu_seqlens: [S+1].  fp32 softmax, output in q.dtype."""
    T, Hq, D = q.shape
    Hkv = k.shape[1]
    g = Hq // Hkv
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    out = torch.empty_like(q)
    cs = cu_seqlens.tolist()
    for s, e in zip(cs[:-1], cs[1:]):
        if e == s:
            continue
        qs = q[s:e].transpose(0, 1).float()                                    # [Hq, L, D]
        ks = k[s:e].transpose(0, 1).float().repeat_interleave(g, dim=0)        # [Hq, L, D]
        vs = v[s:e].transpose(0, 1).float().repeat_interleave(g, dim=0)
        att = torch.matmul(qs, ks.transpose(1, 2)) * scale
        if causal:
            L = e - s
            att = att.masked_fill(torch.ones(L, L, dtype=torch.bool, device=q.device).triu(1), float("-inf"))
        out[s:e] = torch.matmul(att.softmax(-1), vs).transpose(0, 1).to(q.dtype)
    return out


def paged_attention_decode(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                           block_tables: torch.Tensor, context_lens: torch.Tensor,
                           scale: Optional[float] = None) -> torch.Tensor:
    """Single-token decode against a paged KV cache.

    q: [S, Hq, D]; k_cache/v_cache: [num_blocks, Hkv, block_size, D];
    block_tables: [S, max_blocks] int32; context_lens: [S] (number of valid KV tokens incl. current).
    """
    S, Hq, D = q.shape
    nb, Hkv, bs, _ = k_cache.shape
    g = Hq // Hkv
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    out = torch.empty_like(q)
    for i in range(S):
        L = int(context_lens[i])
        nblk = (L + bs - 1) // bs
        blocks = block_tables[i, :nblk].long()
        ks = k_cache[blocks].permute(1, 0, 2, 3).reshape(Hkv, nblk * bs, D)[:, :L].float()
        vs = v_cache[blocks].permute(1, 0, 2, 3).reshape(Hkv, nblk * bs, D)[:, :L].float()
        ks = ks.repeat_interleave(g, dim=0)
        vs = vs.repeat_interleave(g, dim=0)
        att = torch.einsum("hd,hld->hl", q[i].float(), ks) * scale
        out[i] = torch.einsum("hl,hld->hd", att.softmax(-1), vs).to(q.dtype)
    return out


# --------------------------------------------------------------------------------------------
# fused lm-head log-prob (K-LP): never materialise [tokens, vocab] for the caller
# --------------------------------------------------------------------------------------------
def lmhead_logprob(hidden: torch.Tensor, weight: torch.Tensor, targets: torch.Tensor,
                   temperature: float = 1.0, chunk: int = 2048,
                   want_entropy: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """logp[t] = log_softmax(hidden[t] @ weight.T / temperature)[targets[t]] (+ entropy, lse).

    Semantics of GRPO/grpo_trainer.py:544-549 (divide by temperature *before* log-softmax) and the
    entropy stat of :678-679, computed chunk-wise in fp32.
    Returns (logp [T] fp32, entropy [T] fp32, lse [T] fp32 of the temperature-scaled logits).
    """
    T = hidden.shape[0]
    logp = torch.empty(T, dtype=torch.float32, device=hidden.device)
    ent = torch.empty(T, dtype=torch.float32, device=hidden.device)
    lse = torch.empty(T, dtype=torch.float32, device=hidden.device)
    inv_t = 1.0 / temperature
    for s in range(0, T, chunk):
        z = (hidden[s:s + chunk] @ weight.t()).float() * inv_t
        l = torch.logsumexp(z, dim=-1)
        lse[s:s + chunk] = l
        logp[s:s + chunk] = z.gather(1, targets[s:s + chunk, None].long()).squeeze(1) - l
        if want_entropy:
            p = torch.exp(z - l[:, None])
            ent[s:s + chunk] = l - (p * z).sum(-1)
    if not want_entropy:
        ent.zero_()
    return logp, ent, lse


def lmhead_logprob_top_p(hidden: torch.Tensor, weight: torch.Tensor, targets: torch.Tensor, temperature: float, top_p: float,
                         chunk: int = 1024) -> Tuple[torch.Tensor, torch.Tensor]:
    """Log-probs under the distribution the sampler actually draws from: softmax(z / T) restricted to its top-p nucleus and
    renormalised -- ``logp[t] = z_y - logsumexp_{v in nucleus(t)} z_v``.  The reference samples with top_p = 0.95 but scores
    with the full softmax (/root/reference/GRPO/grpo_trainer.py:127 vs :547-549); this is the consistent variant
    (``logprob_top_p_consistent=True``).  Plain autograd over [chunk, V] logits: an experimental switch, not a fast path.  The
    nucleus is treated as a constant set; a target outside it (never produced by the sampler) gets the full-softmax value.
    Returns (logp [T], entropy of the truncated distribution [T])."""
    inv_t = 1.0 / temperature
    outs, ents = [], []
    for s in range(0, hidden.shape[0], chunk):
        z = (hidden[s:s + chunk] @ weight.t()).float() * inv_t
        with torch.no_grad():
            p = torch.softmax(z, -1)
            sp, si = p.sort(-1, descending=True)
            keep_sorted = (sp.cumsum(-1) - sp) < top_p                     # torch / vLLM nucleus rule; the top token always stays
            keep = torch.zeros_like(p, dtype=torch.bool).scatter_(1, si, keep_sorted)
            tgt = targets[s:s + chunk, None].long()
            inside = keep.gather(1, tgt).squeeze(1)
        zk = z.masked_fill(~keep, float("-inf"))
        lse_k, lse_full = torch.logsumexp(zk, -1), torch.logsumexp(z, -1)
        zy = z.gather(1, tgt).squeeze(1)
        outs.append(torch.where(inside, zy - lse_k, zy - lse_full))
        pk = torch.exp(zk - lse_k[:, None])
        ents.append(lse_k - (pk * z.masked_fill(~keep, 0.0)).sum(-1))
    return torch.cat(outs), torch.cat(ents)


def lmhead_logprob_backward(hidden: torch.Tensor, weight: torch.Tensor, targets: torch.Tensor,
                            lse: torch.Tensor, grad_logp: torch.Tensor, temperature: float,
                            need_weight_grad: bool, chunk: int = 2048):
    """d logp / d hidden and d weight, recomputing logits chunk-wise (cut-cross-entropy style)."""
    T = hidden.shape[0]
    inv_t = 1.0 / temperature
    dh = torch.empty_like(hidden)
    dw = torch.zeros_like(weight, dtype=torch.float32) if need_weight_grad else None
    for s in range(0, T, chunk):
        h = hidden[s:s + chunk]
        z = (h @ weight.t()).float() * inv_t
        p = torch.exp(z - lse[s:s + chunk, None])
        g = grad_logp[s:s + chunk, None].float()
        dz = -p * g
        dz.scatter_add_(1, targets[s:s + chunk, None].long(), g.expand(-1, 1).contiguous())
        dz = (dz * inv_t).to(hidden.dtype)
        dh[s:s + chunk] = dz @ weight
        if need_weight_grad:
            dw += (dz.t() @ h).float()
    return dh, (dw.to(weight.dtype) if dw is not None else None)


# --------------------------------------------------------------------------------------------
# advantage scans (K-GAE)
# --------------------------------------------------------------------------------------------
def discounted_suffix_sum(rewards: torch.Tensor, gamma: float = 1.0) -> torch.Tensor:
    """A_t = r_t + gamma * A_{t+1} (ref loop: GRPO/grpo_trainer.py:611-617, REINFORCE :583-588)."""
    out = torch.empty_like(rewards)
    run = torch.zeros_like(rewards[:, 0])
    for t in range(rewards.shape[1] - 1, -1, -1):
        run = rewards[:, t] + gamma * run
        out[:, t] = run
    return out


def gae(rewards: torch.Tensor, values: torch.Tensor, gamma: float, lam: float):
    """delta_t = r_t + gamma V_{t+1} - V_t; A_t = delta_t + gamma lam A_{t+1}; returns = A + V.

    ref: PPO/ppo_trainer.py:688-697.
    """
    B, T = rewards.shape
    adv = torch.empty_like(rewards)
    last = torch.zeros_like(rewards[:, 0])
    for t in range(T - 1, -1, -1):
        nextv = values[:, t + 1] if t < T - 1 else torch.zeros_like(values[:, 0])
        delta = rewards[:, t] + gamma * nextv - values[:, t]
        last = delta + gamma * lam * last
        adv[:, t] = last
    return adv, adv + values


# --------------------------------------------------------------------------------------------
# policy / value losses (K-LOSS)
# --------------------------------------------------------------------------------------------
def policy_loss_token(new_logp, old_logp, adv, mask, cliprange: float,
                      ref_logp=None, kl_coef: float = 0.0):
    """Token-level PPO-clip surrogate (+ optional GRPO k3-KL).  mask=True where the token counts.

    ref: GRPO/grpo_trainer.py:662-671 (GRPO), REINFORCE/reinforce_trainer.py:634-640.
    Returns (loss, stats dict of detached 0-d tensors).
    """
    m = mask.to(new_logp.dtype)
    diff = new_logp - old_logp
    ratio = torch.exp(diff)
    l1 = -adv * ratio
    l2 = -adv * torch.clamp(ratio, 1.0 - cliprange, 1.0 + cliprange)
    per_tok = torch.max(l1, l2)
    refkl = None
    if ref_logp is not None:
        k = new_logp - ref_logp
        refkl = k.detach()
        per_tok = per_tok + kl_coef * (torch.exp(-k) + k - 1.0)
    denom = m.sum()
    loss = (per_tok * m).sum() / denom
    with torch.no_grad():
        stats = {
            "clipfrac": ((l2 > l1).to(m.dtype) * m).sum() / denom,
            "approxkl_masked": 0.5 * ((diff * diff) * m).sum() / denom,
            "approxkl_all": 0.5 * (diff * diff).mean(),
            "ratio_mean_all": ratio.mean(),
            "ratio_mean_masked": (ratio * m).sum() / denom,
        }
        if refkl is not None:
            stats["refkl_all"] = refkl.mean()
            stats["refkl_masked"] = (refkl * m).sum() / denom
    return loss, stats


def policy_loss_sequence(new_logp, old_logp, adv_seq, cliprange: float):
    """RLOO sequence-level ratio: rho = exp(sum_t new - sum_t old) (ref: RLOO/rloo_trainer.py:660-669).

    The caller has already filled masked positions of both log-prob tensors with INVALID_LOGPROB so
    they cancel in the difference.
    """
    diff = new_logp.sum(1) - old_logp.sum(1)
    ratio = torch.exp(diff)
    l1 = -adv_seq * ratio
    l2 = -adv_seq * torch.clamp(ratio, 1.0 - cliprange, 1.0 + cliprange)
    loss = torch.max(l1, l2).mean()
    with torch.no_grad():
        stats = {
            "clipfrac": (l2 > l1).float().mean(),
            "approxkl_all": 0.5 * (diff * diff).mean(),
            "approxkl_masked": 0.5 * (diff * diff).mean(),
            "ratio_mean_all": ratio.mean(),
            "ratio_mean_masked": ratio.mean(),
        }
    return loss, stats


def nll_loss(new_logp) -> torch.Tensor:
    """RAFT: -mean_b sum_t logp (masked slots hold the constant INVALID_LOGPROB; ref raft_trainer.py:636)."""
    return -new_logp.sum(1).mean()


def value_loss(vpred, values_old, returns, mask, cliprange_value: float):
    """0.5 * masked_mean(max((v-R)^2, (clip(v, v_old +- eps)-R)^2)) (ref: PPO/ppo_trainer.py:742-748)."""
    m = mask.to(vpred.dtype)
    vclip = torch.max(torch.min(vpred, values_old + cliprange_value), values_old - cliprange_value)
    l1 = (vpred - returns) ** 2
    l2 = (vclip - returns) ** 2
    denom = m.sum()
    loss = 0.5 * (torch.max(l1, l2) * m).sum() / denom
    with torch.no_grad():
        clipfrac = ((l2 > l1).to(m.dtype) * m).sum() / denom
    return loss, clipfrac


# --------------------------------------------------------------------------------------------
# optimizer / sampling
# --------------------------------------------------------------------------------------------
def adamw_step_(p, g, m, v, lr, beta1, beta2, eps, wd, step, grad_scale: float = 1.0, master=None):
    """In-place decoupled AdamW on one tensor; moments may be fp32 while p is bf16.  ``master`` (fp32, same shape)
    is the authoritative copy when given: p becomes its rounded image."""
    gf = g.float() * grad_scale
    m.mul_(beta1).add_(gf.to(m.dtype), alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gf.to(v.dtype), gf.to(v.dtype), value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    pf = p.float() if master is None else master
    if wd != 0.0:
        pf = pf * (1 - lr * wd)
    denom = (v.float() / bc2).sqrt_().add_(eps)
    pf = pf - (lr / bc1) * (m.float() / denom)
    if master is not None:
        master.copy_(pf)
    p.copy_(pf.to(p.dtype))


def top_p_sample(logits: torch.Tensor, temperature: float, top_p: float,
                 generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Temperature + nucleus sampling; temperature == 0 -> argmax (ReMax baseline / eval)."""
    if temperature == 0.0:
        return logits.argmax(-1)
    probs = torch.softmax(logits.float() / temperature, dim=-1)
    if top_p < 1.0:
        sp, si = probs.sort(dim=-1, descending=True)
        cum = sp.cumsum(-1)
        drop = (cum - sp) >= top_p          # keep the smallest prefix whose mass reaches top_p
        sp = sp.masked_fill(drop, 0.0)
        sp = sp / sp.sum(-1, keepdim=True)
        choice = torch.multinomial(sp, 1, generator=generator)
        return si.gather(1, choice).squeeze(1)
    return torch.multinomial(probs, 1, generator=generator).squeeze(1)


class _LoraLinearRef(torch.autograd.Function):
    """PyTorch oracle / CPU backend of ``ops.lora_linear``: the rank-r update folded into the base GEMM's output
    (``addmm_`` with beta = 1), autograd keeps only x and the [T, r] projection; the base weight is frozen."""

    @staticmethod
    def forward(ctx, x, w, bias, a, b, scaling):
        x2 = x.reshape(-1, x.shape[-1])
        t = x2 @ a.t()
        y = torch.addmm(bias, x2, w.t()) if bias is not None else x2 @ w.t()
        y.addmm_(t, b.t(), alpha=scaling)
        ctx.save_for_backward(x2, w, a, b, t)
        ctx.scaling = scaling
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w, a, b, t = ctx.saved_tensors
        s = ctx.scaling
        g = gy.reshape(-1, gy.shape[-1])
        gt = g @ b                                    # [T, r]
        gx = None
        if ctx.needs_input_grad[0]:
            gx = g @ w
            gx.addmm_(gt, a, alpha=s)
            gx = gx.view(ctx.xshape)
        ga = (gt.t() @ x2) * s if ctx.needs_input_grad[3] else None
        gb = (g.t() @ t) * s if ctx.needs_input_grad[4] else None
        return gx, None, None, ga, gb, None


def lora_linear(x, w, bias, a, b, scaling: float):
    return _LoraLinearRef.apply(x, w, bias, a, b, scaling)
