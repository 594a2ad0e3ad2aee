import pytest
import torch

from nanorlhf_b200.utils import (INVALID_LOGPROB, exact_div, first_true_indices, masked_mean, masked_var, masked_whiten,
                                 response_masks, scatter_terminal_reward, truncate_response)
from nanorlhf_b200.utils.schedules import get_scheduler


def test_masked_mean_and_whiten():
    torch.manual_seed(0)
    x = torch.randn(4, 9)
    m = torch.rand(4, 9) > 0.4
    sel = x[m]
    assert torch.allclose(masked_mean(x, m), sel.mean())
    assert torch.allclose(masked_var(x, m), sel.var(unbiased=True), atol=1e-6)
    w = masked_whiten(x, m)
    assert abs(w[m].mean().item()) < 1e-5 and abs(w[m].var(unbiased=True).item() - 1) < 1e-3
    w2 = masked_whiten(x, m, shift_mean=False)
    assert torch.allclose(w2[m].mean(), sel.mean(), atol=1e-5)


def test_first_true_and_truncate():
    b = torch.tensor([[False, True, True], [False, False, False]])
    assert first_true_indices(b).tolist() == [1, 3]
    r = torch.tensor([[5, 9, 7, 9, 3], [1, 2, 3, 4, 5]])
    t = truncate_response(9, 0, r)
    assert t.tolist() == [[5, 9, 0, 0, 0], [1, 2, 3, 4, 5]]


def test_masks_and_terminal_reward():
    pad = 0
    post = torch.tensor([[4, 4, 9, 0, 0], [4, 4, 4, 4, 4], [9, 0, 0, 0, 0]])
    seq_len, pm, pm1 = response_masks(post, pad)
    assert seq_len.tolist() == [2, 4, 0]
    assert pm[0].tolist() == [False, False, False, True, True]
    assert pm1[0].tolist() == [False, False, False, False, True]
    assert not pm[1].any()
    rew = scatter_terminal_reward(torch.zeros(3, 5), torch.tensor([1.0, 2.0, 3.0]), seq_len)
    # actual_end = seq_len+1 when it fits, else seq_len
    assert rew[0].tolist() == [0, 0, 0, 1, 0] and rew[1].tolist() == [0, 0, 0, 0, 2] and rew[2].tolist() == [0, 3, 0, 0, 0]
    assert INVALID_LOGPROB == 1.0


def test_exact_div():
    assert exact_div(12, 4) == 3
    with pytest.raises(ValueError):
        exact_div(10, 4, "nope")


def test_cosine_with_min_lr():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    s = get_scheduler("cosine_with_min_lr", opt, 0, 10, {"min_lr_rate": 0.1})
    lrs = []
    for _ in range(10):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        s.step()
    assert lrs[0] == 1.0 and all(a >= b for a, b in zip(lrs, lrs[1:]))
    assert abs(opt.param_groups[0]["lr"] - 0.1) < 1e-6


def test_prompt_loader_drop_last_false_keeps_the_tail():
    from nanorlhf_b200.utils.data import PromptLoader
    ds = [{"i": i} for i in range(10)]
    collate = lambda rows: [r["i"] for r in rows]            # noqa: E731
    keep = PromptLoader(ds, 4, collate, seed=0, shuffle=False, drop_last=False)
    drop = PromptLoader(ds, 4, collate, seed=0, shuffle=False, drop_last=True)
    assert len(keep) == 3 and len(drop) == 2
    it = iter(keep)
    seen = [next(it) for _ in range(3)]
    assert seen == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 0, 1]]
    it = iter(drop)
    assert [next(it) for _ in range(3)] == [[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 2, 3]]


def test_max_grad_norm_clips_the_update():
    import torch
    from nanorlhf_b200.parallel.optimizer import FusedAdamW
    w = torch.nn.Parameter(torch.zeros(1024))
    opt = FusedAdamW([{"params": [w]}], lr=1e-2)
    opt.max_grad_norm = 1.0
    opt.zero_grad()
    w.grad.fill_(10.0)                       # norm 320 -> scaled by 1/320
    opt.step()
    assert abs(opt.last_grad_norm - 320.0) < 1e-3
    # Adam's first step is lr * sign(g) whatever the scale; the moments carry the clipped gradient
    assert torch.allclose(opt.flats[0].exp_avg[:1024], torch.full((1024,), 0.1 * 10.0 / 320.0), rtol=1e-4)


def test_top_p_consistent_logprobs():
    """``logprob_top_p_consistent``: log-probs of the top-p-truncated, renormalised softmax (the distribution the sampler draws
    from) -- equal to the full-softmax value plus -log(nucleus mass) inside the nucleus; gradients flow; top_p -> 1 recovers
    the plain log-softmax."""
    import torch
    from nanorlhf_b200.ops import reference as ref
    torch.manual_seed(0)
    T, d, V = 7, 16, 50
    h = torch.randn(T, d, requires_grad=True)
    w = torch.randn(V, d) * 0.5
    z = (h @ w.t()) / 0.9
    p = torch.softmax(z, -1)
    sp, si = p.sort(-1, descending=True)
    keep = torch.zeros_like(p, dtype=torch.bool).scatter_(1, si, (sp.cumsum(-1) - sp) < 0.8)
    tgt_in = si[:, 0]                                                    # the most probable token is always in the nucleus
    lp, ent = ref.lmhead_logprob_top_p(h, w, tgt_in, 0.9, 0.8)
    mass = (p * keep).sum(-1)
    want = torch.log(p.gather(1, tgt_in[:, None]).squeeze(1)) - torch.log(mass)
    assert torch.allclose(lp, want, atol=1e-5) and (lp > torch.log_softmax(z, -1).gather(1, tgt_in[:, None]).squeeze(1)).all()
    pk = (p * keep) / mass[:, None]
    assert torch.allclose(ent, -(pk * torch.log(pk.clamp_min(1e-30))).sum(-1), atol=1e-4)
    lp.sum().backward()
    assert h.grad is not None and torch.isfinite(h.grad).all() and h.grad.abs().sum() > 0
    tgt_out = si[:, -1]                                                  # least probable token: outside -> full-softmax value
    lp_out, _ = ref.lmhead_logprob_top_p(h.detach(), w, tgt_out, 0.9, 0.8)
    assert torch.allclose(lp_out, torch.log_softmax(z.detach(), -1).gather(1, tgt_out[:, None]).squeeze(1), atol=1e-5)
    lp1, _ = ref.lmhead_logprob_top_p(h.detach(), w, tgt_in, 0.9, 1.0)
    assert torch.allclose(lp1, torch.log_softmax(z.detach(), -1).gather(1, tgt_in[:, None]).squeeze(1), atol=1e-5)
