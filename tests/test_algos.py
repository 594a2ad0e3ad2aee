"""Advantage estimators and losses against slow oracles written straight from SURVEY.md section 3.5."""
import math

import pytest
import torch

from nanorlhf_b200.ops import reference as ref


def test_suffix_sum_and_gae_oracle():
    torch.manual_seed(0)
    r, v = torch.randn(3, 7), torch.randn(3, 7)
    out = ref.discounted_suffix_sum(r, 0.9)
    for b in range(3):
        run = 0.0
        for t in reversed(range(7)):
            run = r[b, t].item() + 0.9 * run
            assert abs(out[b, t].item() - run) < 1e-5
    adv, ret = ref.gae(r, v, 0.99, 0.95)
    for b in range(3):
        last = 0.0
        for t in reversed(range(7)):
            nv = v[b, t + 1].item() if t < 6 else 0.0
            delta = r[b, t].item() + 0.99 * nv - v[b, t].item()
            last = delta + 0.99 * 0.95 * last
            assert abs(adv[b, t].item() - last) < 1e-5
    assert torch.allclose(ret, adv + v)


def test_policy_loss_matches_formula_and_grad():
    torch.manual_seed(0)
    new = (torch.randn(3, 6, dtype=torch.float64) * 0.2).requires_grad_(True)
    old = new.detach() + 0.3 * torch.randn(3, 6, dtype=torch.float64)
    refl = new.detach() + 0.1 * torch.randn(3, 6, dtype=torch.float64)
    adv = torch.randn(3, 6, dtype=torch.float64)
    mask = torch.rand(3, 6) > 0.3
    loss, st = ref.policy_loss_token(new, old, adv, mask, 0.2, refl, 0.05)
    tot, n = 0.0, 0
    for i in range(3):
        for j in range(6):
            if mask[i, j]:
                ratio = math.exp(new[i, j].item() - old[i, j].item())
                l = max(-adv[i, j].item() * ratio, -adv[i, j].item() * min(max(ratio, 0.8), 1.2))
                k = new[i, j].item() - refl[i, j].item()
                tot += l + 0.05 * (math.exp(-k) + k - 1)
                n += 1
    assert abs(loss.item() - tot / n) < 1e-9
    assert torch.autograd.gradcheck(lambda x: ref.policy_loss_token(x, old, adv, mask, 0.2, refl, 0.05)[0], (new,), atol=1e-6)


def test_sequence_loss_nll_value_loss():
    torch.manual_seed(0)
    new = torch.randn(4, 5, dtype=torch.float64, requires_grad=True)
    old = new.detach() + 0.05 * torch.randn(4, 5, dtype=torch.float64)
    adv = torch.randn(4, dtype=torch.float64)
    loss, st = ref.policy_loss_sequence(new, old, adv, 0.2)
    ratio = torch.exp(new.sum(1) - old.sum(1))
    want = torch.max(-adv * ratio, -adv * ratio.clamp(0.8, 1.2)).mean()
    assert torch.allclose(loss, want)
    assert torch.allclose(ref.nll_loss(new), -new.sum(1).mean())
    v = torch.randn(4, 5, dtype=torch.float64, requires_grad=True)
    vo = v.detach() + 0.3 * torch.randn(4, 5, dtype=torch.float64)
    R = torch.randn(4, 5, dtype=torch.float64)
    m = torch.rand(4, 5) > 0.2
    l, cf = ref.value_loss(v, vo, R, m, 0.2)
    vc = torch.max(torch.min(v, vo + 0.2), vo - 0.2)
    want = 0.5 * (torch.max((v - R) ** 2, (vc - R) ** 2) * m).sum() / m.sum()
    assert torch.allclose(l, want)
    assert torch.autograd.gradcheck(lambda x: ref.value_loss(x, vo, R, m, 0.2)[0], (v,), atol=1e-6)


def test_lmhead_logprob_matches_materialised_and_grads():
    torch.manual_seed(0)
    T, V, d = 37, 101, 16
    h = torch.randn(T, d, dtype=torch.float64, requires_grad=True)
    w = torch.randn(V, d, dtype=torch.float64, requires_grad=True)
    tgt = torch.randint(0, V, (T,))
    from nanorlhf_b200 import ops
    logp, ent = ops.lmhead_logprob(h, w, tgt, 0.7, True)
    z = (h @ w.t()) / 0.7
    want = torch.log_softmax(z, -1).gather(1, tgt[:, None]).squeeze(1)
    p = torch.softmax(z, -1)
    want_ent = torch.logsumexp(z, -1) - (p * z).sum(-1)
    assert torch.allclose(logp.double(), want, atol=1e-4) and torch.allclose(ent.double(), want_ent, atol=1e-3)
    g = torch.randn(T, dtype=torch.float64)
    (logp.double() * g).sum().backward()
    gh, gw = h.grad.clone(), w.grad.clone()
    h.grad = w.grad = None
    (want * g).sum().backward()
    assert torch.allclose(gh, h.grad, atol=1e-4) and torch.allclose(gw, w.grad, atol=1e-4)


def test_grpo_group_normalise_and_rloo_baseline():
    s = torch.tensor([1.0, 2.0, 3.0, 6.0, 5.0, 5.0, 5.0, 5.0])
    g = s.view(-1, 4)
    z = (g - g.mean(1, keepdim=True)) / g.std(1, keepdim=True)
    z = torch.where(torch.isnan(z), torch.zeros_like(z), z)
    assert torch.allclose(z[1], torch.zeros(4))
    assert abs(z[0].mean().item()) < 1e-6
    base = (g.sum(1, keepdim=True) - g) / 3
    assert torch.allclose((g - base)[0], torch.tensor([1 - 11 / 3, 2 - 10 / 3, 3 - 3.0, 6 - 2.0]))


def test_top_p_sampling_reference():
    torch.manual_seed(0)
    logits = torch.tensor([[2.0, 1.0, 0.0, -5.0]]).repeat(4000, 1)
    g = torch.Generator().manual_seed(1)
    tok = ref.top_p_sample(logits, 1.0, 0.8, g)
    freq = torch.bincount(tok, minlength=4).float() / 4000
    p = torch.softmax(logits[0], -1)
    keep = p[:2] / p[:2].sum()          # 0.665 + 0.245 >= 0.8 -> two tokens kept
    assert freq[2] == 0 and freq[3] == 0 and (freq[:2] - keep).abs().max() < 0.03
    assert ref.top_p_sample(logits[:3], 0.0, 1.0).tolist() == [0, 0, 0]
