"""Flat AdamW: fp32 master weights at the shipped learning rate, step-count invariance of ``train_samples_per_prompt``,
resharding of the saved optimizer state (reference optimizer: GRPO/grpo_trainer.py:258,692)."""
import torch

from nanorlhf_b200.parallel.optimizer import FusedAdamW, shard_bounds


def _run(master: bool, steps: int = 100, lr: float = 6e-6):
    torch.manual_seed(0)
    w0 = (torch.rand(4096) * 0.05 - 0.025)
    p = torch.nn.Parameter(w0.to(torch.bfloat16).clone())
    opt = FusedAdamW([{"params": [p]}], lr=lr, master_weights=master)
    ref = torch.nn.Parameter(w0.to(torch.bfloat16).float().clone())
    ropt = torch.optim.AdamW([ref], lr=lr, weight_decay=0.0)
    g = torch.Generator().manual_seed(1)
    for _ in range(steps):
        grad = torch.randn(4096, generator=g) + 0.5          # a consistent sign component, like a real descent direction
        opt.zero_grad()
        p.grad.copy_(grad.to(torch.bfloat16))
        ref.grad = grad.to(torch.bfloat16).float()
        opt.step()
        ropt.step()
    return w0.to(torch.bfloat16).float(), p.detach().float(), ref.detach(), opt


def test_master_weights_move_like_fp32_adamw_at_lr_6e6():
    w0, p, ref, opt = _run(master=True)
    f = opt.flats[0]
    want = (ref - w0).abs().mean()
    got = (f.master[:4096] - w0).abs().mean()
    assert want > 1e-4                                          # 100 steps x ~lr
    assert abs(got - want) / want < 0.02                        # the fp32 truth tracks torch.optim.AdamW
    assert torch.allclose(p, f.master[:4096].to(torch.bfloat16).float())      # bf16 params = rounded master


def test_without_master_bf16_rounding_stalls():
    """Documents the failure the master copy prevents: |w| ~ 0.02 has a bf16 ulp of 1.2e-4 >> lr."""
    w0, p, ref, _ = _run(master=False)
    assert (p - w0).abs().mean() < 0.25 * (ref - w0).abs().mean()


def test_shard_bounds_any_world():
    for padded in (1024, 8192, 540672000 // 1024 * 1024):
        for world in (1, 2, 3, 5, 6, 7, 8):
            spans = [shard_bounds(padded, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == padded
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo % 8 == 0 and (hi - lo) % 8 == 0 for lo, hi in spans)


def test_full_state_reshards():
    """A whole-buffer optimizer.pt loads into any shard layout (the owner takes its slice)."""
    _, _, _, opt = _run(master=True, steps=3)
    sd = opt.state_dict(full=True)
    q = torch.nn.Parameter(torch.zeros(4096, dtype=torch.bfloat16))
    opt2 = FusedAdamW([{"params": [q]}], lr=6e-6)
    opt2.load_state_dict(sd)
    f, f2 = opt.flats[0], opt2.flats[0]
    assert torch.equal(f.exp_avg, f2.exp_avg) and torch.equal(f.master, f2.master)
    assert torch.equal(q.detach(), f.param[:4096])              # parameters restored from the master copy
    # simulate rank 1 of 2 taking its slice of the same file
    lo, hi = shard_bounds(f.padded, 2, 1)
    opt2._shard_bounds = lambda _f: (lo, hi)
    f2.exp_avg = f2.exp_avg_sq = f2.master = None
    opt2.load_state_dict(sd)
    assert torch.equal(f2.exp_avg, f.exp_avg[lo:hi]) and torch.equal(f2.master, f.master[lo:hi])
