"""Native sampler vs the PyTorch reference sampler (token-exact at temperature 0)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _models():
    from nanorlhf_b200.models.lora import LoraConfig, get_peft_model
    from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM
    cfg = Qwen2Config(vocab_size=2048, hidden_size=256, intermediate_size=512, num_hidden_layers=3,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=128, tie_word_embeddings=True)
    m = Qwen2ForCausalLM.from_config(cfg, torch.bfloat16, "cuda", seed=3)
    m = get_peft_model(m, LoraConfig(r=8, lora_alpha=16, modules_to_save=None))
    with torch.no_grad():
        for mod in m.modules():
            if hasattr(mod, "lora_B"):
                mod.lora_B.weight.normal_(0, 0.02)
    return m


def test_greedy_matches_torch_sampler():
    from nanorlhf_b200.sampler.native_sampler import NativeSampler
    from nanorlhf_b200.sampler.torch_sampler import torch_generate
    m = _models()
    g = torch.Generator().manual_seed(0)
    prompts = [torch.randint(0, 2000, (int(L),), generator=g).tolist() for L in (5, 16, 17, 33, 64, 9, 31, 48)]
    eng = NativeSampler(m, kv_cache_gb=1.0, sync_every=8)
    eng.sync_weights()
    out = eng.generate(prompts, 2, 0.0, 1.0, 24, None, 2047, 1)
    want = torch_generate(m, prompts, 2, 0.0, 1.0, 24, None, 2047, 1)
    agree = (out == want).float().mean().item()
    # bf16 kernels differ in rounding from the fp32-softmax reference: demand near-exact agreement
    first_tok = (out[:, 0] == want[:, 0]).float().mean().item()
    assert first_tok == 1.0, (out[:, :4], want[:, :4])
    assert agree > 0.9, agree


def test_sampling_is_seeded_and_eos_stops():
    from nanorlhf_b200.sampler.native_sampler import NativeSampler
    m = _models()
    prompts = [[1, 2, 3, 4, 5, 6, 7], [8, 9, 10]]
    eng = NativeSampler(m, kv_cache_gb=1.0, sync_every=8)
    eng.sync_weights()
    a = eng.generate(prompts, 4, 0.9, 0.95, 20, None, 2047, 11)
    b = eng.generate(prompts, 4, 0.9, 0.95, 20, None, 2047, 11)
    c = eng.generate(prompts, 4, 0.9, 0.95, 20, None, 2047, 12)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert a.shape == (8, 20) and (a != 2047).all()
    # pick the most frequent first token as "eos": rows that emit it must be padded afterwards
    eos = int(a[:, 3].mode().values)
    d = eng.generate(prompts, 4, 0.9, 0.95, 20, eos, 2047, 11)
    for row in d.tolist():
        if eos in row:
            i = row.index(eos)
            assert all(t == 2047 for t in row[i + 1:])


def test_waves_under_kv_pressure_match_unconstrained():
    """Continuous batching: with a KV pool that holds only a fraction of the requests the scheduler admits them in
    waves (pages of finished sequences are recycled); greedy outputs must equal the all-at-once run."""
    from nanorlhf_b200.sampler.native_sampler import NativeSampler
    m = _models()
    g = torch.Generator().manual_seed(1)
    prompts = [torch.randint(0, 2000, (int(L),), generator=g).tolist() for L in torch.randint(4, 60, (12,), generator=g)]
    big = NativeSampler(m, kv_cache_gb=1.0, sync_every=8)
    big.sync_weights()
    ref_out = big.generate(prompts, 2, 0.0, 1.0, 40, None, 2047, 1)
    # an "eos" that actually occurs, so sequences finish at different steps
    eos = int(ref_out[:, 5:30].flatten().mode().values)
    want = big.generate(prompts, 2, 0.0, 1.0, 40, eos, 2047, 1)
    small = NativeSampler(m, kv_cache_gb=0.0032, sync_every=4)      # ~65 pages: a third of what all 24 sequences need
    small.sync_weights()
    got = small.generate(prompts, 2, 0.0, 1.0, 40, eos, 2047, 1)
    assert small.num_blocks < 100
    assert torch.equal(got, want)
    capped = NativeSampler(m, kv_cache_gb=1.0, sync_every=4, max_num_seqs=6)
    capped.sync_weights()
    assert torch.equal(capped.generate(prompts, 2, 0.0, 1.0, 40, eos, 2047, 1), want)


def test_compaction_on_eos_heavy_batch():
    """Finished rows leave the batch at sync points: with a 64-token vocabulary a sequence meets EOS
    after ~64 tokens on average, far before max_tokens.  Compaction must (a) happen, (b) cut the row-steps executed, and
    (c) leave every sampled token unchanged -- the per-row RNG stream is (seed, sequence id, tokens generated), so the
    batch composition must not matter."""
    from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM
    from nanorlhf_b200.sampler.native_sampler import NativeSampler
    cfg = Qwen2Config(vocab_size=64, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                      num_attention_heads=2, num_key_value_heads=1, head_dim=128, tie_word_embeddings=True)
    m = Qwen2ForCausalLM.from_config(cfg, torch.bfloat16, "cuda", seed=5)
    g = torch.Generator().manual_seed(1)
    prompts = [torch.randint(0, 60, (int(L),), generator=g).tolist() for L in torch.randint(4, 40, (96,), generator=g)]
    eos, pad, max_tokens = 62, 63, 320
    outs, stats = {}, {}
    for compact in (False, True):
        eng = NativeSampler(m, kv_cache_gb=1.0, sync_every=8)
        eng.enable_compaction = compact
        eng.sync_weights()
        outs[compact] = eng.generate(prompts, 4, 1.0, 1.0, max_tokens, eos, pad, 7)
        stats[compact] = dict(eng.stats)
    a, b = outs[False], outs[True]
    # The RNG stream of a row does not depend on the batch it sits in, so compaction may only change a row through kernel
    # numerics (the split-KV factor of the decode attention follows the batch size): near-ties can flip a token, after
    # which that row diverges.  Mis-routed rows would show up as (almost) no agreement.
    # (measured: bf16 logits carry ~0.4 % relative error, a row of ~64 sampled tokens meets a near-tie with p ~ 0.2)
    same_rows = (a == b).all(1).float().mean().item()
    assert same_rows > 0.6, f"only {same_rows:.2f} of the rows survive compaction unchanged"
    assert torch.equal(a[:, 0], b[:, 0])                             # first tokens come from the shared prefill
    assert ((a[:, :4] == b[:, :4]).all(1).float().mean().item()) > 0.95
    has_eos = (b == eos).any(1)
    assert has_eos.float().mean() > 0.9                              # EOS-heavy as intended
    first = torch.where(has_eos, (b == eos).int().argmax(1), torch.full_like(b[:, 0], max_tokens))
    col = torch.arange(max_tokens, device=b.device)[None]
    assert ((b == pad) | (col <= first[:, None])).all()              # nothing but padding after EOS
    assert stats[True]["compactions"] > 0 and stats[False]["compactions"] == 0
    assert stats[True]["row_steps"] < 0.7 * stats[False]["row_steps"], (stats[True]["row_steps"], stats[False]["row_steps"])
