"""Every algorithm end to end on the native (sm_100a) backend: sampler, fused log-prob, loss/GAE kernels, flat AdamW."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(tmp_path, cls, extra=None, **kw):
    from nanorlhf_b200.config import RLConfig
    from nanorlhf_b200.models.lora import LoraConfig, get_peft_model
    from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM, Qwen2ForSequenceClassification
    from nanorlhf_b200.reward.api import TokenIdReward
    from nanorlhf_b200.sampler import engine
    from nanorlhf_b200.trainer import PPOTrainer
    from nanorlhf_b200.utils.data import synthetic_token_dataset
    from nanorlhf_b200.utils.tokenizer import ByteTokenizer
    dev = torch.device("cuda")
    cfg = Qwen2Config(vocab_size=4096, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=128)
    tok = ByteTokenizer(vocab_size=4096)
    tok.special_tokens["[PAD]"], tok.special_tokens["<|im_end|>"] = 4095, 4094
    tok.pad_token_id, tok.eos_token_id, tok.vocab_size = 4095, 4094, 4096
    tok.id_to_special = {v: k for k, v in tok.special_tokens.items()}
    policy = get_peft_model(Qwen2ForCausalLM.from_config(cfg, torch.bfloat16, dev, seed=1),
                            LoraConfig(r=8, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head"]))
    ref = Qwen2ForCausalLM.from_config(cfg, torch.bfloat16, dev, seed=1)
    base = dict(output_dir=str(tmp_path), response_length=24, per_device_train_batch_size=2, gradient_accumulation_steps=2,
                num_mini_batches=2, total_episodes=16, learning_rate=1e-4, report_to="none", sampler="native", kl_coef=0.05)
    base.update(kw)
    a = RLConfig(**base)
    a.quiet = True
    for k, v in (extra or {}).items():
        setattr(a, k, v)
    kwargs = {}
    if cls is PPOTrainer:
        vm = Qwen2ForSequenceClassification.from_causal_lm(ref)
        kwargs["value_model"] = get_peft_model(vm, LoraConfig(r=8, lora_alpha=16, modules_to_save=["score"]))
    return cls(a, tok, policy, ref, synthetic_token_dataset(64, 4000, 8, 40, seed=0), reward_func=TokenIdReward(7), **kwargs)


@pytest.mark.parametrize("name,extra,kw", [
    ("ReinforceTrainer", None, dict(advantage_whiten=True)), ("GRPOTrainer", {"grpo_sample_N": 4}, {}),
    ("RLOOTrainer", {"rloo_sample_N": 4}, {}), ("RemaxTrainer", None, {}), ("RAFTTrainer", {"raft_sample_K": 4}, {}),
    ("PPOTrainer", {"policy_learning_rate": 1e-4, "value_learning_rate": 2e-4}, dict(vf_coef=1.0)),
    ("SparseGRPOTrainer", {"grpo_sample_N": 4}, {}),
])
def test_algorithm_on_native_backend(name, extra, kw, tmp_path):
    import nanorlhf_b200.trainer as T
    from nanorlhf_b200.ops import native
    t = _setup(tmp_path, getattr(T, name), extra, **kw)
    n0 = native.launches()
    m = t.train()
    assert t.state.global_step == 2
    assert native.launches() - n0 > 50, "native kernels did not run"
    for k, v in m.items():
        if isinstance(v, float):
            assert v == v and abs(v) < 1e9, (k, v)
    assert all(torch.isfinite(p).all() for p in t.policy.parameters())
    assert os.path.isdir(os.path.join(str(tmp_path), "checkpoint-2"))


@pytest.mark.parametrize("algo", ["grpo", "ppo"])
def test_graphed_micro_step_matches_eager(tmp_path, algo):
    """CUDA-graph replay of fwd+loss+bwd (trainer/graphed.py) reproduces the eager micro-step: stats and gradients
    (PPO: the critic's forward / backward is part of the same graph)."""
    from nanorlhf_b200.trainer import GRPOTrainer, PPOTrainer
    from nanorlhf_b200.trainer.graphed import GraphedMicroStep
    if algo == "ppo":
        t = _setup(tmp_path, PPOTrainer, {"policy_learning_rate": 1e-4, "value_learning_rate": 2e-4}, vf_coef=1.0,
                   gradient_checkpointing=False, train_cuda_graph="on")
    else:
        t = _setup(tmp_path, GRPOTrainer, {"grpo_sample_N": 4}, gradient_checkpointing=False, train_cuda_graph="on")
    assert t._graph_micro_step() is not None
    dev, pad = t.device, t.tokenizer.pad_token_id
    B, ctx, T_r = 2, 12, 24
    gen = torch.Generator(device="cpu").manual_seed(0)

    def make_mb(qlens, rlens):
        qr = torch.full((B, ctx + T_r), pad, dtype=torch.long)
        pm = torch.ones(B, T_r, dtype=torch.bool)
        for b, (ql, rl) in enumerate(zip(qlens, rlens)):
            qr[b, ctx - ql:ctx + rl] = torch.randint(0, 4000, (ql + rl,), generator=gen)
            pm[b, :rl] = False
        f = lambda: torch.randn(B, T_r, generator=gen) * 0.1 - 2.0                                   # noqa: E731
        mb = {"query_responses": qr.to(dev), "padding_mask": pm.to(dev), "logprobs": f().to(dev),
              "ref_logprobs": f().to(dev), "advantages": torch.randn(B, T_r, generator=gen).to(dev),
              "context_length": ctx}
        if algo == "ppo":
            pm1 = pm.clone()
            for b, rl in enumerate(rlens):
                pm1[b, :min(rl + 1, T_r)] = False
            mb.update(padding_mask_p1=pm1.to(dev), values=torch.randn(B, T_r, generator=gen).to(dev),
                      returns=torch.randn(B, T_r, generator=gen).to(dev))
        return mb

    params = [p for p in t.model.parameters() if p.requires_grad]

    def run(step, mb):
        t.optimizer.zero_grad()
        vec = step(mb, ctx, pad).clone()
        return vec, torch.cat([p.grad.detach().float().reshape(-1) for p in params]).clone()

    t.model.train()
    g = GraphedMicroStep(t)
    mb1, mb2 = make_mb([5, 9], [24, 17]), make_mb([7, 8], [20, 20])          # same bucket, different data
    v_eager, g_eager = run(g, mb1)                                           # first sight of the bucket: eager
    v_graph, g_graph = run(g, mb1)                                           # capture + replay
    assert g.replays == 1 and g.eager == 1
    assert torch.allclose(v_eager, v_graph, rtol=1e-4, atol=1e-5), (v_eager, v_graph)
    assert (g_eager - g_graph).abs().max() <= 1e-3 * g_eager.abs().max() + 1e-7
    v2_graph, g2_graph = run(g, mb2)                                         # pure replay on new data
    v2_eager, g2_eager = run(GraphedMicroStep(t), mb2)                       # a fresh instance runs eagerly
    assert g.replays == 2
    assert torch.allclose(v2_eager, v2_graph, rtol=1e-4, atol=1e-5)
    assert (g2_eager - g2_graph).abs().max() <= 1e-3 * g2_eager.abs().max() + 1e-7
    assert g2_eager.abs().max() > 0


def test_grpo_trains_with_cuda_graph_micro_steps(tmp_path):
    from nanorlhf_b200.trainer import GRPOTrainer
    t = _setup(tmp_path, GRPOTrainer, {"grpo_sample_N": 4}, gradient_checkpointing=False, train_cuda_graph="on",
               total_episodes=32)
    m = t.train()
    assert t._graphed is not None and t._graphed.replays > 0
    for k, v in m.items():
        if isinstance(v, float):
            assert v == v and abs(v) < 1e9, (k, v)
    assert all(torch.isfinite(p).all() for p in t.policy.parameters())


def test_async_checkpoint_roundtrip(tmp_path):
    """save_steps=1 with the asynchronous writer (pinned-host snapshot on a side stream + background serialisation): every
    checkpoint directory is complete, and the adapter / optimizer state on disk is the state AT THAT UPDATE (the snapshot
    must not race with the next update's optimizer steps)."""
    import nanorlhf_b200.trainer as T
    from nanorlhf_b200.models.hf_io import load_state_dict
    t = _setup(tmp_path, T.GRPOTrainer, {"grpo_sample_N": 4}, save_strategy="steps", save_steps=1, total_episodes=24)
    assert t.args.async_checkpoint
    snaps = {}
    orig = t._save_checkpoint

    def spy(model, trial=None, metrics=None):
        snaps[t.state.global_step] = {k: v.detach().clone() for k, v in t.policy.adapter_state_dict().items()}
        snaps[(t.state.global_step, "m")] = t.optimizer.flats[0].exp_avg.clone()
        return orig(model, trial, metrics)

    t._save_checkpoint = spy
    t.train()
    for step in (1, 2, 3):
        d = os.path.join(str(tmp_path), f"checkpoint-{step}")
        for f in ("adapter_model.safetensors", "adapter_config.json", "optimizer.pt", "scheduler.pt", "rng_state.pth", "trainer_state.json"):
            assert os.path.exists(os.path.join(d, f)), (step, f)
        sd = load_state_dict(os.path.join(d, "adapter_model.safetensors"))
        for k, v in snaps[step].items():
            assert torch.equal(sd[k].to(v.device), v), (step, k)
        opt = torch.load(os.path.join(d, "optimizer.pt"), map_location="cpu", weights_only=False)
        assert torch.equal(opt["state"]["group0.exp_avg"], snaps[(step, "m")].cpu())
