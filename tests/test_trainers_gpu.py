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
    engine._ENGINES.clear()
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
