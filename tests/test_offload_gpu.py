"""K-OFF tiering engine on a GPU: host-resident roles round-trip through pinned memory on the side stream."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tiering_roundtrip_and_training_equivalence(tmp_path):
    from nanorlhf_b200.runtime.offload import TieringEngine
    dev = torch.device("cuda")
    m = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.Linear(512, 64)).to(dev, torch.bfloat16)
    x = torch.randn(8, 256, device=dev, dtype=torch.bfloat16)
    want = m(x)
    eng = TieringEngine(dev)
    eng.register("ref", m, "host")
    before = torch.cuda.memory_allocated()
    eng.evict("ref")
    assert all(p.numel() == 0 for p in m.parameters())
    assert torch.cuda.memory_allocated() < before
    eng.fetch("ref")
    assert torch.equal(m(x), want)
    eng.evict("ref")
    eng.prefetch("ref")            # async H2D on the side stream
    eng.fetch("ref")
    assert torch.equal(m(x), want)
    assert eng.stats()["offload/h2d_gb"] > 0


def test_trainer_with_host_offloaded_roles_matches_resident(tmp_path):
    from nanorlhf_b200.config import RLConfig
    from nanorlhf_b200.models.lora import LoraConfig, get_peft_model
    from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM
    from nanorlhf_b200.reward.api import TokenIdReward
    from nanorlhf_b200.sampler import engine
    from nanorlhf_b200.trainer import ReinforceTrainer
    from nanorlhf_b200.utils.data import synthetic_token_dataset
    from nanorlhf_b200.utils.tokenizer import ByteTokenizer
    dev = torch.device("cuda")
    cfg = Qwen2Config(vocab_size=2048, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=1, head_dim=128)
    tok = ByteTokenizer(vocab_size=2048)
    tok.special_tokens["[PAD]"], tok.special_tokens["<|im_end|>"] = 2047, 2046
    tok.pad_token_id, tok.eos_token_id, tok.vocab_size = 2047, 2046, 2048
    tok.id_to_special = {v: k for k, v in tok.special_tokens.items()}
    outs = []
    for policy_kind in ("resident", "host"):
        engine.reseed_stream(42)
        policy = get_peft_model(Qwen2ForCausalLM.from_config(cfg, torch.bfloat16, dev, seed=1), LoraConfig(r=8, lora_alpha=16, modules_to_save=None))
        with torch.no_grad():
            g = torch.Generator(device=dev).manual_seed(3)
            for p in policy.parameters():
                if p.requires_grad:
                    p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * 0.02)
        ref = Qwen2ForCausalLM.from_config(cfg, torch.bfloat16, dev, seed=1)
        a = RLConfig(output_dir=str(tmp_path / policy_kind), response_length=12, per_device_train_batch_size=2, gradient_accumulation_steps=1,
                     num_mini_batches=2, total_episodes=8, learning_rate=1e-3, report_to="none", save_strategy="no",
                     offload_policy=policy_kind, sampler="native", kl_coef=0.05)
        a.quiet = True
        t = ReinforceTrainer(a, tok, policy, ref, synthetic_token_dataset(16, 2000, 8, 20, seed=0), reward_func=TokenIdReward(7))
        m = t.train()
        outs.append((m["eval_objective/scores_old"], m["objective/kl_old"],
                     torch.cat([p.detach().float().reshape(-1) for p in t.policy.parameters() if p.requires_grad])))
    assert abs(outs[0][0] - outs[1][0]) < 1e-6 and abs(outs[0][1] - outs[1][1]) < 1e-4
    assert torch.allclose(outs[0][2], outs[1][2], atol=1e-5)
