"""Reduced-size smoke runs of the other BASELINE.json configs on one B200 (writes gpurun_out/configs_smoke.json):
  config 3  Qwen2.5-1.5B PPO with value-model initialisation, bf16
  config 4  Qwen2.5-7B sparse-GRPO (r1-v0 style) with rule reward, dynamic mini-batching, fp8 rollout
  config 5  Qwen2.5-1.5B RLOO / ReMax / RAFT with host-offloaded ref + optimizer state
Sizes (prompts per update, response length) are cut so each finishes in about a minute; shapes are the real ones."""
import json
import os
import sys
import time
from dataclasses import dataclass

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanorlhf_b200.config import RLConfig, ValueFinetuneConfig  # noqa: E402
from nanorlhf_b200.models.lora import LoraConfig, get_peft_model  # noqa: E402
from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM, Qwen2ForSequenceClassification  # noqa: E402
from nanorlhf_b200.reward.api import TokenIdReward  # noqa: E402
from nanorlhf_b200.sampler import engine  # noqa: E402
from nanorlhf_b200.trainer import PPOTrainer, RAFTTrainer, RemaxTrainer, RLOOTrainer, SparseGRPOTrainer  # noqa: E402
from nanorlhf_b200.trainer.value_initializer import finetuned_value_model  # noqa: E402
from nanorlhf_b200.utils.data import synthetic_token_dataset  # noqa: E402
from nanorlhf_b200.utils.tokenizer import ByteTokenizer  # noqa: E402

dev = torch.device("cuda")
which = sys.argv[1:] or ["ppo", "r1", "sweep"]
results = {}


def tok_for(shape):
    tok = ByteTokenizer(vocab_size=shape.vocab_size - 1)
    tok.special_tokens["<|im_end|>"], tok.special_tokens["[PAD]"] = shape.vocab_size - 2, shape.vocab_size - 1
    tok.id_to_special = {v: k for k, v in tok.special_tokens.items()}
    tok.eos_token_id, tok.pad_token_id, tok.vocab_size = shape.vocab_size - 2, shape.vocab_size - 1, shape.vocab_size
    return tok


def base_cfg(name, **kw):
    d = dict(output_dir=f"/tmp/nrl_smoke_{name}", response_length=256, per_device_train_batch_size=4, gradient_accumulation_steps=4,
             num_mini_batches=2, total_episodes=64, report_to="none", save_strategy="no", resume="never", sampler="native",
             watchdog_timeout_s=0, gradient_checkpointing=False)
    d.update(kw)
    c = RLConfig(**d)
    c.quiet = True
    return c


def run(name, trainer):
    t0 = time.time()
    m = trainer.train()
    torch.cuda.synchronize()
    results[name] = {"seconds": time.time() - t0, "episodes_per_s": m["throughput/episodes_per_s"], "updates": trainer.state.global_step,
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                     **{k: m[k] for k in m if k.startswith(("loss/", "objective/kl", "eval_objective/scores"))}}
    print(name, json.dumps(results[name]), flush=True)


if "ppo" in which:
    shape = Qwen2Config.qwen2_5_1_5b()
    tok = tok_for(shape)
    policy = get_peft_model(Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0), LoraConfig(r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head"]))
    ref = Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0)
    vm = get_peft_model(Qwen2ForSequenceClassification.from_causal_lm(ref), LoraConfig(r=64, lora_alpha=16, modules_to_save=["score"]))
    cfg = base_cfg("ppo", vf_coef=1.0, lam=0.95)
    cfg.policy_learning_rate, cfg.value_learning_rate = 6e-6, 9e-6
    ds = synthetic_token_dataset(256, shape.vocab_size - 2, 24, 96, seed=1)
    rf = TokenIdReward(7)
    t0 = time.time()
    vm = finetuned_value_model(vm, policy, ref, rf, ds, tok, cfg, ValueFinetuneConfig(train_data_size=32, num_train_epochs=2,
                               per_device_train_batch_size=8, gradient_accumulation_steps=1, learning_rate=1e-3), verbose=False)
    results["ppo_value_init_s"] = time.time() - t0
    run("ppo_1.5b", PPOTrainer(cfg, tok, policy, ref, ds, value_model=vm, reward_func=rf))
    del policy, ref, vm
    torch.cuda.empty_cache()

if "r1" in which:
    shape = Qwen2Config.qwen2_5_7b()
    tok = tok_for(shape)
    policy = get_peft_model(Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0), LoraConfig(r=64, lora_alpha=16, modules_to_save=None))
    ref = Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0)
    cfg = base_cfg("r1", kl_coef=0.0, response_length=512, rollout_dtype="fp8", kv_cache_dtype="fp8", total_episodes=32)
    cfg.grpo_sample_N = 4
    ds = synthetic_token_dataset(128, shape.vocab_size - 2, 24, 96, seed=1)

    class RandomBinaryReward:           # rule-style 0/1 reward so that some groups have zero advantage (sparse filter)
        accepts_ids = True

        def __call__(self, q, r, tokenizer):
            return ((r[:, :8].sum(1) % 3) == 0).float()

    run("sparse_grpo_7b_fp8_rollout", SparseGRPOTrainer(cfg, tok, policy, ref, ds, reward_func=RandomBinaryReward()))
    del policy, ref
    torch.cuda.empty_cache()

if "sweep" in which:
    shape = Qwen2Config.qwen2_5_1_5b()
    tok = tok_for(shape)
    for name, cls, extra in (("rloo", RLOOTrainer, {"rloo_sample_N": 4}), ("remax", RemaxTrainer, {}), ("raft", RAFTTrainer, {"raft_sample_K": 4})):
        policy = get_peft_model(Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0), LoraConfig(r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head"]))
        ref = Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0)
        cfg = base_cfg(name, offload_ref="host", offload_optimizer="host", total_episodes=32)
        for k, v in extra.items():
            setattr(cfg, k, v)
        ds = synthetic_token_dataset(128, shape.vocab_size - 2, 24, 96, seed=1)
        run(f"{name}_1.5b_host_offload", cls(cfg, tok, policy, ref, ds, reward_func=TokenIdReward(7)))
        del policy, ref
        torch.cuda.empty_cache()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", "configs_smoke.json"), "w"), indent=1)
