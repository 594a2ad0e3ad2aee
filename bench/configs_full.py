"""Full-size runs of the other BASELINE.json configurations (3, 4, 5) through the public trainer API.

  python bench/configs_full.py [ppo] [r1] [rloo] [remax] [raft]           (one GPU)
  torchrun --nproc-per-node N --master-addr 127.0.0.1 bench/configs_full.py ...   (data parallel)

  config 3  ppo    Qwen2.5-1.5B PPO, 1500-token responses, DeBERTa-v3-large reward, value-model initialisation with the
                   reference's Value_Finetune_Config (500 prompts, 8 epochs, bs 32 x accum 12, lr 1e-3, plateau LR, early stop)
  config 4  r1     Qwen2.5-7B sparse GRPO, 8000-token responses, rule-style 0/1 reward, dynamic micro-buckets,
                   gradient checkpointing, fp8 KV pages (+ fp8 rollout GEMMs with R1_ROLLOUT=fp8)
  config 5  rloo / remax / raft   Qwen2.5-1.5B, ref + reward model + optimizer state tiered to pinned host memory

Random-init weights of the named architectures, synthetic prompts (no network).  Prompts per rank default to 256
(PROMPTS env) -- the headline benchmark's batch -- and UPDATES (default 2) timed updates follow one warm-up update.
Writes gpurun_out/configs_full_<world>gpu.json: episodes/s (device-timed, max over ranks), phase split, peak memory,
value-init wall time."""
import json
import os
import sys
import time
from dataclasses import dataclass

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanorlhf_b200.config import RLConfig, ValueFinetuneConfig  # noqa: E402
from nanorlhf_b200.models.deberta_v3 import DebertaV3Config, DebertaV3ForSequenceClassification  # noqa: E402
from nanorlhf_b200.models.lora import LoraConfig, get_peft_model  # noqa: E402
from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM, Qwen2ForSequenceClassification  # noqa: E402
from nanorlhf_b200.parallel.comm import Comm  # noqa: E402
from nanorlhf_b200.reward.model_reward import ModelReward  # noqa: E402
from nanorlhf_b200.trainer import PPOTrainer, RAFTTrainer, RemaxTrainer, RLOOTrainer, SparseGRPOTrainer  # noqa: E402
from nanorlhf_b200.trainer.value_initializer import finetuned_value_model  # noqa: E402
from nanorlhf_b200.utils.clocks import ClockSampler  # noqa: E402
from nanorlhf_b200.utils.data import synthetic_token_dataset  # noqa: E402
from nanorlhf_b200.utils.tokenizer import ByteTokenizer  # noqa: E402

comm = Comm.from_env()
dev = comm.device
which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["ppo", "r1", "rloo", "remax", "raft"]
PROMPTS = int(os.environ.get("PROMPTS", "256"))
UPDATES = int(os.environ.get("UPDATES", "2"))
results = {"world": comm.world_size, "prompts_per_rank": PROMPTS, "timed_updates": UPDATES}


def tok_for(shape):
    tok = ByteTokenizer(vocab_size=shape.vocab_size - 1)
    tok.special_tokens["<|im_end|>"], tok.special_tokens["[PAD]"] = shape.vocab_size - 2, shape.vocab_size - 1
    tok.id_to_special = {v: k for k, v in tok.special_tokens.items()}
    tok.eos_token_id, tok.pad_token_id, tok.vocab_size = shape.vocab_size - 2, shape.vocab_size - 1, shape.vocab_size
    return tok


def base_cfg(name, prompts, **kw):
    mini = max(1, prompts // 32)
    d = dict(output_dir=f"/tmp/nrl_full_{name}_{os.getpid()}", response_length=1500, per_device_train_batch_size=4,
             gradient_accumulation_steps=8, num_mini_batches=mini, total_episodes=4 * 8 * mini * comm.world_size * (UPDATES + 1),
             report_to="none", save_strategy="no", resume="never", sampler="native", watchdog_timeout_s=0,
             gradient_checkpointing=False, learning_rate=6e-6)
    d.update(kw)
    c = RLConfig(**d)
    c.quiet = True
    return c


def deberta_reward():
    rm = DebertaV3ForSequenceClassification.from_config(DebertaV3Config.large(), torch.bfloat16, dev, seed=1)
    return ModelReward(rm, None, reward_batch_size=16, device=dev, token_budget=65536)


def timed_updates(name, trainer, extra=None):
    it = iter(trainer.dataloader)
    trainer.train_one_update(1, next(it))                    # warm-up (graph capture, lazy inits)
    comm.barrier()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    clk = ClockSampler(dev.index or 0, 500).start() if comm.is_main else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    phase, last = {}, {}
    for u in range(2, 2 + UPDATES):
        last = trainer.train_one_update(u, next(it))
        for k, v in last.items():
            if k.startswith("time/") and k.endswith("_s") and "wall" not in k:
                phase[k[5:-2]] = phase.get(k[5:-2], 0.0) + v / UPDATES
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 1e3, torch.cuda.max_memory_allocated() / 2**30], dtype=torch.float64, device=dev)
    comm.all_reduce_(t, "max")
    secs, peak = t.tolist()
    row = {"episodes_per_s": trainer.args.batch_size * UPDATES / secs, "s_per_update": secs / UPDATES, "global_batch": trainer.args.batch_size,
           "phases_s": phase, "peak_mem_gb": peak, "clocks": clk.stop() if clk else None,
           **{k: last[k] for k in last if k.startswith(("loss/", "objective/kl", "eval_objective/scores", "val/num_eos"))}, **(extra or {})}
    results[name] = row
    if comm.is_main:
        print(name, json.dumps(row), flush=True)
    trainer.heartbeat.close()


if "ppo" in which:
    shape = Qwen2Config.qwen2_5_1_5b()
    tok = tok_for(shape)
    policy = get_peft_model(Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0),
                            LoraConfig(r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head", "score"]))
    ref = Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0)
    vm = get_peft_model(Qwen2ForSequenceClassification.from_causal_lm(ref),
                        LoraConfig(r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head", "score", "wte", "wpe"]))
    cfg = base_cfg("ppo", PROMPTS, vf_coef=1.0, lam=0.95, cliprange_value=0.2)
    cfg.policy_learning_rate, cfg.value_learning_rate = 6e-6, 9e-6
    ds = synthetic_token_dataset(max(PROMPTS * comm.world_size * 2, 512), shape.vocab_size - 2, 24, 160, seed=1)
    rf = deberta_reward()
    torch.cuda.synchronize()
    t0 = time.time()
    vm = finetuned_value_model(vm, policy, ref, rf, ds, tok, cfg, ValueFinetuneConfig(), verbose=False)
    torch.cuda.synchronize()
    vinit = time.time() - t0
    timed_updates("ppo_1.5b", PPOTrainer(cfg, tok, policy, ref, ds, value_model=vm, reward_func=rf, comm=comm),
                  {"value_init_wall_s": vinit, "value_init_reference": "about 15 minutes on 1 x A100-40G (PPO/ppo.py:370)"})
    del policy, ref, vm, rf
    torch.cuda.empty_cache()

if "r1" in which:
    shape = Qwen2Config.qwen2_5_7b()
    tok = tok_for(shape)
    policy = get_peft_model(Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0),
                            LoraConfig(r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head", "score"]))
    ref = Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0)
    r1_prompts = int(os.environ.get("R1_PROMPTS", "64"))
    cfg = base_cfg("r1", r1_prompts, kl_coef=0.0, response_length=int(os.environ.get("R1_RESPONSE", "8000")),
                   rollout_dtype=os.environ.get("R1_ROLLOUT", "bf16"), kv_cache_dtype="fp8", gradient_checkpointing=True,
                   learning_rate=9e-6, token_budget_fwd=22 * 2316, token_budget_train=4 * 2316)
    cfg.grpo_sample_N = 4
    ds = synthetic_token_dataset(max(r1_prompts * comm.world_size * 2, 256), shape.vocab_size - 2, 24, 160, seed=1)

    class RandomBinaryReward:           # rule-style 0/1 reward: some groups get zero advantage and are dropped (sparse filter)
        accepts_ids = True

        def __call__(self, q, r, tokenizer):
            return ((r[:, :8].sum(1) % 3) == 0).float()

    timed_updates("sparse_grpo_7b_8000tok", SparseGRPOTrainer(cfg, tok, policy, ref, ds, reward_func=RandomBinaryReward(), comm=comm),
                  {"response_length": cfg.response_length, "rollout_dtype": cfg.rollout_dtype, "kv_cache_dtype": "fp8"})
    del policy, ref
    torch.cuda.empty_cache()

for name, cls, extra in (("rloo", RLOOTrainer, {"rloo_sample_N": 4}), ("remax", RemaxTrainer, {}), ("raft", RAFTTrainer, {"raft_sample_K": 4})):
    if name not in which:
        continue
    shape = Qwen2Config.qwen2_5_1_5b()
    tok = tok_for(shape)
    policy = get_peft_model(Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0),
                            LoraConfig(r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head", "score"]))
    ref = Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0)
    cfg = base_cfg(name, PROMPTS, offload_ref="host", offload_reward="host", offload_optimizer="host")
    for k, v in extra.items():
        setattr(cfg, k, v)
    ds = synthetic_token_dataset(max(PROMPTS * comm.world_size * 2, 512), shape.vocab_size - 2, 24, 160, seed=1)
    timed_updates(f"{name}_1.5b_host_offload", cls(cfg, tok, policy, ref, ds, reward_func=deberta_reward(), comm=comm),
                  {"offload": "ref + reward + optimizer state in pinned host memory between their phases"})
    del policy, ref
    torch.cuda.empty_cache()

if comm.is_main:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"configs_full_{comm.world_size}gpu.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(results)
    json.dump(old, open(path, "w"), indent=1)
comm.barrier()
comm.close()
