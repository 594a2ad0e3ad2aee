"""PPO entry script with value-model initialisation (reference: /root/reference/PPO/ppo.py:78-399).

Run:  python PPO/ppo.py [--key=value ...]      |      torchrun --nproc-per-node 8 PPO/ppo.py

``Value_Finetune_Config`` / ``finetune_args`` configure the critic pre-fit (``finetuned_value_model``,
nanorlhf_b200/trainer/value_initializer.py); ``MyPPOConfig`` / ``training_args`` configure PPO itself with
separate policy / value learning rates and separate LoRA knobs for the critic.
"""
import os
import sys
from dataclasses import dataclass, field
from typing import List, Optional

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nanorlhf_b200 import entry
from nanorlhf_b200.config import DEFAULT_LORA_TARGETS, RLConfig, ValueFinetuneConfig
from nanorlhf_b200.trainer import PPOTrainer
from nanorlhf_b200.trainer.value_initializer import finetuned_value_model
from nanorlhf_b200.utils.callbacks import EarlyStoppingCallback

os.environ.setdefault("WANDB_PROJECT", "rlhf")
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
base_model = "Qwen/Qwen2.5-1.5B-Instruct"


@dataclass
class Value_Finetune_Config(ValueFinetuneConfig):
    pass


finetune_args = Value_Finetune_Config(
    train_data_size=500, train_split_rate=0.8, num_train_epochs=8,
    per_device_train_batch_size=32, per_device_eval_batch_size=50, gradient_accumulation_steps=12,
    learning_rate=1e-3, lr_scheduler_type="reduce_lr_on_plateau",
    lr_scheduler_kwargs={"mode": "min", "factor": 0.5, "patience": 0},
    early_stopping_patience=3, report_to="none",
)


@dataclass
class MyPPOConfig(RLConfig):
    policy_learning_rate: float = 6e-6
    value_learning_rate: float = 9e-6
    value_use_lora: bool = True
    value_lora_r: int = 64
    value_lora_alpha: int = 16
    value_lora_dropout: float = 0.0
    value_lora_bias: str = "none"
    value_lora_target_modules: List[str] = field(default_factory=lambda: list(DEFAULT_LORA_TARGETS))
    value_modules_to_save: Optional[List[str]] = field(default_factory=lambda: ["embed_tokens", "lm_head", "score", "wte", "wpe"])
    advantage_whiten: bool = False


training_args = MyPPOConfig(
    exp_name="ppo-v1",
    sft_model_path=base_model,
    reward_model_path="OpenAssistant/reward-model-deberta-v3-large-v2",
    output_dir=f"{base_model}/{os.environ['WANDB_PROJECT']}",
    kl_coef=0.01, cliprange=0.2, vf_coef=1, cliprange_value=0.2, gamma=1.0, lam=0.95,
    temperature=0.9, response_length=1500, whiten_rewards=False,
    per_device_train_batch_size=4, gradient_accumulation_steps=8, num_mini_batches=16, num_ppo_epochs=1,
    total_episodes=250000,
    learning_rate=6e-6, warmup_steps=0, lr_scheduler_type="cosine_with_min_lr", lr_scheduler_kwargs={"min_lr_rate": 0.1},
    bf16=True, gradient_checkpointing=True,
    use_lora=True, lora_r=64, lora_alpha=16, lora_dropout=0.0, modules_to_save=["embed_tokens", "lm_head", "score"],
    report_to="none", save_steps=1, save_total_limit=8, logging_steps=1, eval_steps=1,
    metric_for_best_model="eval_objective/rlhf_reward_old", greater_is_better=True, load_best_model_at_end=True,
    stop_token="eos", reward_batch_size=16, save_value_model=True,
    train_dataset_name="Anthropic/hh-rlhf", train_dataset_split="train[:100%]",
)

_reward = None


def reward_func(pmt_and_responses, eos_token):
    global _reward
    if _reward is None:
        _reward = entry.load_reward_func(training_args)
    return _reward.score_strings(pmt_and_responses, eos_token)


if __name__ == "__main__":
    training_args.apply_overrides()
    entry.prepare_output_dir(training_args)
    tokenizer, policy, ref_policy = entry.load_tokenizer_and_policies(training_args)
    value_model = entry.load_value_model(
        training_args, ref_policy, training_args.value_use_lora,
        dict(r=training_args.value_lora_r, lora_alpha=training_args.value_lora_alpha,
             target_modules=training_args.value_lora_target_modules, lora_dropout=training_args.value_lora_dropout,
             bias=training_args.value_lora_bias, task_type="CAUSAL_LM", modules_to_save=training_args.value_modules_to_save))
    train_dataset = entry.load_prompt_dataset(training_args, tokenizer)
    rf = entry.load_reward_func(training_args)
    rf = rf if rf.accepts_ids else reward_func

    # fit the critic to the initial policy's Monte-Carlo returns before PPO starts (reference: ppo.py:371-380)
    value_model = finetuned_value_model(value_model, policy, ref_policy, rf, train_dataset, tokenizer,
                                        training_args, finetune_args)

    trainer = PPOTrainer(
        config=training_args, processing_class=tokenizer, policy=policy, ref_policy=ref_policy,
        train_dataset=train_dataset, value_model=value_model, reward_func=rf,
        callbacks=[EarlyStoppingCallback(early_stopping_patience=training_args.early_stopping_patience)],
    )
    trainer.train()
