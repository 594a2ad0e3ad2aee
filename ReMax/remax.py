"""Remax entry script -- "ALL setting is on the file you run" (reference README.md:34).

Run:  python ReMax/remax.py [--key=value ...]          (1 GPU or CPU)
      torchrun --nproc-per-node 8 ReMax/remax.py       (data parallel, one rank per B200)

Mirrors /root/reference/ReMax/remax.py:86-150: a ``RemaxConfig`` dataclass instantiated as ``training_args``,
a ``reward_func(pmt_and_responses, eos_token)`` callback and a ``__main__`` that builds tokenizer / policy /
ref policy / dataset and calls ``RemaxTrainer(...).train()``.  Models and data come from local directories when
present and fall back to synthetic stand-ins otherwise (nanorlhf_b200/entry.py).
"""
import os
import sys
from dataclasses import dataclass

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nanorlhf_b200 import entry
from nanorlhf_b200.config import RLConfig
from nanorlhf_b200.trainer import RemaxTrainer
from nanorlhf_b200.utils.callbacks import EarlyStoppingCallback

os.environ.setdefault("WANDB_PROJECT", "rlhf")
base_model = "Qwen/Qwen2.5-1.5B-Instruct"


@dataclass
class RemaxConfig(RLConfig):
    advantage_whiten: bool = False


training_args = RemaxConfig(
    exp_name="remax-v1",
    sft_model_path=base_model,
    reward_model_path="OpenAssistant/reward-model-deberta-v3-large-v2",
    output_dir=f"{base_model}/{os.environ['WANDB_PROJECT']}",
    # algorithm
    kl_coef=0.01,
    cliprange=0.2,
    temperature=0.9,
    response_length=1500,
    whiten_rewards=False,
    # batch arithmetic: 4 x 8 x 16 = 512 prompts per rank per update
    per_device_train_batch_size=4,
    gradient_accumulation_steps=8,
    num_mini_batches=16,
    num_ppo_epochs=1,
    total_episodes=250000,
    # optimisation
    learning_rate=6e-6,
    warmup_steps=0,
    lr_scheduler_type="cosine_with_min_lr",
    lr_scheduler_kwargs={"min_lr_rate": 0.1},
    bf16=True,
    gradient_checkpointing=True,
    # LoRA
    use_lora=True, lora_r=64, lora_alpha=16, lora_dropout=0.0,
    modules_to_save=["embed_tokens", "lm_head", "score"],
    # bookkeeping
    report_to="none",
    save_steps=1, save_total_limit=8, logging_steps=1, eval_steps=1,
    metric_for_best_model="eval_objective/rlhf_reward_old", greater_is_better=True, load_best_model_at_end=True,
    stop_token="eos",
    reward_batch_size=16,
    train_dataset_name="Anthropic/hh-rlhf", train_dataset_split="train[:100%]",
)

_reward = None


def reward_func(pmt_and_responses, eos_token):
    """Same contract as the reference (list[str], eos_token) -> FloatTensor; the DeBERTa-v3 RM is built lazily."""
    global _reward
    if _reward is None:
        _reward = entry.load_reward_func(training_args)
    return _reward.score_strings(pmt_and_responses, eos_token)


if __name__ == "__main__":
    training_args.apply_overrides()
    entry.prepare_output_dir(training_args)
    tokenizer, policy, ref_policy = entry.load_tokenizer_and_policies(training_args)
    train_dataset = entry.load_prompt_dataset(training_args, tokenizer)
    rf = entry.load_reward_func(training_args)           # id-level fast path when no RM tokenizer is on disk
    trainer = RemaxTrainer(
        config=training_args,
        processing_class=tokenizer,
        policy=policy,
        ref_policy=ref_policy,
        train_dataset=train_dataset,
        reward_func=rf if rf.accepts_ids else reward_func,
        callbacks=[EarlyStoppingCallback(early_stopping_patience=training_args.early_stopping_patience)],
    )
    trainer.train()
