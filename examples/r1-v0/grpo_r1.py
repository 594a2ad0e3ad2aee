"""r1-v0: "sparse GRPO" + dynamic mini-batching with a rule-based math reward and 8000-token responses.

Reference: /root/reference/examples/r1-v0/grpo_r1.py (config :96-155, reward :250-273, accuracy probe
:276-341) and grpo_r1_trainer.py.  Same surface: ``GRPOConfig`` / ``training_args``,
``reward_func(pmt_and_responses, responses_ids, tokenizer)``, ``accuracy_func(model, args) -> float`` and
``GRPOTrainer(..., accuracy_func=...)``; the trainer is the data-parallel-safe ``SparseGRPOTrainer``.

Run:  python examples/r1-v0/grpo_r1.py [--key=value ...]   |   torchrun --nproc-per-node 8 examples/r1-v0/grpo_r1.py
Datasets (MetaMathQA for training, MATH-500 for the accuracy probe) are read from local directories when
present; otherwise synthetic arithmetic problems with exact answers stand in (no network on the GPU box).
"""
import os
import sys
from dataclasses import dataclass
from typing import Optional

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from nanorlhf_b200 import entry
from nanorlhf_b200.config import RLConfig
from nanorlhf_b200.reward import rule_math
from nanorlhf_b200.trainer import SparseGRPOTrainer
from nanorlhf_b200.utils.callbacks import EarlyStoppingCallback
from nanorlhf_b200.utils.data import R1_TEMPLATE, prepare_prompt_dataset

os.environ.setdefault("WANDB_PROJECT", "r1-v0")
base_model = "Qwen/Qwen2-1.5B"          # the reference script uses Qwen2-1.5B (grpo_r1.py:92)


@dataclass
class GRPOConfig(RLConfig):
    grpo_sample_N: int = 4
    memory_log: Optional[str] = None     # path of a JSONL file: per-update peak allocated / reserved HBM and phase times
    accuracy_before_train: bool = True
    q_lora: bool = False                 # 4-bit base weights (bitsandbytes) in the reference; rejected here, see __main__
    accuracy_dataset_name: str = "HuggingFaceH4/MATH-500"
    eval_every: int = 10                 # hard-coded in the reference (grpo_r1_trainer.py:824)
    reward_match: str = "equiv"          # "exact" = the shipped reference's effective rule
    advantage_whiten: bool = False


training_args = GRPOConfig(
    exp_name="r1-v0",
    sft_model_path=base_model,
    output_dir=f"{base_model}/{os.environ['WANDB_PROJECT']}",
    kl_coef=0.0, cliprange=0.2, temperature=0.9, response_length=8000,
    per_device_train_batch_size=4, gradient_accumulation_steps=8, num_mini_batches=16, num_ppo_epochs=1,
    total_episodes=250000, learning_rate=9e-6, lr_scheduler_type="cosine_with_min_lr", lr_scheduler_kwargs={"min_lr_rate": 0.1},
    bf16=True, gradient_checkpointing=True,
    use_lora=True, lora_r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head", "score"],
    report_to="none", save_steps=1, save_total_limit=6, logging_steps=1, eval_steps=1,
    metric_for_best_model="eval_objective/rlhf_reward_old", greater_is_better=True,
    stop_token="eos", train_dataset_name="meta-math/MetaMathQA", train_dataset_split="train[:100%]",
    token_budget_fwd=22 * 2316, token_budget_train=4 * 2316,
)


def load_problems(args):
    """(train rows with query/response, eval rows with problem/answer)."""
    train = evals = None
    try:
        from datasets import load_dataset
        if os.path.isdir(args.train_dataset_name):
            train = [dict(r) for r in load_dataset(args.train_dataset_name, split=args.train_dataset_split)]
        if os.path.isdir(args.accuracy_dataset_name):
            evals = [dict(r) for r in load_dataset(args.accuracy_dataset_name, split="test")]
    except Exception as e:
        print(f"[r1] local datasets unavailable: {e}")
    if train is None:
        print("[r1] MetaMathQA not on disk: synthetic arithmetic problems")
        train = rule_math.synthetic_arithmetic_problems(4096, seed=0)
    if evals is None:
        evals = rule_math.synthetic_arithmetic_problems(100, seed=1)
    return train, evals


class R1Trainer(SparseGRPOTrainer):
    """Adds the accuracy probe before training and every ``eval_every`` updates (grpo_r1_trainer.py:471-475,824-840)."""

    def before_training(self):
        if self.accuracy_func is not None and self.args.accuracy_before_train:
            acc = float(self.accuracy_func(self.model, self.args))
            self.log({"initial_accuracy": acc})

    def after_update(self, update, metrics):
        if self.accuracy_func is not None and update % self.args.eval_every == 0:
            acc = float(self.accuracy_func(self.model, self.args))
            self.log({"eval_accuracy_new": acc,
                      "eval_response_length": getattr(self.accuracy_func, "last_mean_response_tokens", 0.0)})


GRPOTrainer = R1Trainer      # the name the reference exports

_answers = {}


def reward_func(pmt_and_responses, responses_ids, tokenizer):
    """1.0 when the \\boxed{} answer of the response matches the ground truth, else 0.0."""
    return rule_math.RuleMathReward(_answers, match=training_args.reward_match)(pmt_and_responses, responses_ids, tokenizer)


class _StringRewardAdapter:
    """The trainer calls reward callbacks as (texts, eos_token); the r1 callback wants (texts, ids, tokenizer)."""

    def __init__(self, tokenizer):
        self.tokenizer = tokenizer

    def __call__(self, pmt_and_responses, eos_token):
        return reward_func(pmt_and_responses, None, self.tokenizer)


if __name__ == "__main__":
    training_args.apply_overrides()
    if training_args.q_lora:
        # reference: prepare_model_for_kbit_training on a bitsandbytes 4-bit base (grpo_r1.py:378-381).  A 1.5B / 7B bf16 base
        # is 3 / 15 GB of a 180 GB B200, so the memory saving buys nothing here and no 4-bit GEMM path is built.
        raise NotImplementedError("q_lora=True (4-bit base weights) is not supported: keep the frozen base in bf16 on B200")
    entry.prepare_output_dir(training_args)
    tokenizer, policy, ref_policy = entry.load_tokenizer_and_policies(training_args)
    train_rows, eval_rows = load_problems(training_args)
    for r in train_rows:
        ans = rule_math.extract_answer_is(r["response"])
        if ans is not None:
            _answers[r["query"]] = ans
    train_dataset = prepare_prompt_dataset([r["query"] for r in train_rows], tokenizer, R1_TEMPLATE)
    accuracy_func = rule_math.make_accuracy_func(eval_rows, tokenizer, R1_TEMPLATE, max_tokens=training_args.response_length,
                                                 match=training_args.reward_match)
    trainer = GRPOTrainer(
        config=training_args, processing_class=tokenizer, policy=policy, ref_policy=ref_policy,
        train_dataset=train_dataset, reward_func=_StringRewardAdapter(tokenizer), accuracy_func=accuracy_func,
        callbacks=[EarlyStoppingCallback(early_stopping_patience=training_args.early_stopping_patience)],
    )
    trainer.train()
    rule_math.shutdown_pool()
