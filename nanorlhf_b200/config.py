"""Standalone configuration dataclasses.

The reference subclasses trl's ``PPOConfig`` (itself an HF ``TrainingArguments``) once per
algorithm and instantiates it at module import inside the entry script
(/root/reference/GRPO/grpo.py:86-155, PPO/ppo.py:78-166).  trl / HF ``Trainer`` are not
dependencies here, so ``RLConfig`` owns every field the reference consumes (SURVEY.md section 2.2
``PPOConfig`` row and App. A) with the same names and defaults, promotes the constants the
reference hard-codes (top-p, token budgets, scratch dirs ...) to fields, and adds
``--key=value`` CLI / ``NANORLHF_KEY`` env overrides (the reference has neither).
"""
from __future__ import annotations

import dataclasses
import json
import os
import sys
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

DEFAULT_LORA_TARGETS = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]


@dataclass
class RLConfig:
    # ---- experiment ------------------------------------------------------------------------
    exp_name: str = "rlhf"
    seed: int = 42
    output_dir: str = "outputs/rlhf"
    logging_dir: Optional[str] = None
    run_name: Optional[str] = None
    report_to: Any = "none"                     # "none" | "wandb" | "tensorboard" | list
    overwrite: bool = False                     # reference rmtree()s output_dir (GRPO/grpo.py:204)
    resume: str = "auto"                        # auto | never | <checkpoint path>

    # ---- models ----------------------------------------------------------------------------
    sft_model_path: str = "Qwen/Qwen2.5-1.5B-Instruct"
    reward_model_path: str = "OpenAssistant/reward-model-deberta-v3-large-v2"
    bf16: bool = True
    gradient_checkpointing: bool = True
    gradient_checkpointing_kwargs: Optional[dict] = None

    # ---- LoRA (GRPO/grpo.py:90-99) -----------------------------------------------------------
    use_lora: bool = True
    lora_r: int = 64
    lora_alpha: int = 16
    lora_dropout: float = 0.0
    lora_bias: str = "none"
    lora_target_modules: List[str] = field(default_factory=lambda: list(DEFAULT_LORA_TARGETS))
    modules_to_save: Optional[List[str]] = field(default_factory=lambda: ["embed_tokens", "lm_head", "score"])

    # ---- data ------------------------------------------------------------------------------
    train_dataset_name: str = "Anthropic/hh-rlhf"
    train_dataset_split: str = "train[:100%]"
    dataset_num_proc: int = 6
    dataloader_drop_last: bool = True

    # ---- rollout ---------------------------------------------------------------------------
    response_length: int = 1500
    temperature: float = 0.9
    top_p: float = 0.95                         # hard-coded in the reference (grpo_trainer.py:127)
    logprob_top_p_consistent: bool = False      # experimental: score with the top-p-truncated, renormalised softmax the sampler draws
                                                # from (the reference samples with top_p but scores with the full softmax); slow path
    stop_token: Optional[str] = "eos"
    stop_token_id: Optional[int] = None
    missing_eos_penalty: Optional[float] = None
    changing_seed: bool = True                  # seed=random.randint(1,5000) per rollout (:127)
    rollout_dtype: str = "bf16"                 # sampler GEMMs: bf16 | fp8 (e4m3, per-token x per-channel scales)
    kv_cache_dtype: str = "fp8"                 # sampler KV pages: fp8 (e4m3 + per-token scales: half the bytes of the HBM-bound
                                                # decode attention, 1.6x faster, 2x capacity) | bf16.  Old log-probs are always
                                                # recomputed by the bf16 training engine, so this only shapes the samples.
    sampler: str = "auto"                       # auto | native | torch
    kv_block_size: int = 16

    # ---- reward ----------------------------------------------------------------------------
    reward_batch_size: int = 16
    reward_dtype: str = "bf16"                  # parity switch: "fp32" (reference keeps RM in fp32)

    # ---- algorithm -------------------------------------------------------------------------
    kl_coef: float = 0.01
    cliprange: float = 0.2
    cliprange_value: float = 0.2
    vf_coef: float = 0.1
    gamma: float = 1.0
    lam: float = 0.95
    whiten_rewards: bool = False
    advantage_whiten: bool = False
    num_ppo_epochs: int = 1
    num_mini_batches: int = 16
    total_episodes: int = 250000
    local_rollout_forward_batch_size: int = 16  # ignored by the reference too (token budget instead)
    token_budget_fwd: int = 22 * 2316           # padded-token budget of a no-grad forward (:534)
    token_budget_train: int = 4 * 2316          # r1 micro-bucket budget (grpo_r1_trainer.py:700)
    train_samples_per_prompt: int = 1           # GRPO/RLOO keep 1 random of N (grpo_trainer.py:504)
    stats_include_padding: bool = False         # parity: entropy/approxkl over padded positions
    grpo_std_eps: float = 0.0
    num_sample_generations: int = 0

    # ---- optimisation ----------------------------------------------------------------------
    learning_rate: float = 6e-6
    weight_decay: float = 0.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    max_grad_norm: Optional[float] = None       # None = never clip (the reference); a value clips the global (all-rank) gradient norm
    warmup_steps: int = 0
    lr_scheduler_type: str = "cosine_with_min_lr"
    lr_scheduler_kwargs: Dict[str, Any] = field(default_factory=lambda: {"min_lr_rate": 0.1})
    per_device_train_batch_size: int = 4
    gradient_accumulation_steps: int = 8
    optimizer_state_dtype: str = "fp32"         # parity switch: "bf16" (reference, GRPO/grpo.py:222)

    # ---- eval / save / log -----------------------------------------------------------------
    eval_strategy: str = "steps"
    eval_steps: int = 1
    save_strategy: str = "steps"
    save_steps: int = 1
    save_total_limit: Optional[int] = 8
    save_only_model: bool = False
    async_checkpoint: bool = True               # CUDA: snapshot to pinned host memory on a side stream, serialise in a background thread
    save_value_model: bool = True
    logging_steps: int = 1
    metric_for_best_model: str = "eval_objective/rlhf_reward_old"
    greater_is_better: bool = True
    load_best_model_at_end: bool = True
    early_stopping_patience: int = 1000000
    disable_tqdm: bool = True
    push_to_hub: bool = False

    # ---- runtime (B200) --------------------------------------------------------------------
    comm: str = "fused"                         # fused (symmetric-memory kernels) | nccl
    ddp_bucket_mb: int = 25                     # comm="nccl": gradient buckets all-reduced from backward hooks on the last micro-step
                                                # of a window (DDP's reducer, 25 MB default); 0 = one all-reduce inside step()
    weight_sync: str = "sharded"                # sharded (K-BC under fused DP: layer-sharded merge + multimem.st into every rank's arena,
                                                # 7.4 ms at 2 GPUs / 5.8 ms at 8, bit-identical to the local merge) | local (every rank
                                                # merges its own arena, 13.7 ms); single-GPU runs always merge locally
    train_cuda_graph: str = "auto"              # auto | on | off : replay the micro-step (fwd+loss+bwd) as a CUDA graph
    offload_policy: str = "resident"            # per-role residency: resident | host
    offload_ref: Optional[str] = None
    offload_reward: Optional[str] = None
    offload_optimizer: Optional[str] = None
    scratch_dir: str = "/tmp/nanorlhf_scratch"  # profiler traces + value pre-fit scratch (reference: /data/temp_vllm_model, /data/cache_value_model)
    profile: str = "none"                       # none | nvtx (ranges per phase) | torch (chrome trace of update `profile_update`)
    profile_update: int = 2
    memory_log: Optional[str] = None            # JSONL path: per-update peak HBM + phase times (r1's unused `memory_log`, grpo_r1.py:98-100)
    watchdog_timeout_s: float = 1800.0

    # ---- derived (filled by the trainer; reference: grpo_trainer.py:220-240) ------------------
    world_size: int = 1
    local_batch_size: int = 0
    micro_batch_size: int = 0
    batch_size: int = 0
    mini_batch_size: int = 0
    local_mini_batch_size: int = 0
    num_total_batches: int = 0

    def __post_init__(self):
        if self.logging_dir is None:
            self.logging_dir = os.path.join(self.output_dir, "logs")

    # ---- helpers ---------------------------------------------------------------------------
    def role_residency(self, role: str) -> str:
        v = getattr(self, f"offload_{role}", None)
        return v or self.offload_policy

    def to_dict(self) -> Dict[str, Any]:
        return dataclasses.asdict(self)

    def to_json_string(self) -> str:
        return json.dumps(self.to_dict(), indent=2, default=str)

    def apply_overrides(self, argv: Optional[List[str]] = None, env: Optional[Dict[str, str]] = None):
        """``--key=value`` / ``--key value`` CLI flags and ``NANORLHF_KEY=value`` env variables."""
        env = os.environ if env is None else env
        argv = sys.argv[1:] if argv is None else argv
        fields = {f.name: f for f in dataclasses.fields(self)}
        pending: Dict[str, str] = {}
        for k, v in env.items():
            if k.startswith("NANORLHF_") and k[9:].lower() in fields and k[9:].lower() != "backend":
                pending[k[9:].lower()] = v
        i = 0
        while i < len(argv):
            a = argv[i]
            if a.startswith("--"):
                if "=" in a:
                    k, v = a[2:].split("=", 1)
                elif i + 1 < len(argv) and not argv[i + 1].startswith("--"):
                    k, v = a[2:], argv[i + 1]
                    i += 1
                else:
                    k, v = a[2:], "true"
                k = k.replace("-", "_")
                if k not in fields:
                    raise ValueError(f"unknown config flag --{k}")
                pending[k] = v
            i += 1
        for k, v in pending.items():
            setattr(self, k, _coerce(v, getattr(self, k), fields[k].type))
        if "output_dir" in pending and "logging_dir" not in pending:
            self.logging_dir = os.path.join(self.output_dir, "logs")
        return self


def _coerce(text: str, current: Any, annot: Any):
    if isinstance(current, bool):
        return text.lower() in ("1", "true", "yes", "on")
    if isinstance(current, int) and not isinstance(current, bool):
        return int(text)
    if isinstance(current, float):
        return float(text)
    if isinstance(current, (list, dict)):
        return json.loads(text)
    if current is None:
        s = str(annot)
        if text.lower() in ("none", "null"):
            return None
        if "int" in s:
            return int(text)
        if "float" in s:
            return float(text)
        if "List" in s or "Dict" in s or "dict" in s:
            return json.loads(text)
    return text


@dataclass
class ValueFinetuneConfig:
    """Value-model pre-fit settings (ref ``Value_Finetune_Config``: PPO/ppo.py:78-110)."""
    train_data_size: int = 500
    train_split_rate: float = 0.8
    num_train_epochs: int = 8
    per_device_train_batch_size: int = 32
    per_device_eval_batch_size: int = 50
    gradient_accumulation_steps: int = 12
    learning_rate: float = 1e-3
    lr_scheduler_type: str = "reduce_lr_on_plateau"
    lr_scheduler_kwargs: Dict[str, Any] = field(default_factory=lambda: {"mode": "min", "factor": 0.5, "patience": 0})
    early_stopping_patience: int = 3
    output_dir: str = "/tmp/nanorlhf_scratch/cache_value_model"
    report_to: Any = "none"
    seed: int = 42
    value_init_truncate: bool = True            # parity switch: reference runs with stop_token_id=None
    weight_decay: float = 0.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    token_budget_fwd: int = 28 * 2316           # value_initializer.py:270
