"""Shared plumbing of the per-algorithm entry scripts (GRPO/grpo.py ...).

The reference's scripts download models / datasets from the HF hub at import
(/root/reference/GRPO/grpo.py:159,209-224,247).  The GPU box has no network, so every loader here
takes a local directory when one exists and otherwise falls back to a *synthetic* stand-in of the same
shape (random-init weights of the named architecture, byte-level tokenizer, hh-rlhf-shaped prompts) and
says so on stdout.  ``NANORLHF_MODEL_SHAPE`` (tiny | 125m | 1.5b | 7b) overrides the architecture.
"""
from __future__ import annotations

import copy
import os
import shutil
from typing import Optional

import torch

from .models.deberta_v3 import DebertaV3Config, DebertaV3ForSequenceClassification
from .models.lora import LoraConfig, get_peft_model
from .models.qwen2 import Qwen2Config, Qwen2ForCausalLM, Qwen2ForSequenceClassification
from .reward.model_reward import ModelReward
from .utils.data import (QWEN_CHAT_TEMPLATE, extract_hh_question, prepare_prompt_dataset, synthetic_hh_questions)
from .utils.tokenizer import ByteTokenizer, load_tokenizer

_SHAPES = {"tiny": Qwen2Config.tiny, "125m": Qwen2Config.plumbing_125m, "1.5b": Qwen2Config.qwen2_5_1_5b,
           "7b": Qwen2Config.qwen2_5_7b}


def default_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    return torch.device("cpu")


def prepare_output_dir(args):
    """The reference ``rmtree``s output_dir on every launch (GRPO/grpo.py:204); here that needs --overwrite."""
    if args.overwrite and int(os.environ.get("RANK", "0")) == 0:
        shutil.rmtree(args.output_dir, ignore_errors=True)
    os.makedirs(args.output_dir, exist_ok=True)


def _shape_for(name: str, vocab_size: Optional[int]) -> Qwen2Config:
    forced = os.environ.get("NANORLHF_MODEL_SHAPE")
    if forced:
        key = forced.lower()
    elif "7b" in name.lower():
        key = "7b"
    elif "1.5b" in name.lower():
        key = "1.5b"
    else:
        key = "125m" if not torch.cuda.is_available() else "1.5b"
    cfg = _SHAPES[key]()
    if vocab_size is not None and vocab_size > cfg.vocab_size or key in ("tiny", "125m"):
        cfg.vocab_size = max(vocab_size or 0, 320)
    return cfg


def load_tokenizer_and_policies(args, dtype=torch.bfloat16, device=None):
    """Returns (tokenizer, policy, ref_policy).  LoRA is applied to the policy when ``args.use_lora``."""
    device = device or default_device()
    path = args.sft_model_path
    local = os.path.isdir(path) and os.path.exists(os.path.join(path, "config.json"))
    tokenizer = load_tokenizer(path)
    if tokenizer.pad_token_id is None or tokenizer.pad_token_id == tokenizer.eos_token_id:
        tokenizer.add_special_tokens({"pad_token": "[PAD]"})
    tokenizer.padding_side = "left"
    if local:
        policy = Qwen2ForCausalLM.from_pretrained(path, dtype, device)
        ref_policy = Qwen2ForCausalLM.from_pretrained(path, dtype, device)
    else:
        cfg = _shape_for(path, len(tokenizer))
        cfg.name_or_path = path
        print(f"[entry] '{path}' is not a local checkpoint: random-init {cfg.num_hidden_layers}L d={cfg.hidden_size} "
              f"vocab={cfg.vocab_size} (synthetic weights)")
        if dtype == torch.bfloat16 and device.type == "cpu":
            dtype = torch.float32
        policy = Qwen2ForCausalLM.from_config(cfg, dtype, device, seed=args.seed)
        ref_policy = Qwen2ForCausalLM.from_config(cfg, dtype, device, seed=args.seed)
    ref_policy.eval()
    if args.use_lora:
        policy = get_peft_model(policy, LoraConfig(r=args.lora_r, lora_alpha=args.lora_alpha, target_modules=args.lora_target_modules,
                                                   lora_dropout=args.lora_dropout, bias=args.lora_bias, task_type="CAUSAL_LM",
                                                   modules_to_save=args.modules_to_save, base_model_name_or_path=path))
        policy.print_trainable_parameters()
        if args.gradient_checkpointing:
            policy.enable_input_require_grads()
    return tokenizer, policy, ref_policy


def load_value_model(args, ref_policy, use_lora: bool, lora_kwargs: dict):
    """Critic = policy backbone + score head (PPO/ppo.py:280-287), optionally LoRA-wrapped (:316-332)."""
    vm = Qwen2ForSequenceClassification.from_causal_lm(ref_policy)
    if use_lora:
        vm = get_peft_model(vm, LoraConfig(**lora_kwargs))
    return vm


def load_reward_func(args, device=None, tiering=None):
    """DeBERTa-v3 reward callback; random-init large model when the checkpoint is not on disk."""
    device = device or default_device()
    path = args.reward_model_path
    dtype = torch.float32 if (args.reward_dtype == "fp32" or device.type == "cpu") else torch.bfloat16
    if os.path.isdir(path) and os.path.exists(os.path.join(path, "config.json")):
        rm = DebertaV3ForSequenceClassification.from_pretrained(path, dtype, device)
        from transformers import AutoTokenizer
        rm_tok = AutoTokenizer.from_pretrained(path)
    else:
        small = os.environ.get("NANORLHF_MODEL_SHAPE", "").lower() in ("tiny", "125m") or device.type == "cpu"
        cfg = DebertaV3Config.tiny(vocab_size=1024) if small else DebertaV3Config.large()
        print(f"[entry] reward model '{path}' not on disk: random-init DeBERTa-v3 "
              f"({cfg.num_hidden_layers}L d={cfg.hidden_size}), id-level scoring")
        rm = DebertaV3ForSequenceClassification.from_config(cfg, dtype, device, seed=1234)
        rm_tok = None
    return ModelReward(rm, rm_tok, args.reward_batch_size, device, tiering)


def load_prompt_dataset(args, tokenizer, template: str = QWEN_CHAT_TEMPLATE, n_synthetic: int = 4096,
                        field: str = "chosen", extract=extract_hh_question):
    """hh-rlhf prompts from a local copy of the dataset when available, else synthetic ones."""
    name = args.train_dataset_name
    questions = None
    if os.path.isdir(name):
        try:
            from datasets import load_dataset
            ds = load_dataset(name, split=args.train_dataset_split)
            questions = [extract(r[field]) if extract else r[field] for r in ds]
        except Exception as e:
            print(f"[entry] could not read local dataset {name}: {e}")
    if questions is None:
        print(f"[entry] dataset '{name}' not on disk: {n_synthetic} synthetic hh-rlhf-shaped prompts")
        questions = synthetic_hh_questions(n_synthetic, seed=args.seed)
    return prepare_prompt_dataset(questions, tokenizer, template)
