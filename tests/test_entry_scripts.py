"""The per-algorithm entry scripts keep the reference's public names and run end to end offline (tiny shapes)."""
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("path,cfg,extra", [
    ("GRPO/grpo.py", "GRPOConfig", "grpo_sample_N"), ("RLOO/rloo.py", "RLOOConfig", "rloo_sample_N"),
    ("RAFT/raft.py", "RAFTConfig", "raft_sample_K"), ("ReMax/remax.py", "RemaxConfig", None),
    ("REINFORCE/reinforce.py", "ReinforceConfig", None), ("PPO/ppo.py", "MyPPOConfig", "value_learning_rate"),
    ("examples/r1-v0/grpo_r1.py", "GRPOConfig", "grpo_sample_N"),
])
def test_public_names(path, cfg, extra):
    m = _load(path, "entry_" + cfg + path.replace("/", "_").replace(".", "_").replace("-", "_"))
    assert hasattr(m, cfg) and hasattr(m, "training_args") and callable(m.reward_func)
    a = m.training_args
    assert a.per_device_train_batch_size == 4 and a.gradient_accumulation_steps == 8 and a.num_mini_batches == 16
    assert a.lr_scheduler_type == "cosine_with_min_lr" and a.temperature == 0.9
    if extra:
        assert hasattr(a, extra)
    if "reinforce" in path:
        assert a.advantage_whiten is True
    if "ppo" in path:
        assert hasattr(m, "Value_Finetune_Config") and hasattr(m, "finetune_args") and m.finetune_args.train_data_size == 500
        assert a.vf_coef == 1 and a.lam == 0.95
    if "r1" in path:
        assert a.response_length == 8000 and a.kl_coef == 0.0 and callable(getattr(m, "GRPOTrainer"))


@pytest.mark.parametrize("script,extra", [("REINFORCE/reinforce.py", []), ("GRPO/grpo.py", ["--grpo_sample_N=2"])])
def test_entry_script_runs_offline(script, extra, tmp_path):
    env = dict(os.environ, NANORLHF_MODEL_SHAPE="tiny", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, os.path.join(ROOT, script), f"--output_dir={tmp_path}", "--response_length=8",
           "--per_device_train_batch_size=2", "--gradient_accumulation_steps=1", "--num_mini_batches=2", "--total_episodes=8",
           "--sampler=torch", "--learning_rate=1e-3"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.isdir(os.path.join(tmp_path, "checkpoint-2"))


def test_reference_helper_names_are_importable():
    """module-level helpers a reference user imports from the trainer module."""
    import torch
    from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM
    from nanorlhf_b200.trainer import (INVALID_LOGPROB, PolicyAndValueWrapper, forward, state_to_device,  # noqa: F401
                                       vllm_generate)
    from nanorlhf_b200.utils.tokenizer import ByteTokenizer
    assert INVALID_LOGPROB == 1.0
    tok = ByteTokenizer()
    m = Qwen2ForCausalLM.from_config(Qwen2Config.tiny(vocab_size=tok.vocab_size), torch.float32, seed=0)
    prompts = [[5, 6, 7], [8, 9]]
    a = vllm_generate(2, m, tok, prompts, 0.0, 4, backend="torch")            # GRPO-style call: n first
    b = vllm_generate(m, tok, prompts, 0.0, 4, backend="torch")               # PPO-style call: n = 1
    assert a.shape == (4, 4) and b.shape == (2, 4)
    assert torch.equal(a[0], a[1]) and torch.equal(a[0], b[0])                # greedy: samples of a prompt coincide
