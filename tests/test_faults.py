"""Failure detection + restart: kill one rank mid-update (fault injection), the survivor must exit non-zero,
and a relaunch resumes from the last checkpoint (the reference has none of this)."""
import os
import tempfile
import time

import pytest
import torch
import torch.multiprocessing as mp


def _worker(rank, world, port, tmp, fault):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if fault:
        os.environ["NANORLHF_FAULT"] = fault
    else:
        os.environ.pop("NANORLHF_FAULT", None)
    torch.set_num_threads(2)
    from nanorlhf_b200.config import RLConfig
    from nanorlhf_b200.models.lora import LoraConfig, get_peft_model
    from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM
    from nanorlhf_b200.parallel.comm import Comm
    from nanorlhf_b200.reward.api import LengthReward
    from nanorlhf_b200.trainer import ReinforceTrainer
    from nanorlhf_b200.utils.data import synthetic_hh_dataset
    from nanorlhf_b200.utils.tokenizer import ByteTokenizer
    comm = Comm.from_env(torch.device("cpu"), timeout_s=20)
    tok = ByteTokenizer()
    cfg = Qwen2Config.tiny(vocab_size=tok.vocab_size)
    policy = get_peft_model(Qwen2ForCausalLM.from_config(cfg, torch.float32, seed=1), LoraConfig(r=4, lora_alpha=8, modules_to_save=None))
    ref = Qwen2ForCausalLM.from_config(cfg, torch.float32, seed=1)
    a = RLConfig(output_dir=tmp, response_length=6, per_device_train_batch_size=2, gradient_accumulation_steps=1, num_mini_batches=1,
                 total_episodes=12, learning_rate=1e-3, sampler="torch", report_to="none", watchdog_timeout_s=15)
    a.quiet = True
    t = ReinforceTrainer(a, tok, policy, ref, synthetic_hh_dataset(tok, 32, max_prompt_tokens=16), reward_func=LengthReward(4), comm=comm)
    start = t.state.global_step
    t.train()
    if rank == 0:
        with open(os.path.join(tmp, "done.txt"), "w") as f:
            f.write(f"{t.state.global_step} resumed={t._resumed}")
    comm.close()


def test_rank_failure_is_detected_and_restart_resumes():
    port = 29800 + os.getpid() % 100
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.time()
        with pytest.raises(Exception):
            mp.spawn(_worker, args=(2, port, tmp, "1:reward:2:exit"), nprocs=2, join=True)     # rank 1 dies in update 2
        assert time.time() - t0 < 120, "survivor did not notice the dead rank in time"
        assert os.path.isdir(os.path.join(tmp, "checkpoint-1")) and not os.path.exists(os.path.join(tmp, "done.txt"))
        mp.spawn(_worker, args=(2, port + 1, tmp, ""), nprocs=2, join=True)                     # relaunch: resume=auto
        assert open(os.path.join(tmp, "done.txt")).read() == "3 resumed=True"
