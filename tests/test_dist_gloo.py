"""Host-side data-parallel logic under gloo, world_size 2, on the CPU box ("multi-node without a cluster")."""
import os
import tempfile

import pytest
import torch
import torch.multiprocessing as mp


def _worker(rank, world, port, tmp, algo):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from dataclasses import dataclass
    from nanorlhf_b200.config import RLConfig
    from nanorlhf_b200.models.lora import LoraConfig, get_peft_model
    from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM
    from nanorlhf_b200.parallel.comm import Comm
    from nanorlhf_b200.reward.api import LengthReward
    from nanorlhf_b200.trainer import GRPOTrainer, SparseGRPOTrainer
    from nanorlhf_b200.utils.data import synthetic_hh_dataset
    from nanorlhf_b200.utils.tokenizer import ByteTokenizer
    comm = Comm.from_env(torch.device("cpu"))
    tok = ByteTokenizer()
    cfg = Qwen2Config.tiny(vocab_size=tok.vocab_size)
    policy = get_peft_model(Qwen2ForCausalLM.from_config(cfg, torch.float32, seed=1), LoraConfig(r=4, lora_alpha=8, modules_to_save=None))
    ref = Qwen2ForCausalLM.from_config(cfg, torch.float32, seed=1)
    a = RLConfig(output_dir=tmp, response_length=8, per_device_train_batch_size=2, gradient_accumulation_steps=1,
                 num_mini_batches=2, total_episodes=16, learning_rate=1e-3, sampler="torch", report_to="none")
    a.grpo_sample_N = 4
    a.quiet = True
    cls = SparseGRPOTrainer if algo == "sparse" else GRPOTrainer
    # rank-dependent reward => rank-dependent number of surviving rows in sparse GRPO (deadlock regression)
    reward = LengthReward(4 + 3 * rank)
    t = cls(a, tok, policy, ref, synthetic_hh_dataset(tok, 32, max_prompt_tokens=20), reward_func=reward, comm=comm)
    t.train()
    flat = torch.cat([p.detach().reshape(-1) for p in t.policy.parameters() if p.requires_grad])
    both = comm.all_gather_cat(flat[None])
    assert torch.equal(both[0], both[1]), "ranks diverged: gradients were not averaged identically"
    assert t.state.global_step == 2 and t.state.episode == 16
    assert t.optimizer.comm_mode == "nccl" and t.optimizer.overlap_launched > 0      # bucket all-reduces came from backward hooks
    comm.close()


@pytest.mark.parametrize("algo", ["grpo", "sparse"])
def test_two_rank_training_keeps_replicas_identical(algo):
    port = 29600 + (os.getpid() % 200) + (0 if algo == "grpo" else 1)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, port, tmp, algo), nprocs=2, join=True)


def _bcast_worker(rank, world, port, q):
    import os
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from nanorlhf_b200.parallel.comm import Comm
    comm = Comm.from_env(torch.device("cpu"))
    torch.manual_seed(100 + rank)                      # replicas deliberately start different
    m = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4)).to(torch.float32)
    m[2].weight.requires_grad_(False)
    m.register_buffer("step", torch.tensor([rank], dtype=torch.int64))
    comm.broadcast_module_(m, 0)
    sig = torch.cat([p.detach().reshape(-1) for p in m.parameters()] + [m.step.float()]).double().sum().item()
    q.put((rank, sig))
    comm.barrier()
    comm.close()


def test_broadcast_module_makes_replicas_identical():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert got[0] == got[1]


def _overlap_worker(rank, world, port, q):
    import os
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from nanorlhf_b200.parallel.comm import Comm
    from nanorlhf_b200.parallel.optimizer import FusedAdamW
    comm = Comm.from_env(torch.device("cpu"))

    def run(overlap: bool):
        torch.manual_seed(0)
        m = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, 4))
        opt = FusedAdamW([{"params": list(m.parameters())}], lr=1e-2, comm=comm, comm_mode="nccl", master_weights=False)
        if overlap:
            assert opt.enable_bucket_overlap(bucket_bytes=1024)            # 3 buckets: one per Linear
            assert len(opt._buckets[0]) >= 3
        g = torch.Generator().manual_seed(10 + rank)                       # every rank sees different data
        for _step in range(3):
            opt.zero_grad()
            for micro in range(2):                                          # accumulation window of 2
                x = torch.randn(8, 16, generator=g)
                if overlap and micro == 1:
                    opt.arm_overlap()
                (m(x).square().mean() / 2).backward()
            opt.step()
        return torch.cat([p.detach().reshape(-1) for p in m.parameters()]), opt.overlap_launched

    p_plain, n0 = run(False)
    p_overlap, n1 = run(True)
    q.put((rank, bool(torch.equal(p_plain, p_overlap)), n0, n1, p_overlap.double().sum().item()))
    comm.barrier()
    comm.close()


def test_bucketed_overlap_equals_single_allreduce():
    """comm="nccl": all-reducing ~25 MB buckets from backward hooks on the last micro-step (DDP's reducer, reference
    GRPO/grpo_trainer.py:690) gives bit-identical parameters to one all-reduce inside step(); every bucket is launched
    during the backward; earlier micro-steps of the window do not communicate."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + (os.getpid() % 100)
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert all(same for _r, same, _n0, _n1, _s in res), res
    assert all(n0 == 0 and n1 >= 9 for _r, _same, n0, n1, _s in res), res         # 3 steps x >= 3 buckets, all from hooks
    assert res[0][4] == res[1][4]                                                # replicas stay identical
