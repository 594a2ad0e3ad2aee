"""Phase micro-profiles (torch.profiler kernel tables): (a) one decode step at S sequences / ctx tokens,
(b) one training micro-batch fwd+bwd (LoRA, checkpointing), (c) one DeBERTa reward batch.
Output: gpurun_out/profile_phases.txt (top kernels by device time per phase)."""
import math
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanorlhf_b200.models.deberta_v3 import DebertaV3Config, DebertaV3ForSequenceClassification  # noqa: E402
from nanorlhf_b200.models.lora import LoraConfig, get_peft_model  # noqa: E402
from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM, response_logprobs  # noqa: E402
from nanorlhf_b200.ops import native  # noqa: E402
from nanorlhf_b200.sampler.native_sampler import NativeSampler  # noqa: E402

dev = torch.device("cuda")
which = sys.argv[1:] or ["decode", "train", "reward"]
out_lines = []


def table(prof, title, n=22):
    rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0 and not e.key.startswith(("aten::", "autograd::", "_")) and "Backward" not in e.key]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    out_lines.append(f"==== {title}: total kernel time {tot:.2f} ms")
    for k, ms, c in rows[:n]:
        out_lines.append(f"  {ms:9.3f} ms {100 * ms / tot:5.1f}%  n={c:<6d} {k[:110]}")
    print("\n".join(out_lines[-(n + 1):]), flush=True)


shape = Qwen2Config.qwen2_5_1_5b()
policy = get_peft_model(Qwen2ForCausalLM.from_config(shape, torch.bfloat16, dev, seed=0),
                        LoraConfig(r=64, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head", "score"]))

if "decode" in which:
    S, ctx, max_tokens = int(os.environ.get("S", "2048")), int(os.environ.get("CTX", "1000")), 16
    eng = NativeSampler(policy, use_cuda_graph=False, rollout_dtype=os.environ.get("ROLLOUT", "bf16"),
                        kv_cache_dtype=os.environ.get("KV", "bf16"))
    eng.sync_weights()
    per = (ctx + max_tokens) // 16 + 2
    eng._ensure_kv(S * per + 8)
    st = eng._alloc_state(S, per, max_tokens, shape.vocab_size - 1)
    st["block_tables"].copy_(torch.arange(S * per, device=dev, dtype=torch.int32).view(S, per))
    st["positions"].fill_(ctx - 1)
    st["ctx_lens"].fill_(ctx)
    st["finished"].zero_()
    st["tokens"].copy_(torch.randint(0, 1000, (S,), device=dev, dtype=torch.int32))
    for _ in range(2):
        eng._decode_step(st, 0.9, 0.95, 1, None, shape.vocab_size - 1, 10**6)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            eng._decode_step(st, 0.9, 0.95, 1, None, shape.vocab_size - 1, 10**6)
        torch.cuda.synchronize()
    table(prof, f"decode x4 steps, S={S}, ctx={ctx}, rollout_dtype={os.environ.get('ROLLOUT', 'bf16')}, kv={os.environ.get('KV', 'bf16')}")
    del eng, st
    torch.cuda.empty_cache()

if "train" in which:
    policy.train()
    if os.environ.get("GC", "1") == "1":
        policy.gradient_checkpointing_enable()
    pad = shape.vocab_size - 1
    qr = torch.randint(0, 150000, (4, 1650), device=dev)
    params = [p for p in policy.parameters() if p.requires_grad]

    def step():
        lp, ent = response_logprobs(policy, qr, 150, pad, 0.9, want_entropy=True)[:2]
        (lp.mean()).backward()
    step()
    for p in params:
        p.grad = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    out_lines.append(f"train micro-batch wall (CUDA events, 3 steps): {e0.elapsed_time(e1) / 3:.2f} ms/step")
    print(out_lines[-1], flush=True)
    for p in params:
        p.grad = None
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    table(prof, "train micro-batch fwd+bwd (4 x 1650 tokens, LoRA r64 + embed/lm_head, grad checkpointing)", 30)
    for p in params:
        p.grad = None
    # no-grad logprob pass over 30 sequences
    qr2 = torch.randint(0, 150000, (30, 1650), device=dev)
    policy.eval()
    with torch.no_grad():
        response_logprobs(policy, qr2, 150, pad, 0.9)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            response_logprobs(policy, qr2, 150, pad, 0.9)
            torch.cuda.synchronize()
    table(prof, "no-grad logprob pass (30 x 1650 tokens, policy with LoRA)", 16)

if "reward" in which:
    rm = DebertaV3ForSequenceClassification.from_config(DebertaV3Config.large(), torch.bfloat16, dev, seed=1)
    ids = torch.randint(3, 100000, (16, 1660), device=dev)
    with torch.no_grad():
        rm(ids)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            rm(ids)
            torch.cuda.synchronize()
    table(prof, "DeBERTa-v3-large reward forward (16 x 1660 tokens)", 16)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "profile_phases.txt"), "a") as f:
    f.write("\n".join(out_lines) + "\n")
