"""Data-parallel equivalence on >= 2 GPUs (torchrun): W ranks x (G / W rows each) through the fused K-AR optimizer step
must land on the same parameters as ONE rank accumulating the W row groups locally.

Every rank builds the same tiny LoRA policy and the same batch of G sequences; rank r runs forward + policy loss + backward
on rows r::W and calls FusedAdamW.step() (reduce-scatter + AdamW + all-gather in one kernel).  Rank 0 also replays all W
groups on a private copy with gradient accumulation (loss / W) and a local optimizer.  Adam's first step is lr * g / |g|, so a
handful of elements whose gradient rounds across zero may differ by 2 lr; everything else must agree to float rounding."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanorlhf_b200 import ops  # noqa: E402
from nanorlhf_b200.models.lora import LoraConfig, get_peft_model  # noqa: E402
from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM, response_logprobs  # noqa: E402
from nanorlhf_b200.parallel.comm import Comm  # noqa: E402
from nanorlhf_b200.parallel.optimizer import FusedAdamW, build_param_groups  # noqa: E402

comm = Comm.from_env()
dev, W, R = comm.device, comm.world_size, comm.rank
cfg = Qwen2Config(vocab_size=4096, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                  num_key_value_heads=1, head_dim=128, tie_word_embeddings=True)
LR, CTX, TR, G = 1e-3, 16, 48, 4 * W


def build():
    m = get_peft_model(Qwen2ForCausalLM.from_config(cfg, torch.bfloat16, dev, seed=11),
                       LoraConfig(r=8, lora_alpha=16, modules_to_save=["embed_tokens", "lm_head"]))
    with torch.no_grad():
        g = torch.Generator(device=dev).manual_seed(3)
        for mod in m.modules():
            if hasattr(mod, "lora_B"):
                mod.lora_B.weight.copy_(torch.randn(mod.lora_B.weight.shape, generator=g, device=dev) * 0.02)
    return m.train()


g = torch.Generator().manual_seed(5)
qr = torch.randint(5, 4000, (G, CTX + TR), generator=g).to(dev)
old = (torch.randn(G, TR, generator=g) * 0.05 - 8.0).to(dev)
adv = torch.randn(G, 1, generator=g).expand(G, TR).contiguous().to(dev)
mask = torch.ones(G, TR, dtype=torch.bool, device=dev)


def micro(model, rows, scale):
    lp = response_logprobs(model, qr[rows], CTX, 4095, 0.9, want_entropy=False)[0]
    loss, _ = ops.policy_loss_token(lp, old[rows], adv[rows], mask[rows], 0.2)
    (loss * scale).backward()


dp = build()
comm.broadcast_module_(dp, 0)
opt = FusedAdamW(build_param_groups(dp.named_parameters(), 0.0, LR), lr=LR, comm=comm, comm_mode="fused")
opt.zero_grad()
micro(dp, torch.arange(R, G, W, device=dev), 1.0)
opt.step()
torch.cuda.synchronize()
got = torch.cat([f.param.float() for f in opt.flats])

res = {"world": W}
ok = True
if comm.is_main:
    ref = build()
    ropt = FusedAdamW(build_param_groups(ref.named_parameters(), 0.0, LR), lr=LR)
    before = torch.cat([f.param.float() for f in ropt.flats]).clone()
    ropt.zero_grad()
    for r in range(W):
        micro(ref, torch.arange(r, G, W, device=dev), 1.0 / W)
    ropt.step()
    want = torch.cat([f.param.float() for f in ropt.flats])
    diff = (got - want).abs()
    moved = (want - before).abs()
    res.update(max_abs_diff=float(diff.max()), mean_abs_diff=float(diff.mean()), mean_step=float(moved.mean()),
               frac_off_by_more_than_half_lr=float((diff > 0.5 * LR).float().mean()))
    ok = res["frac_off_by_more_than_half_lr"] < 0.01 and res["mean_abs_diff"] < 0.05 * res["mean_step"] + 1e-6
    res["ok"] = bool(ok)
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"dp_equiv_{W}.json"), "w"))
# parameters must be bit-identical on every rank after the fused all-gather
mine = got.to(torch.bfloat16).view(torch.int16).long().sum()
both = torch.stack([mine, -mine])
comm.all_reduce_(both, "max")
same = bool((both[0] == -both[1]).item())
v = torch.tensor([1 if (ok and same) else 0], device=dev)
comm.all_reduce_(v, "min")
del opt
torch.cuda.synchronize()
comm.barrier()
comm.close()
sys.stdout.flush()
os._exit(0 if bool(v.item()) else 1)
