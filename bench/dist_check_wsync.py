"""2+ GPU check of the sharded weight refresh (K-BC) against the local merge; times both.  torchrun launch."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanorlhf_b200.models.lora import LoraConfig, LoraLinear, get_peft_model  # noqa: E402
from nanorlhf_b200.models.qwen2 import Qwen2Config, Qwen2ForCausalLM  # noqa: E402
from nanorlhf_b200.parallel.comm import Comm  # noqa: E402
from nanorlhf_b200.parallel.weight_sync import ShardedWeightSync, refresh_sampler_arena  # noqa: E402
from nanorlhf_b200.sampler.native_sampler import NativeSampler  # noqa: E402

comm = Comm.from_env()
dev = comm.device
cfg = Qwen2Config.qwen2_5_1_5b()
policy = get_peft_model(Qwen2ForCausalLM.from_config(cfg, torch.bfloat16, dev, seed=0), LoraConfig(r=64, lora_alpha=16, modules_to_save=None))
with torch.no_grad():
    g = torch.Generator(device=dev).manual_seed(5)
    for m in policy.modules():
        if isinstance(m, LoraLinear):
            m.lora_B.weight.copy_(torch.randn(m.lora_B.weight.shape, generator=g, device=dev, dtype=torch.float32).mul_(0.02))
local = NativeSampler(policy)
refresh_sampler_arena(local)
want = [(lw.wqkv.clone(), lw.wo.clone(), lw.wgu.clone(), lw.wdown.clone()) for lw in local.layers]
shared = NativeSampler(policy)
sync = ShardedWeightSync(shared, comm)
sync.refresh()
torch.cuda.synchronize()
ok = all(torch.equal(a, b) for lw, w in zip(shared.layers, want) for a, b in zip((lw.wqkv, lw.wo, lw.wgu, lw.wdown), w))


def timed(fn, iters=5):
    fn()
    comm.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device=dev, dtype=torch.float64)
    comm.all_reduce_(t, "max")
    return float(t)


res = {"world": comm.world_size, "identical_to_local_merge": bool(ok), "local_merge_ms": timed(lambda: refresh_sampler_arena(local)),
       "sharded_peer_store_ms": timed(sync.refresh), "arena_gb": sync.flat.numel() * 2 / 1e9}
if comm.is_main:
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"wsync_check_{comm.world_size}.json"), "w"))
comm.barrier()
comm.close()
sys.exit(0 if ok else 1)
