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
comm.broadcast_module_(policy, 0)          # replicas start identical (the trainer does the same for every model it holds)
local = NativeSampler(policy)
refresh_sampler_arena(local)
want = [(lw.wqkv.clone(), lw.wo.clone(), lw.wgu.clone(), lw.wdown.clone()) for lw in local.layers]
shared = NativeSampler(policy)
sync = ShardedWeightSync(shared, comm)


def mismatches():
    """{(layer, matrix): number of differing elements} on this rank."""
    bad = {}
    for li, (lw, w) in enumerate(zip(shared.layers, want)):
        for name, a, b in zip(("wqkv", "wo", "wgu", "wdown"), (lw.wqkv, lw.wo, lw.wgu, lw.wdown), w):
            n = int((a != b).sum())
            if n:
                bad[f"L{li}.{name}"] = n
    return bad


# replicas must hold identical inputs: checksum of every parameter, compared across ranks
chk = torch.stack([p.detach().float().sum() for p in policy.parameters()]).double().sum().reshape(1)
allchk = comm.all_gather_cat(chk)
inputs_identical = bool((allchk == allchk[0]).all())
sync.refresh()
torch.cuda.synchronize()
bad_first = mismatches()
import time as _time
_time.sleep(0.2)
torch.cuda.synchronize()
bad_after_wait = mismatches()
sync.refresh()
torch.cuda.synchronize()
bad_second = mismatches()
# staleness probes: change the adapters, refresh, compare -- once as shipped, once draining the stream before the closing barrier,
# once through plain peer stores (no multicast)
probes = {}
for tag in ("multicast", "multicast_drained", "peer_stores"):
    with torch.no_grad():
        for m in policy.modules():
            if isinstance(m, LoraLinear):
                m.lora_B.weight.mul_(1.25)
    refresh_sampler_arena(local)
    want = [(lw.wqkv.clone(), lw.wo.clone(), lw.wgu.clone(), lw.wdown.clone()) for lw in local.layers]
    sync.drain_before_barrier = tag == "multicast_drained"
    saved = sync.mc_base
    if tag == "peer_stores":
        sync.mc_base = 0
    sync.refresh()
    torch.cuda.synchronize()
    probes[tag] = len(mismatches())
    sync.mc_base = saved
sync.drain_before_barrier = False
ok = not bad_second and not bad_first and probes["multicast"] == 0
diag = {"probes_mismatching_matrices": probes, "rank": comm.rank, "inputs_identical_across_ranks": inputs_identical, "first_refresh_mismatch": dict(list(bad_first.items())[:6]),
        "n_first": len(bad_first), "n_after_200ms": len(bad_after_wait), "n_second_refresh": len(bad_second)}
if bad_first or bad_second or not inputs_identical or any(probes.values()):
    print("[wsync diag]", json.dumps(diag), file=sys.stderr, flush=True)


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


t_local, t_sharded = timed(lambda: refresh_sampler_arena(local)), timed(sync.refresh)
arena_gb = sync.flat.numel() * 2 / 1e9
# K-BC roofline: every rank must RECEIVE (world-1)/world of the merged arena over NVLink (770 GB/s/dir measured) and read
# 1/world of the base weights from HBM; the slower of the two bounds the refresh
W = comm.world_size
floor_ms = max(arena_gb * (W - 1) / W / 770.0, arena_gb / W / 6583.0) * 1e3
verdict = torch.tensor([1 if ok else 0], device=dev)
comm.all_reduce_(verdict, "min")
ok = bool(verdict.item())
res = {"world": W, "identical_to_local_merge": ok, "inputs_identical_across_ranks": inputs_identical, "rank0_diag": diag, "multicast": sync.stats["multicast"], "local_merge_ms": t_local,
       "sharded_multicast_ms": t_sharded, "arena_gb": arena_gb, "nvlink_floor_ms": floor_ms,
       "frac_of_nvlink_roofline": floor_ms / t_sharded}
if comm.is_main:
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"wsync_check_{comm.world_size}.json"), "w"))
del sync, shared, local
torch.cuda.synchronize()
comm.barrier()
comm.close()
sys.stdout.flush()
os._exit(0 if ok else 1)
