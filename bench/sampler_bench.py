"""Sampling kernels (streaming vs shared-memory-resident cluster) and small-batch decode GEMMs (single pass vs split-K)
at the benchmark's shapes.  CUDA events after warm-up; inputs larger than L2 (311 MB of logits per launch).
Writes gpurun_out/sampler_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanorlhf_b200.ops import native  # noqa: E402

native.load()
ext = native.ext()
dev = "cuda"


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3          # us


res = {"sampler_us": {}, "gemm_us": {}}
V = 151936
for S in (256, 1024, 2048):
    for name, std in (("flat_random_init", 0.05), ("peaked", 4.0)):
        lg = (torch.randn(S, V, device=dev) * std).bfloat16()
        rid = torch.arange(S, device=dev, dtype=torch.int32)
        row = {}
        for impl, label in ((1, "stream"), (2, "cluster_smem")):
            row[label] = timeit(lambda: ext.sample(lg, 0.9, 0.95, 1, 0, rid, rid, None, impl))
        row["hbm_floor_us"] = S * V * 2 / 6.58e12 * 1e6
        res["sampler_us"][f"S{S}_{name}"] = row
        del lg
for (M, N, K, with_bias, what) in ((64, 4608, 3584, True, "7B qkv"), (64, 3584, 3584, False, "7B o_proj"), (64, 3584, 18944, False, "7B down"),
                                   (256, 3584, 18944, False, "7B down S=256"), (256, 4608, 3584, True, "7B qkv S=256"),
                                   (64, 1536, 8960, False, "1.5B down S=64"), (1024, 1536, 8960, False, "1.5B down S=1024"),
                                   (1024, 1536, 1536, False, "1.5B o_proj S=1024"), (6912, 64, 1536, False, "LoRA t = x A^T")):
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16() if with_bias else None
    row = {"auto": timeit(lambda: ext.gemm_tc(a, b, False, False, None, None, bias)),
           "single_pass": timeit(lambda: ext.gemm_tc(a, b, False, False, None, None, bias, 0, 1.0, None, None, False, 0, 0, 0)),
           "weight_stream_floor_us": N * K * 2 / 6.58e12 * 1e6}
    res["gemm_us"][f"{what} [{M}x{N}x{K}]"] = row
# LoRA adapter-input gradient dA = t'^T x  ([r, T] x [T, K_in], both MN-major)
for (r, T, Kin) in ((64, 6912, 1536), (192, 6912, 1536), (64, 6912, 8960)):
    tp = torch.randn(T, r, device=dev).bfloat16()
    x = torch.randn(T, Kin, device=dev).bfloat16()
    res["gemm_us"][f"LoRA dA [{r}x{Kin}x{T}]"] = {
        "auto": timeit(lambda: ext.gemm_tc(tp, x, True, True)),
        "single_pass": timeit(lambda: ext.gemm_tc(tp, x, True, True, None, None, None, 0, 1.0, None, None, False, 0, 0, 0))}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/sampler_bench.json", "w"), indent=1)
print(json.dumps(res, indent=1))
