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
        for impl, label in ((1, "stream_histogram"), (2, "cluster_smem"), (24, "cluster_smem_4cta")):
            row[label] = timeit(lambda: ext.sample(lg, 0.9, 0.95, 1, 0, rid, rid, None, impl))
        row["hbm_floor_us"] = S * V * 2 / 6.58e12 * 1e6
        res["sampler_us"][f"S{S}_{name}"] = row
        del lg
def sweep(a, b, a_mn, b_mn, bias):
    row = {}
    for label, sk in (("single_pass", 0), ("auto", -1), ("split2", 2), ("split4", 4), ("split8", 8), ("split12", 12)):
        row[label] = timeit(lambda: ext.gemm_tc(a, b, a_mn, b_mn, None, None, bias, 0, 1.0, None, None, False, 0, 0, sk))
    return row


for (M, N, K, with_bias, what) in ((64, 4608, 3584, True, "7B qkv"), (64, 3584, 3584, False, "7B o_proj"), (64, 3584, 18944, False, "7B down"),
                                   (256, 3584, 18944, False, "7B down S=256"), (256, 4608, 3584, True, "7B qkv S=256"),
                                   (64, 1536, 8960, False, "1.5B down S=64"), (1024, 1536, 8960, False, "1.5B down S=1024"),
                                   (1024, 1536, 1536, False, "1.5B o_proj S=1024"), (6912, 64, 1536, False, "LoRA t = x A^T"),
                                   (6912, 192, 1536, False, "LoRA t (q,k,v grouped)"), (6912, 64, 8960, False, "LoRA t (down)")):
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16() if with_bias else None
    row = sweep(a, b, False, False, bias)
    row["weight_stream_floor_us"] = N * K * 2 / 6.58e12 * 1e6
    res["gemm_us"][f"{what} [{M}x{N}x{K}]"] = row
# LoRA weight gradients: dA = t'^T x  ([r, T] x [T, K_in]) and dB = g^T t ([N_out, T] x [T, r]), both operands MN-major
for (Mo, No, T, what) in ((64, 1536, 6912, "dA"), (192, 1536, 6912, "dA q,k,v grouped"), (64, 8960, 6912, "dA down"),
                          (1536, 64, 6912, "dB q/o/down"), (256, 64, 6912, "dB k/v"), (8960, 64, 6912, "dB gate/up")):
    a = torch.randn(T, Mo, device=dev).bfloat16()
    b = torch.randn(T, No, device=dev).bfloat16()
    res["gemm_us"][f"LoRA {what} [{Mo}x{No}x{T}]"] = sweep(a, b, True, True, None)
# t' = s dy B  ([T, N_out] x [N_out, r], B MN-major)
for (T, r, Nout) in ((6912, 64, 1536), (6912, 64, 8960)):
    a = torch.randn(T, Nout, device=dev).bfloat16()
    b = torch.randn(Nout, r, device=dev).bfloat16()
    res["gemm_us"][f"LoRA t' [{T}x{r}x{Nout}]"] = sweep(a, b, False, True, None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/sampler_bench.json", "w"), indent=1)
print(json.dumps(res))
