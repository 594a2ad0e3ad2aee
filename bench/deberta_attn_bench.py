"""Fused DeBERTa disentangled attention (content + c2p + p2c) micro-benchmark at the reward-model shape."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanorlhf_b200.models.deberta_v3 import build_bucket_lut  # noqa: E402
from nanorlhf_b200.ops import native  # noqa: E402

ext = native.ext()
dev = "cuda"
out = []
for lens in ([1660] * 16, [1000] * 16, [400] * 32):
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device=dev, dtype=torch.int32)
    T = sum(lens)
    q, k, v = (torch.randn(T, 16, 64, device=dev, dtype=torch.bfloat16) for _ in range(3))
    ra, rb = (torch.randn(16, T, 512, device=dev, dtype=torch.bfloat16) for _ in range(2))
    lut = build_bucket_lut(max(lens), 256, 512, 256, dev)
    fn = lambda: ext.attn_varlen_fwd(q, k, v, cu, max(lens), 1 / math.sqrt(192), False, ra, rb, lut)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    flops = sum(4 * 16 * 64 * L * L for L in lens)
    fn2 = lambda: ext.deberta_attn_fwd(q, k, v, cu, max(lens), 1 / math.sqrt(192), ra, rb, lut)  # noqa: E731
    for _ in range(3):
        fn2()
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        fn2()
    b.record()
    torch.cuda.synchronize()
    ms2 = a.elapsed_time(b) / 10
    out.append({"lens": f"{len(lens)}x{lens[0]}", "cp_async_ms": ms, "tma_ms": ms2, "tma_content_tflops": flops / ms2 / 1e9})
    print(json.dumps(out[-1]), flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "deberta_attn_bench.json"), "w"), indent=1)
