"""One tiny launch of every native kernel family -- the target of bench/sanitize.sh (compute-sanitizer)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanorlhf_b200.models.deberta_v3 import build_bucket_lut  # noqa: E402
from nanorlhf_b200.ops import native  # noqa: E402

ext = native.ext()
dev = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16
# tcgen05 GEMM family
a, b = torch.randn(200, 128, device=dev, dtype=bf), torch.randn(264, 128, device=dev, dtype=bf)
native.gemm_bf16(a, b)
h = torch.randn(300, 128, device=dev, dtype=bf, requires_grad=True)
w = (torch.randn(1000, 128, device=dev) * 0.05).to(bf).requires_grad_(True)
t = torch.randint(0, 1000, (300,), device=dev)
lp, ent = native.lmhead_logprob(h, w, t, 0.9, True)
lp.sum().backward()
# tcgen05 attention forward / backward
lens = [5, 130, 300]
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
T = sum(lens)
q, k, v = torch.randn(T, 4, 128, device=dev, dtype=bf), torch.randn(T, 2, 128, device=dev, dtype=bf), torch.randn(T, 2, 128, device=dev, dtype=bf)
o, lse = ext.attn_fwd_tc(q, k, v, cu, 1 / math.sqrt(128))
ext.attn_bwd_tc(torch.randn_like(o), q, k, v, o, lse, cu, 1 / math.sqrt(128))
# TMA-fed DeBERTa attention
qd, kd, vd = (torch.randn(T, 2, 64, device=dev, dtype=bf) for _ in range(3))
ra, rb = (torch.randn(2, T, 512, device=dev, dtype=bf) for _ in range(2))
ext.deberta_attn_fwd(qd, kd, vd, cu, max(lens), 1 / math.sqrt(192), ra, rb, build_bucket_lut(max(lens), 256, 512, 256, dev))
# elementwise / RL kernels
x = torch.randn(64, 256, device=dev, dtype=bf)
native.rmsnorm(x, torch.ones(256, device=dev, dtype=bf), 1e-6)
native.add_layernorm(x, x, torch.ones(256, device=dev, dtype=bf), torch.zeros(256, device=dev, dtype=bf), 1e-6)
torch.cuda.synchronize()
print("sanitize smoke ok", native.launches(), "launches")
