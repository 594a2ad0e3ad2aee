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
# ---- general tcgen05 GEMM (gemm_tc.cu): every operand-major form, both CTA-group sizes, dual-source K, fp32 accumulate,
#      split-K, batched 3D TMA, SwiGLU epilogue ----
A, B = torch.randn(300, 200, device=dev, dtype=bf), torch.randn(264, 200, device=dev, dtype=bf)
for cg, bn in ((1, 64), (1, 192), (2, 128), (2, 192), (2, 256)):
    ext.gemm_tc(A, B, False, False, None, None, None, 0, 1.0, None, None, False, cg, bn, 0)
Bm = torch.randn(200, 264, device=dev, dtype=bf)
for cg, bn in ((1, 128), (2, 192), (2, 256)):
    ext.gemm_tc(A, Bm, False, True, None, None, None, 0, 1.0, None, None, False, cg, bn, 0)
Am = torch.randn(200, 304, device=dev, dtype=bf)
for cg, bn in ((1, 64), (2, 256)):
    ext.gemm_tc(Am, Bm, True, True, None, None, None, 0, 1.0, None, None, False, cg, bn, 0)
acc = torch.zeros(304, 264, device=dev)
ext.gemm_tc(Am, Bm, True, True, None, None, None, 0, 1.0, None, acc, True, 2, 256, 0)
A2, B2 = torch.randn(300, 64, device=dev, dtype=bf), torch.randn(264, 64, device=dev, dtype=bf)
ext.gemm_tc(A, B, False, False, A2, B2, None, 0, 1.0, None, None, False, 2, 256, 0)
ext.gemm_tc(torch.randn(2000, 64, device=dev, dtype=bf), torch.randn(2000, 128, device=dev, dtype=bf), True, True, None, None, None, 0, 1.0, None, None, False, 0, 0, 5)
ext.gemm_tc_batched(torch.randn(300, 4, 64, device=dev, dtype=bf).transpose(0, 1), torch.randn(136, 4, 64, device=dev, dtype=bf).transpose(0, 1), None, 2, 256)
ext.gemm_tc_swiglu(A, torch.randn(256, 200, device=dev, dtype=bf), None, 2, 256)
# 1-CTA family epilogues still in use: fused SwiGLU (fp8 path), LoRA merge (K-BC), fp8 GEMM + row quantiser
wl, la, lb = torch.randn(264, 128, device=dev, dtype=bf), torch.randn(16, 128, device=dev, dtype=bf), torch.randn(264, 16, device=dev, dtype=bf)
ext.lora_merge(wl, la, lb, 0.25, torch.empty_like(wl))
xq, xs = ext.quant_rows_e4m3(torch.randn(200, 128, device=dev, dtype=bf))
wq, ws = ext.quant_rows_e4m3(torch.randn(264, 128, device=dev, dtype=bf))
ext.gemm_fp8(xq, xs, wq, ws, None, False)
# ---- elementwise fwd + bwd: rope, swiglu ----
pos = torch.arange(T, device=dev)
from nanorlhf_b200.ops import reference as ref  # noqa: E402
cos, sin = ref.rope_cos_sin(pos, 128, 1e6)
qg = q.clone().requires_grad_(True)
native.apply_rope(qg, cos, sin).sum().backward()
gu = torch.randn(64, 512, device=dev, dtype=bf, requires_grad=True)
native.swiglu(gu).sum().backward()
g1, u1 = (torch.randn(64, 256, device=dev, dtype=bf, requires_grad=True) for _ in range(2))
native.swiglu_pair(g1, u1).sum().backward()
xr = torch.randn(64, 256, device=dev, dtype=bf, requires_grad=True)
native.rmsnorm(xr, torch.ones(256, device=dev, dtype=bf), 1e-6).sum().backward()
# ---- RL kernels: GAE scan, policy loss, value loss, flat AdamW (with and without fp32 master) ----
rew, val = torch.randn(8, 40, device=dev), torch.randn(8, 40, device=dev)
native.gae_scan(rew, val, 1.0, 0.95)
nl = torch.randn(8, 40, device=dev, requires_grad=True)
m8 = torch.rand(8, 40, device=dev) > 0.2
loss, _ = native.policy_loss_token(nl, torch.randn(8, 40, device=dev), torch.randn(8, 40, device=dev), m8, 0.2, torch.randn(8, 40, device=dev), 0.01)
loss.backward()
vp = torch.randn(8, 40, device=dev, requires_grad=True)
native.value_loss(vp, val, rew, m8, 0.2)[0].backward()
P, G = torch.randn(4096, device=dev).to(bf), torch.randn(4096, device=dev).to(bf)
native.adamw_flat(P, G, torch.zeros(4096, device=dev), torch.zeros(4096, device=dev), 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0, P.float())
# ---- sampler kernels: KV page writers, paged decode (bf16 + fp8 register-dequant), top-p sampler, arg-max ----
S, Hq, Hkv, nblk = 5, 4, 2, 40
ctx = torch.tensor([1, 16, 17, 100, 130], device=dev, dtype=torch.int32)
table = torch.arange(S * 9, device=dev, dtype=torch.int32).view(S, 9)
kc, vc = (torch.zeros(nblk + 8, Hkv, 16, 128, device=dev, dtype=bf) for _ in range(2))
kk, vv = torch.randn(nblk * 16, Hkv, 128, device=dev, dtype=bf), torch.randn(nblk * 16, Hkv, 128, device=dev, dtype=bf)
slots = torch.arange(nblk * 16, device=dev, dtype=torch.int32)
native.kv_cache_write(kk, vv, kc, vc, slots)
qd1 = torch.randn(S, Hq, 128, device=dev, dtype=bf)
native.paged_decode(qd1, kc, vc, table, ctx, splits=1)
native.paged_decode(qd1, kc, vc, table, ctx, splits=2)
kq, vq = (torch.zeros(nblk + 8, Hkv, 16, 128, device=dev, dtype=torch.uint8) for _ in range(2))
ks8, vs8 = (torch.ones(nblk + 8, Hkv, 16, device=dev) for _ in range(2))
ext.kv_cache_write_fp8(kk, vv, kq, vq, ks8, vs8, slots, None)
ext.paged_decode_fp8(qd1, kq, vq, ks8, vs8, table, ctx, 1 / math.sqrt(128), 1)
ext.paged_decode_fp8(qd1, kq, vq, ks8, vs8, table, ctx, 1 / math.sqrt(128), 2)
lg = torch.randn(6, 5000, device=dev, dtype=bf)
native.sample(lg, 0.9, 0.95, 7, 0, impl=1)               # streaming histogram kernel
native.sample(lg, 0.0, 1.0, 7, 0)
lgw = torch.randn(6, 151936, device=dev, dtype=bf)
for impl in (21, 22, 24):                                # cluster kernel: 1 / 2 / 4 CTAs per row (bulk loads, DSMEM exchange)
    native.sample(lg if impl == 21 else lgw, 0.9, 0.95, 7, 0, impl=impl)
native.sample(lgw, 1.0, 0.5, 7, 0, impl=2)               # many redraws
# split-K with bias in the fix-up, and the fused decode-step RoPE + KV page write
ext.gemm_tc(torch.randn(64, 2048, device=dev, dtype=bf), torch.randn(264, 2048, device=dev, dtype=bf), False, False, None, None,
            torch.randn(264, device=dev, dtype=bf))
qkv = torch.randn(5, (Hq + 2 * Hkv) * 128, device=dev, dtype=bf)
cs, sn = ref.rope_cos_sin(torch.arange(5, device=dev), 128, 1e6)
sl5 = torch.arange(5, device=dev, dtype=torch.int32) * 7
ext.rope_kv_write(qkv.clone(), cs, sn, kc, vc, None, None, sl5, Hq, Hkv)
ext.rope_kv_write(qkv.clone(), cs, sn, kq, vq, ks8, vs8, sl5, Hq, Hkv)
torch.cuda.synchronize()
print("sanitize smoke ok", native.launches(), "launches")
