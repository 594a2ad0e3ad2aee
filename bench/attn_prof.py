"""Cycle accounting of the tcgen05 attention forward: per-role time spent in each mbarrier wait (clock64).

Needs the instrumented build:  NRL_ATTN_PROFILE=1 python -m nanorlhf_b200.csrc.build  (production builds carry no
clock reads; the counters then stay zero).  NRL_ATTN_DBG=<bits> knocks out S MMAs (1), PV MMAs (2), exps (4), K/V loads (8)
for timing experiments -- results are wrong by construction, only the time is meaningful."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanorlhf_b200.ops import native  # noqa: E402

ext = native.ext()
lens = [4096] * 4
Hq, Hkv, D = 12, 2, 128
T = sum(lens)
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
q = torch.randn(T, Hq, D, device="cuda").bfloat16()
k = torch.randn(T, Hkv, D, device="cuda").bfloat16()
v = torch.randn(T, Hkv, D, device="cuda").bfloat16()
gx = T // 256 + len(lens)
prof = torch.zeros(Hq * gx, 40, dtype=torch.int64, device="cuda")
for _ in range(2):
    ext.attn_fwd_tc(q, k, v, cu, 1 / math.sqrt(D), prof)
torch.cuda.synchronize()
pr = prof.cpu().view(-1, 4, 10).double()
nb = pr[:, 1, 9]
live = nb > 0
names = ["kv_empty", "q_full", "kv_full", "p_ready", "s_full", "p_reuse", "rescale", "epi", "total", "n_blocks"]
for role, rn in enumerate(["producer", "mma", "softmaxA", "softmaxB"]):
    sel = pr[live][:, role]
    blocks = nb[live]
    per_block = sel[:, :9].sum(0) / blocks.sum()
    print(rn, {n: round(float(x), 1) for n, x in zip(names[:9], per_block)})
big = pr[nb == nb.max()]
print("largest CTAs (n_blocks=%d): total cycles/block by role:" % int(nb.max()), [round(float(big[:, r, 8].mean() / nb.max()), 1) for r in range(4)])
small = pr[(nb > 0) & (nb <= 8)]
print("small CTAs (<=8 blocks): total cycles by role:", [round(float(small[:, r, 8].mean()), 1) for r in range(4)], "blocks", float(nb[(nb > 0) & (nb <= 8)].mean()))
