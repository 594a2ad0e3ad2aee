"""Per-kernel timing vs cuBLAS on the model's GEMM shapes (CUDA events, L2 flushed between iterations).
Writes gpurun_out/gemm_bench.json.  Roofline denominators come from MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanorlhf_b200.ops import native  # noqa: E402
from nanorlhf_b200.utils.clocks import ClockSampler  # noqa: E402

native.load()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


shapes = [("qkv_decode", 2048, 2048, 1536), ("o_decode", 2048, 1536, 1536), ("gate_up_decode", 2048, 17920, 1536),
          ("down_decode", 2048, 1536, 8960), ("gate_up_train", 6912, 17920, 1536), ("down_train", 6912, 1536, 8960),
          ("gate_up_logprob", 50000, 17920, 1536), ("lm_head_decode", 2048, 151936, 1536), ("square_8k", 8192, 8192, 8192)]
rows = []
clk = ClockSampler(0).start()
for name, M, N, K in shapes:
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    row = {"name": name, "M": M, "N": N, "K": K}
    for bn in ((128, 192, 256, 0) if os.environ.get('NRL_SKIP_2CTA') == '1' else (128, 192, 256, 512, 0)):                # 512 = cta_group::2 (256x256 per CTA pair); 0 = dispatcher's choice
        key = {512: "2cta", 0: "auto"}.get(bn, f"bn{bn}")
        try:
            ms = timeit(lambda: native.ext().gemm_bf16(a, b, None, out, bn))
        except Exception as e:  # noqa: BLE001 -- a variant that cannot run this shape is recorded, not fatal
            row[f"ours_{key}_error"] = str(e)[:200]
            torch.cuda.synchronize()
            continue
        row[f"ours_{key}_ms"] = ms
        row[f"ours_{key}_tflops"] = fl / ms / 1e9
    ms = timeit(lambda: torch.matmul(a, b.t(), out=out))
    row["cublas_ms"], row["cublas_tflops"] = ms, fl / ms / 1e9
    best = row["ours_auto_tflops"]
    row["ours_frac_of_measured_peak"] = best / peaks["bf16_tflops"]
    row["ours_vs_cublas"] = best / row["cublas_tflops"]
    rows.append(row)
    print(json.dumps(row), flush=True)
    del a, b, out
# fused lm-head log-prob vs materialised logits
T, V, d = 32768, 151936, 1536
h = torch.randn(T, d, device="cuda", dtype=torch.bfloat16)
w = (torch.randn(V, d, device="cuda") * 0.02).bfloat16()
tgt = torch.randint(0, V, (T,), device="cuda", dtype=torch.int32)
ms = timeit(lambda: native.ext().lmhead_logprob_fwd(h, w, tgt, 1 / 0.9, 0), iters=5)


def eager():
    for s in range(0, T, 4096):
        z = (h[s:s + 4096] @ w.t()).float() / 0.9
        lp = torch.log_softmax(z, -1).gather(1, tgt[s:s + 4096, None].long())
    return lp


ms_e = timeit(eager, iters=3, warm=1)
row = {"name": "lmhead_logprob_fused", "T": T, "V": V, "d": d, "ours_ms": ms, "ours_tflops": 2.0 * T * V * d / ms / 1e9,
       "eager_chunked_ms": ms_e, "speedup": ms_e / ms, "ours_frac_of_measured_peak": 2.0 * T * V * d / ms / 1e9 / peaks["bf16_tflops"]}
rows.append(row)
print(json.dumps(row), flush=True)
rows.append({"clocks": clk.stop()})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "gemm_bench.json"), "w"), indent=1)
