"""gemm_tc (csrc/gemm_tc.cu) vs cuBLAS on the model's GEMM shapes, every operand-major form the trainer uses.
CUDA events, L2 flushed between iterations, median of 10, clocks recorded.  Writes gpurun_out/gemm_tc_bench.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanorlhf_b200.ops import native  # noqa: E402
from nanorlhf_b200.utils.clocks import ClockSampler  # noqa: E402

native.load()
ext = native.ext()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
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


def rnd(*shape):
    return torch.randn(*shape, device="cuda", dtype=torch.bfloat16)


TILES = [(1, 64), (1, 128), (1, 192), (1, 256), (2, 128), (2, 192), (2, 256)]
rows = []
clk = ClockSampler(0).start()

# (name, form, M, N, K): TN = x W^T, NN = dy W (dgrad), NT = dy^T x (wgrad)
shapes = [("qkv_decode", "TN", 1024, 2048, 1536), ("o_decode", "TN", 1024, 1536, 1536), ("gate_up_decode", "TN", 1024, 17920, 1536),
          ("down_decode", "TN", 1024, 1536, 8960), ("down_decode_2k", "TN", 2048, 1536, 8960), ("lm_head_decode", "TN", 1024, 151936, 1536),
          ("q_train", "TN", 6912, 1536, 1536), ("gate_train", "TN", 6912, 8960, 1536), ("down_train", "TN", 6912, 1536, 8960),
          ("gate_logprob", "TN", 50000, 8960, 1536), ("down_logprob", "TN", 50000, 1536, 8960), ("square_8k", "TN", 8192, 8192, 8192),
          ("deberta_ffn1", "TN", 26560, 4096, 1024), ("deberta_ffn2", "TN", 26560, 1024, 4096), ("deberta_qkv", "TN", 26560, 1024, 1024),
          ("q_dgrad", "NN", 6912, 1536, 1536), ("gate_dgrad", "NN", 6912, 1536, 8960), ("down_dgrad", "NN", 6912, 8960, 1536),
          ("lmhead_dH", "NN", 8192, 1536, 151936),
          ("loraA_wgrad", "NT", 64, 1536, 6912), ("loraB_wgrad", "NT", 1536, 64, 6912), ("loraB_gate_wgrad", "NT", 8960, 64, 6912),
          ("lmhead_dW", "NT", 151936, 1536, 8192)]
for name, form, M, N, K in shapes:
    a_mn, b_mn = form == "NT", form in ("NN", "NT")
    a = rnd(K, M) if a_mn else rnd(M, K)
    b = rnd(K, N) if b_mn else rnd(N, K)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    row = {"name": name, "form": form, "M": M, "N": N, "K": K}
    best = 0.0
    for cg, bn in TILES + [(0, 0)]:
        key = f"cg{cg}_bn{bn}" if cg else "auto"
        try:
            ms = timeit(lambda: ext.gemm_tc(a, b, a_mn, b_mn, None, None, None, 0, 1.0, out, None, False, cg, bn))
        except Exception as e:  # noqa: BLE001 -- not every tile shape is instantiated for every form
            torch.cuda.synchronize()
            continue
        row[f"{key}_tflops"] = round(fl / ms / 1e9, 1)
        if cg:
            best = max(best, fl / ms / 1e9)
    am = a.t() if a_mn else a
    bm = b if b_mn else b.t()
    ms = timeit(lambda: torch.matmul(am, bm, out=out))
    row["cublas_tflops"] = round(fl / ms / 1e9, 1)
    row["best_vs_cublas"] = round(best / row["cublas_tflops"], 3)
    row["auto_vs_cublas"] = round(row.get("auto_tflops", 0) / row["cublas_tflops"], 3)
    row["auto_frac_of_measured_peak"] = round(row.get("auto_tflops", 0) / peaks["bf16_tflops"], 3)
    rows.append(row)
    print(json.dumps(row), flush=True)
    del a, b, out

# LoRA forward: one dual-source-K GEMM (+ the rank-64 projection) vs cuBLAS base GEMM + 2 adapter GEMMs (addmm_)
for name, M, N, K in [("lora_q_fwd", 6912, 1536, 1536), ("lora_gate_fwd", 6912, 8960, 1536), ("lora_down_fwd", 6912, 1536, 8960)]:
    x, w, A, B = rnd(M, K), rnd(N, K), rnd(64, K), rnd(N, 64)

    def ours():
        t = ext.gemm_tc(x, A, False, False, None, None, None, 0, 0.25)
        return ext.gemm_tc(x, w, False, False, t, B)

    def cublas():
        t = x @ A.t()
        y = x @ w.t()
        return y.addmm_(t, B.t(), alpha=0.25)

    o, c = timeit(ours), timeit(cublas)
    row = {"name": name, "M": M, "N": N, "K": K, "ours_ms": o, "cublas_3gemm_ms": c, "speedup": round(c / o, 3)}
    rows.append(row)
    print(json.dumps(row), flush=True)
rows.append({"clocks": clk.stop()})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "gemm_tc_bench.json"), "w"), indent=1)
