"""Forward attention micro-benchmark: tcgen05 kernel vs the mma.sync kernel vs flash-attn (library baseline)."""
import json
import math
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))

from nanorlhf_b200.ops import native


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ext = native.ext()
    Hq, Hkv, D = 12, 2, 128
    out = []
    for lens in ([1700] * 16, [2316] * 22, [512] * 32, [4096] * 4):
        T = sum(lens)
        cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
        q = torch.randn(T, Hq, D, device="cuda").bfloat16()
        k = torch.randn(T, Hkv, D, device="cuda").bfloat16()
        v = torch.randn(T, Hkv, D, device="cuda").bfloat16()
        sc = 1.0 / math.sqrt(D)
        flops = sum(4 * Hq * D * (L * (L + 1) / 2) for L in lens)
        row = {"lens": f"{len(lens)}x{lens[0]}", "gflop": flops / 1e9}
        t = timeit(lambda: ext.attn_fwd_tc(q, k, v, cu, sc))
        row["tcgen05_ms"], row["tcgen05_tflops"] = t, flops / t / 1e9
        o, lse = ext.attn_fwd_tc(q, k, v, cu, sc)
        g = torch.randn_like(o)
        bflops = 2.5 * flops
        t = timeit(lambda: ext.attn_bwd_tc(g, q, k, v, o, lse, cu, sc))
        row["bwd_tcgen05_ms"], row["bwd_tcgen05_tflops"] = t, bflops / t / 1e9
        if __import__("os").environ.get("ATTN_ONLY") == "tc":
            print(json.dumps(row), flush=True)
            continue
        t = timeit(lambda: ext.attn_varlen_fwd(q, k, v, cu, max(lens), sc))
        row["mma_sync_ms"], row["mma_sync_tflops"] = t, flops / t / 1e9
        t = timeit(lambda: ext.attn_varlen_bwd(g, q, k, v, o, lse, cu, max(lens), sc))
        row["bwd_mma_sync_ms"], row["bwd_mma_sync_tflops"] = t, bflops / t / 1e9
        try:
            from flash_attn import flash_attn_varlen_func
            t = timeit(lambda: flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), softmax_scale=sc, causal=True))
            row["flash_attn_ms"], row["flash_attn_tflops"] = t, flops / t / 1e9
        except Exception as e:  # noqa: BLE001
            row["flash_attn"] = repr(e)[:80]
        print(json.dumps(row), flush=True)
        out.append(row)
    json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/attn_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
