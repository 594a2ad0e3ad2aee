"""Committed SASS evidence: for every hot kernel, the part of its `cuobjdump -sass` listing around the Blackwell
instructions (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit,
LDGMC/STGMC... = multimem, HMMA = mma.sync) plus per-kernel mnemonic counts.  No GPU needed.
Usage: python bench/sass_extract.py  ->  profiles/sass_r2/<kernel>.sass"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "nanorlhf_b200", "csrc", "build")
OUT = os.path.join(ROOT, "profiles", "sass_r2")
os.makedirs(OUT, exist_ok=True)
# (object file, regex on the demangled kernel name, output name)
TARGETS = [
    ("gemm_tc.cu.o", r"gemm_tc_cg2_kernel<256, false, false, 0>", "gemm_tc_cg2_bn256_tn"),
    ("gemm_tc.cu.o", r"gemm_tc_cg2_kernel<192, false, true, 0>", "gemm_tc_cg2_bn192_dgrad_mnmajorB"),
    ("gemm_tc.cu.o", r"gemm_tc_cg2_kernel<256, true, true, 1>", "gemm_tc_cg2_bn256_wgrad_f32acc"),
    ("gemm_tc.cu.o", r"gemm_tc_cg2_kernel<256, false, false, 2>", "gemm_tc_cg2_bn256_swiglu"),
    ("gemm_tc.cu.o", r"gemm_tc_splitk_kernel<64, true, true>", "gemm_tc_splitk_bn64_wgrad"),
    ("gemm_tc.cu.o", r"gemm_tc_batched_cg2_kernel<256>", "gemm_tc_batched_cg2_bn256"),
    ("gemm_sm100.cu.o", r"gemm_bf16_tn_kernel<256, 1, false>", "lmhead_logprob_fused"),
    ("gemm_sm100.cu.o", r"gemm_bf16_tn_kernel<256, 2, false>", "lmhead_dlogits"),
    ("gemm_sm100.cu.o", r"gemm_bf16_tn_kernel<256, 3, false>", "lora_merge_multicast_kbc"),
    ("gemm_sm100.cu.o", r"gemm_bf16_tn_kernel<256, 0, true>", "gemm_fp8_f8f6f4"),
    ("attention_fwd_tc.cu.o", r"attn_fwd_tc_kernel", "attention_fwd_tcgen05"),
    ("attention_bwd_tc.cu.o", r"attn_bwd_dkdv_tc_kernel", "attention_bwd_dkdv_tcgen05"),
    ("attention_bwd_tc.cu.o", r"attn_bwd_dq_tc_kernel", "attention_bwd_dq_tcgen05"),
    ("attention_decode.cu.o", r"paged_decode_kernel", "paged_decode_bf16"),
    ("attention_decode_fp8.cu.o", r"paged_decode_fp8_kernel", "paged_decode_fp8_regdequant"),
    ("attention_varlen.cu.o", r"deberta_attn_fwd_tma_kernel<64>", "deberta_disentangled_attention_tma"),
    ("comm.cu.o", r"allreduce_adam_mc_kernel<float>", "kar_allreduce_adam_nvls_multimem"),
    ("comm.cu.o", r"allreduce_adam_p2p_kernel<float, 8>", "kar_allreduce_adam_p2p_8"),
    ("sampling.cu.o", r"sample_top_p_kernel<__nv_bfloat16>", "sample_top_p"),
    ("sampling.cu.o", r"sample_top_p_smem_kernel", "sample_top_p_cluster_smem"),
    ("gemm_tc.cu.o", r"gemm_tc_fp8_cg2_kernel<256, 2>", "gemm_tc_fp8_cg2_swiglu"),
    ("rope_kv.cu.o", r"rope_kv_write", "rope_kv_write_decode"),
    ("rl_kernels.cu.o", r"policy_loss_kernel", "policy_loss"),
]
KEY = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UCGABAR[A-Z_]*|SYNCS|HMMA|LDGSTS|LDSM|MULTIMEM|LDGMC|STGMC|REDG?|ATOMG|F2FP|MUFU)|STRONG\.SYS")


def listing(obj):
    raw = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    dem = subprocess.run(["cu++filt"], input=raw, capture_output=True, text=True).stdout or raw
    return dem


cache = {}
index = []
for obj, pat, name in TARGETS:
    if obj not in cache:
        cache[obj] = listing(obj)
    text = cache[obj]
    funcs = re.split(r"(?m)^\s*Function : ", text)
    def norm(t):          # "<(int)256, (bool)1>" -> "<256,true>" so the patterns can be written the C++ way
        t = re.sub(r"\((?:int|unsigned int|long)\)", "", t)
        t = t.replace("(bool)1", "true").replace("(bool)0", "false")
        return t.replace(" ", "")
    body = next((f for f in funcs[1:] if norm(pat) in norm(f.split("\n", 1)[0])), None)
    if body is None:
        index.append(f"{name}: NOT FOUND ({pat})")
        continue
    lines = body.split("\n")
    instr = [l for l in lines if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    counts = {}
    for l in instr:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            op = m.group(1).split(".")[0]
            counts[op] = counts.get(op, 0) + 1
    keep = set()
    for i, l in enumerate(instr):
        if KEY.search(l):
            keep.update(range(max(0, i - 2), min(len(instr), i + 3)))
    out = [f"// {lines[0].strip()}", f"// object: csrc/build/{obj}   instructions: {len(instr)}",
           "// mnemonic counts: " + ", ".join(f"{k}={v}" for k, v in sorted(counts.items(), key=lambda kv: -kv[1]) if KEY.search(k) or v >= 20)]
    prev = -2
    for i in sorted(keep)[:600]:
        if i != prev + 1:
            out.append("        ...")
        out.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", instr[i]).rstrip())
        prev = i
    with open(os.path.join(OUT, name + ".sass"), "w") as f:
        f.write("\n".join(out) + "\n")
    tags = [k for k in ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UCGABAR", "HMMA", "LDGSTS") if any(c.startswith(k) for c in counts)]
    mm = "multimem" if re.search(r"MULTIMEM|LDGMC|STGMC|\.MC\b|LDG\.E\..*HPADD|MMEM", "\n".join(instr)) else ""
    index.append(f"{name}: {len(instr)} instr; " + " ".join(f"{t}x{sum(v for c, v in counts.items() if c.startswith(t))}" for t in tags) + (" " + mm if mm else ""))
with open(os.path.join(OUT, "INDEX.txt"), "w") as f:
    f.write("\n".join(index) + "\n")
print("\n".join(index))
