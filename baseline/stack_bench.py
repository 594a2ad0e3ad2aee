"""Same-box anchors: the *reference's dependency stack* (vLLM + HF transformers/flash-attn-2 + cuBLAS + NCCL +
torch AdamW) timed phase by phase at the shapes of the headline benchmark -- with ZERO code from this repository.

Why this exists: the reference itself (``/root/reference/GRPO/grpo.py``) cannot run on the benchmark box -- it is not
an installable package, imports trl / peft / accelerate (absent from the image and the offline wheelhouse) and pulls
its models and datasets from the HF hub (no network); ``bench.py --impl reference`` therefore reports ``unavailable``
(DESIGN.md section 7).  What *can* be measured honestly is every library call the reference's update is made of
(/root/reference/GRPO/grpo_trainer.py): each phase below names the reference lines it stands for and is deliberately
CHARITABLE to the reference -- no CPU<->GPU model swaps, no disk round trip of the weights, no LoRA adapter GEMMs,
no ``empty_cache()`` storms -- so the resulting per-phase times are lower bounds of what the reference would need.

  python baseline/stack_bench.py [--prompts 256] [--phases rollout,logprob,train,reward,optim]
  torchrun --nproc-per-node N baseline/stack_bench.py --phases optim        (all-reduce + AdamW at N GPUs)

Each phase runs in its own subprocess (vLLM keeps the GPU memory it reserved) and appends to
``gpurun_out/stack_bench.json``.  All timings: CUDA events (or wall clock around blocking library calls that own
their streams, i.e. ``LLM.generate``), >= 3 warm-up iterations where an iteration is cheap, nvidia-smi clocks recorded.
Weights are random-init of the named architectures (no network); inputs are synthetic ids of the benchmark's shapes.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "stack_bench.json")
WORK = os.environ.get("STACK_BENCH_DIR", "/tmp/stack_bench")

QWEN_15B = dict(vocab_size=151936, hidden_size=1536, intermediate_size=8960, num_hidden_layers=28, num_attention_heads=12,
                num_key_value_heads=2, max_position_embeddings=32768, rms_norm_eps=1e-6, rope_theta=1000000.0,
                tie_word_embeddings=True, bos_token_id=151643, eos_token_id=151645, pad_token_id=151643)
DEBERTA_LARGE = dict(vocab_size=128100, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                     max_position_embeddings=512, relative_attention=True, position_buckets=256, norm_rel_ebd="layer_norm",
                     share_att_key=True, pos_att_type=["p2c", "c2p"], position_biased_input=False, max_relative_positions=-1,
                     type_vocab_size=0, layer_norm_eps=1e-7, pooler_hidden_size=1024, pooler_dropout=0, pooler_hidden_act="gelu",
                     num_labels=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


# ---------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, idx=0):
        self.rows, self.idx, self.proc = [], idx, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx),
                                          "-lms", "500"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.rows.append([c.strip() for c in ln.split(",")]) for ln in self.proc.stdout],
                             daemon=True).start()
        except Exception:
            pass
        return self

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        ok = [r for r in self.rows if len(r) >= 9]
        sm = [float(r[1]) for r in ok if r[1].replace(".", "").isdigit()]
        reasons = sorted({n for r in ok for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9])
                          if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max([float(r[2]) for r in ok if r[2].replace(".", "").isdigit()] or [0]),
                "reasons": reasons, "samples": len(sm)}


def record(key, value):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    data[key] = value
    json.dump(data, open(OUT, "w"), indent=1)
    print(json.dumps({key: value}), flush=True)


def model_dir():
    """Random-init Qwen2.5-1.5B in HF format (written once with transformers; no repo code)."""
    d = os.path.join(WORK, "qwen2.5-1.5b-random")
    if os.path.exists(os.path.join(d, "model.safetensors")):
        return d
    import torch
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(0)
    cfg = Qwen2Config(**QWEN_15B)
    cfg.torch_dtype = torch.bfloat16
    m = Qwen2ForCausalLM(cfg).to(torch.bfloat16)
    os.makedirs(d, exist_ok=True)
    m.save_pretrained(d, safe_serialization=True)
    with open(os.path.join(d, "generation_config.json"), "w") as f:
        json.dump({"bos_token_id": 151643, "eos_token_id": [151645], "pad_token_id": 151643}, f)
    return d


def synthetic_prompts(n, seed=1, lo=24, hi=160, vocab=151000):
    import random
    rng = random.Random(seed)
    return [[rng.randrange(5, vocab) for _ in range(rng.randint(lo, hi))] for _ in range(n)]


def cuda_time(fn, iters=3, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


# ---------------------------------------------------------------------------------------------------------
# phase: rollout  (reference: vllm_generate, /root/reference/GRPO/grpo_trainer.py:122-166)
# ---------------------------------------------------------------------------------------------------------
def phase_rollout(args):
    path = model_dir()
    clocks = Clocks().start()
    from vllm import LLM, SamplingParams
    t0 = time.perf_counter()
    kw = dict(model=path, dtype="bfloat16", skip_tokenizer_init=True, max_model_len=2048, gpu_memory_utilization=0.85,
              enable_prefix_caching=True, seed=0)
    if args.max_num_seqs:
        kw["max_num_seqs"] = args.max_num_seqs
        kw["max_num_batched_tokens"] = max(8192, args.max_num_seqs)
    llm = LLM(**kw)
    boot_s = time.perf_counter() - t0
    prompts = [{"prompt_token_ids": p} for p in synthetic_prompts(args.prompts)]
    sp = SamplingParams(temperature=0.9, top_p=0.95, n=args.samples, max_tokens=args.response_length, logprobs=1, detokenize=False,
                        ignore_eos=False, seed=1234)
    # one small warm-up generate (captures graphs / compiles), then the timed full-size call
    llm.generate(prompts[:8], SamplingParams(temperature=0.9, top_p=0.95, n=args.samples, max_tokens=16, detokenize=False), use_tqdm=False)
    times, toks = [], 0
    for it in range(args.rollout_iters):
        sp.seed = 1234 + it
        t1 = time.perf_counter()
        outs = llm.generate(prompts, sp, use_tqdm=False)
        times.append(time.perf_counter() - t1)
        toks = sum(len(o.token_ids) for r in outs for o in r.outputs)
    gen_s = min(times)
    record(f"rollout_vllm{'_mns' + str(args.max_num_seqs) if args.max_num_seqs else '_stock'}", {
        "what": "vLLM LLM.generate, Qwen2.5-1.5B random init, bf16, T=0.9 top_p=0.95, prompt_token_ids 24-160, "
                f"{args.prompts} prompts x n={args.samples}, max_tokens={args.response_length}",
        "engine_boot_s": boot_s, "generate_s": gen_s, "generate_s_all": times, "generated_tokens": toks,
        "tok_per_s": toks / gen_s, "vllm_kwargs": {k: v for k, v in kw.items() if k != "model"}, "clocks": clocks.stop(),
        "ref_lines": "GRPO/grpo_trainer.py:122-166 (the reference additionally saves + merges on CPU + boots the engine every update)"})


# ---------------------------------------------------------------------------------------------------------
# phases: logprob / train  (reference: grpo_trainer.py:538-556 and :630-693)
# ---------------------------------------------------------------------------------------------------------
def _hf_policy(train: bool):
    import torch
    from transformers import AutoModelForCausalLM
    m = AutoModelForCausalLM.from_pretrained(model_dir(), torch_dtype=torch.bfloat16, attn_implementation="flash_attention_2").cuda()
    m.train(train)
    return m


def _batch(bs, ctx, resp, vocab=151000, pad=151643, seed=0):
    """Left-padded queries (24..160 real tokens) + full-length responses, like the benchmark's rollouts."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(5, vocab, (bs, ctx + resp), generator=g)
    qlen = torch.randint(24, ctx + 1, (bs,), generator=g)
    col = torch.arange(ctx)[None, :]
    ids[:, :ctx] = torch.where(col >= (ctx - qlen)[:, None], ids[:, :ctx], torch.full_like(ids[:, :ctx], pad))
    return ids.cuda(), pad


def _ref_forward(model, qr, pad):
    """The reference's ``forward`` helper (grpo_trainer.py:90-120), re-typed from its description."""
    import torch
    mask = qr != pad
    pos = mask.cumsum(1) - mask.long()
    ids = torch.masked_fill(qr, ~mask, 0)
    return model(input_ids=ids, attention_mask=mask, position_ids=pos, return_dict=False, use_cache=False)[0]


def phase_logprob(args):
    import torch
    clocks = Clocks().start()
    m = _hf_policy(False)
    ctx, resp = 160, args.response_length
    bs = 22 * 2316 // (ctx + resp)                       # the reference's chunk size (grpo_trainer.py:534)
    qr, pad = _batch(bs, ctx, resp)

    @torch.no_grad()
    def one():
        logits = _ref_forward(m, qr, pad)[:, ctx - 1:-1]
        logits = logits.clone()
        logits /= 0.9
        lp = torch.nn.functional.log_softmax(logits, dim=-1)
        return torch.gather(lp, 2, qr[:, ctx:].unsqueeze(-1)).squeeze(-1)

    ms = cuda_time(one, iters=3, warm=2)
    real_tokens = int((qr != pad).sum())
    record("logprob_hf_fa2", {"what": f"HF Qwen2 + flash_attention_2 no-grad forward + logits/T + log_softmax + gather, batch {bs} x {ctx + resp}",
                              "ms_per_chunk": ms, "sequences_per_chunk": bs, "real_tokens_per_chunk": real_tokens,
                              "ms_per_sequence": ms / bs, "clocks": clocks.stop(),
                              "ref_lines": "GRPO/grpo_trainer.py:538-556 (run twice per chunk: policy with unmerged LoRA, then ref)"})


def phase_train(args):
    import torch
    clocks = Clocks().start()
    res = {}
    for ckpt in (True, False):
        m = _hf_policy(True)
        for p in m.parameters():
            p.requires_grad_(False)
        # modules_to_save = [embed_tokens, lm_head]: full trainable copies (peft un-ties them); LoRA adapters are NOT modelled
        m.get_input_embeddings().weight.requires_grad_(True)
        if ckpt:
            m.gradient_checkpointing_enable()
            m.enable_input_require_grads()
        ctx, resp, bs = 160, args.response_length, 4
        qr, pad = _batch(bs, ctx, resp)
        old = torch.randn(bs, resp, device="cuda") * 0.01 - 11.0
        adv = torch.randn(bs, 1, device="cuda").expand(bs, resp)

        def one():
            logits = _ref_forward(m, qr, pad)[:, ctx - 1:-1]
            logits = logits / 0.9
            lp = torch.gather(torch.nn.functional.log_softmax(logits, dim=-1), 2, qr[:, ctx:].unsqueeze(-1)).squeeze(-1)
            ratio = torch.exp(lp - old)
            loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2)).mean()
            with torch.no_grad():                       # the reference's entropy statistic: a second softmax + logsumexp over V
                pd = torch.nn.functional.softmax(logits, dim=-1)
                _ = (torch.logsumexp(logits, dim=-1) - torch.sum(pd * logits, dim=-1)).mean()
            (loss / 8).backward()

        try:
            ms = cuda_time(one, iters=3, warm=2)
            res["ckpt_on" if ckpt else "ckpt_off"] = ms
        except Exception as e:  # noqa: BLE001
            res["ckpt_on_error" if ckpt else "ckpt_off_error"] = repr(e)[:300]
        del m
        torch.cuda.empty_cache()
    record("train_hf_fa2", {"what": "HF Qwen2 + flash_attention_2 micro-step fwd+loss+bwd, 4 x 1660 tokens, frozen base, trainable embed/lm_head, "
                                    "NO LoRA adapters (peft unavailable) -- gradient checkpointing on = the reference's setting",
                            "ms_per_micro_step": res, "clocks": clocks.stop(), "ref_lines": "GRPO/grpo_trainer.py:630-693"})


# ---------------------------------------------------------------------------------------------------------
# phase: reward  (reference: GRPO/grpo.py:162-198)
# ---------------------------------------------------------------------------------------------------------
def phase_reward(args):
    import torch
    from transformers import DebertaV2Config, DebertaV2ForSequenceClassification
    clocks = Clocks().start()
    torch.manual_seed(0)
    res = {}
    L = 24 + 92 + args.response_length                     # [CLS] q [SEP] r [SEP] at the benchmark's mean prompt length
    for dt in (torch.float32, torch.bfloat16):
        m = DebertaV2ForSequenceClassification(DebertaV2Config(**DEBERTA_LARGE)).to(dt).cuda().eval()
        ids = torch.randint(5, 128000, (16, L), device="cuda")
        mask = torch.ones_like(ids)

        @torch.no_grad()
        def one():
            return m(input_ids=ids, attention_mask=mask).logits.squeeze()

        try:
            res[str(dt).split(".")[-1]] = cuda_time(one, iters=3, warm=2)
        except Exception as e:  # noqa: BLE001
            res[str(dt).split(".")[-1] + "_error"] = repr(e)[:300]
        del m
        torch.cuda.empty_cache()
    record("reward_hf_deberta", {"what": f"HF DebertaV2ForSequenceClassification (deberta-v3-large shape, random init) forward, 16 x {L}",
                                 "ms_per_batch_of_16": res, "clocks": clocks.stop(),
                                 "ref_lines": "GRPO/grpo.py:162-198 (fp32 is the reference's dtype; it also swaps the RM CPU<->GPU each call)"})


# ---------------------------------------------------------------------------------------------------------
# phase: optim  (reference: DDP all-reduce + torch AdamW, grpo_trainer.py:690-693)
# ---------------------------------------------------------------------------------------------------------
def phase_optim(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    clocks = Clocks(local).start()
    n = 540_672_000
    p = torch.zeros(n, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    p.grad = torch.randn(n, device="cuda").bfloat16()
    opt = torch.optim.AdamW([p], lr=6e-6, fused=True)

    def one():
        if world > 1:
            dist.all_reduce(p.grad)
            p.grad.div_(world)
        opt.step()

    ms = cuda_time(one, iters=10, warm=3)
    ar = cuda_time(lambda: dist.all_reduce(p.grad), iters=10, warm=3) if world > 1 else 0.0
    t = torch.tensor([ms, ar], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        record(f"optim_nccl_adamw_{world}gpu", {"what": "dist.all_reduce (NCCL) of 540.7M bf16 grads + torch.optim.AdamW(fused=True) step",
                                                 "ms_allreduce_plus_adamw": float(t[0]), "ms_allreduce_only": float(t[1]), "clocks": clocks.stop(),
                                                 "ref_lines": "GRPO/grpo_trainer.py:690-693"})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


PHASES = {"rollout": phase_rollout, "logprob": phase_logprob, "train": phase_train, "reward": phase_reward, "optim": phase_optim}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phases", default="rollout,logprob,train,reward,optim")
    ap.add_argument("--phase", default=None, help="(internal) run one phase in this process")
    ap.add_argument("--prompts", type=int, default=256)
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--response-length", type=int, default=1500)
    ap.add_argument("--max-num-seqs", type=int, default=0, help="0 = vLLM stock default (what the reference uses)")
    ap.add_argument("--rollout-iters", type=int, default=2)
    ap.add_argument("--phase-timeout", type=int, default=900)
    args = ap.parse_args()
    if args.phase:
        PHASES[args.phase](args)
        return 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:          # under torchrun only the collective phase makes sense
        phase_optim(args)
        return 0
    for ph in args.phases.split(","):
        cmd = [sys.executable, os.path.abspath(__file__), "--phase", ph, "--prompts", str(args.prompts), "--samples", str(args.samples),
               "--response-length", str(args.response_length), "--max-num-seqs", str(args.max_num_seqs),
               "--rollout-iters", str(args.rollout_iters)]
        t0 = time.time()
        try:
            r = subprocess.run(cmd, timeout=args.phase_timeout, capture_output=True, text=True)
            if r.returncode != 0:
                record(f"{ph}_error", {"rc": r.returncode, "stderr_tail": r.stderr[-1500:]})
        except subprocess.TimeoutExpired:
            record(f"{ph}_error", {"timeout_s": args.phase_timeout})
        print(f"[stack_bench] {ph}: {time.time() - t0:.0f}s", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
