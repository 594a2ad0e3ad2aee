"""The sm_100a rollout engine (replaces vLLM; SURVEY.md section 2.6 first row, K1-K5).

* lives in the trainer's process / CUDA context, constructed once per policy (no boot, no teardown);
* weights: a fused arena (qkv / gate_up concatenated, LoRA merged) refreshed on device by
  ``sync_weights`` (parallel/weight_sync.py, K-BC); embed / lm_head alias the live parameters;
* KV: paged (16-token pages), pages handed out by the C++ ``Scheduler`` (csrc/runtime.cpp) with prompt
  pages shared between the N samples of a prompt; "reserve" policy => decode never allocates, so
* the decode step (28 layers x {fused add+RMSNorm, tcgen05 GEMM(+bias), RoPE, KV page write, paged
  decode attention, SwiGLU} + lm_head GEMM + top-p sampler + state update) is ONE CUDA graph per batch
  bucket; the host only looks at the ``finished`` flags every ``sync_every`` steps;
* sampling semantics of the reference's ``vllm_generate``: temperature, top-p 0.95, n samples per
  prompt, seeded, stop at EOS (kept), right-pad with pad_id, prompt-major / sample-minor order
  (/root/reference/GRPO/grpo_trainer.py:122-166); temperature 0 = greedy (ReMax baseline / eval).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from ..models.lora import LoraLinear
from ..ops import gemm as _gemm, native, reference as ref
from ..ops.attention import attention_varlen


def _unwrap(model):
    m = getattr(model, "policy", model)
    return getattr(m, "base_model", m) if hasattr(m, "peft_config") else m


class _LayerWeights:
    __slots__ = ("wqkv", "bqkv", "wo", "wgu", "wgu_i", "wdown", "ln1", "ln2", "q8")


class NativeSampler:
    PAGE = 16

    def __init__(self, model, rollout_dtype: str = "bf16", kv_block_size: int = 16, max_num_seqs: int = 4096,
                 kv_memory_fraction: float = 0.85, prefill_token_budget: int = 32768, sync_every: int = 64,
                 use_cuda_graph: bool = True, kv_cache_gb: Optional[float] = None, kv_cache_dtype: Optional[str] = None):
        if kv_block_size != self.PAGE:
            raise ValueError("the decode attention kernel is built for 16-token pages")
        native.load()
        self.model = model
        self.lm = _unwrap(model)
        self.cfg = self.lm.config
        self.device = next(self.lm.parameters()).device
        if self.cfg.head_dim != 128:
            raise ValueError("native sampler kernels are specialised for head_dim 128 (Qwen2.5 family)")
        self.rollout_dtype = rollout_dtype              # GEMM operands: bf16 | fp8 (e4m3, per-token x per-channel scales)
        # KV pages: bf16 | fp8.  fp8 pages halve KV memory (2x the resident sequences); the fp8 decode kernel is
        # not faster than the bf16 one yet (in-smem dequant makes it issue-bound), so bf16 stays the default.
        self.kv_dtype = kv_cache_dtype or "bf16"
        self.max_num_seqs = max_num_seqs
        self.kv_memory_fraction = kv_memory_fraction
        self.kv_cache_gb = kv_cache_gb
        self.prefill_token_budget = prefill_token_budget
        self.sync_every = sync_every
        self.use_cuda_graph = use_cuda_graph
        self.sharded_sync = None                        # parallel.weight_sync.ShardedWeightSync under fused data parallelism
        self.enable_compaction = True                   # drop finished rows at sync points (tests switch it off for A/B)
        self.layers: List[_LayerWeights] = []
        self._weights_version = None
        self.k_cache: List[torch.Tensor] = []
        self.v_cache: List[torch.Tensor] = []
        self.num_blocks = 0
        self._graphs: Dict[int, tuple] = {}
        self._state_pool: Dict[tuple, dict] = {}
        self.stats = {"decode_steps": 0, "decode_tokens": 0, "prefill_tokens": 0, "graph_replays": 0, "compactions": 0,
                      "row_steps": 0}
        self._build_arena()

    # ------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------
    def _build_arena(self):
        cfg, dev = self.cfg, self.device
        d, D = cfg.hidden_size, cfg.head_dim
        nq, nkv, F = cfg.num_attention_heads * D, cfg.num_key_value_heads * D, cfg.intermediate_size
        for _ in range(cfg.num_hidden_layers):
            lw = _LayerWeights()
            lw.wqkv = torch.empty(nq + 2 * nkv, d, dtype=torch.bfloat16, device=dev)
            lw.bqkv = torch.zeros(nq + 2 * nkv, dtype=torch.bfloat16, device=dev)
            lw.wo = torch.empty(d, nq, dtype=torch.bfloat16, device=dev)
            lw.wgu = torch.empty(2 * F, d, dtype=torch.bfloat16, device=dev)
            # gate/up rows interleaved per 32 features for the fused SwiGLU GEMM epilogue
            lw.wgu_i = torch.empty(2 * F, d, dtype=torch.bfloat16, device=dev) if (F % 64 == 0) else None
            lw.wdown = torch.empty(d, F, dtype=torch.bfloat16, device=dev)
            lw.q8 = None          # fp8 rollout: {name: (e4m3 weight, per-channel scale)} built at weight refresh
            self.layers.append(lw)

    # ---- linear layers of the sampler: bf16 tcgen05 GEMM, or e4m3 x e4m3 (per-token x per-channel scales) ----
    def _lin(self, lw, name, x, bias=None):
        if lw.q8 is not None:
            native._count(2)
            xq, xs = native.ext().quant_rows_e4m3(x)
            wq, ws = lw.q8[name]
            return native.ext().gemm_tc_fp8(xq, xs, wq, ws, bias, False)          # kind::f8f6f4 on CTA pairs (gemm_tc.cu)
        # general tcgen05 GEMM (cta_group::2 where it pays); with few output tiles (small decode batches: the tail of a rollout,
        # the 7B configs) the dispatcher splits the contraction so that all SMs stream the weight (gemm_tc.cu split-K)
        return _gemm.gemm(x, getattr(lw, name), bias=bias)

    def quantize_arena(self):
        """fp8 rollout (rollout_dtype="fp8"): e4m3 copies of the merged arena with per-output-channel scales."""
        for lw in self.layers:
            lw.q8 = {}
            for name in ("wqkv", "wo", "wdown") + (("wgu_i",) if lw.wgu_i is not None else ("wgu",)):
                native._count()
                lw.q8[name] = native.ext().quant_rows_e4m3(getattr(lw, name))

    def release(self):
        """Free the arena, the KV pages and the captured graphs (the engine is unusable afterwards)."""
        self._graphs.clear()
        self.layers, self.k_cache, self.v_cache, self.num_blocks = [], [], [], 0

    def weights_fingerprint(self):
        """Cheap change detector: the optimizer bumps ``_nrl_version`` on the policy after every step."""
        return getattr(self.lm, "_nrl_version", 0)

    @torch.no_grad()
    def sync_weights(self, force: bool = False):
        """Refresh the fused sampler arena from the live policy (LoRA merged on device)."""
        from ..parallel.weight_sync import refresh_sampler_arena
        ver = self.weights_fingerprint()
        if not force and self._weights_version == ver:
            return
        if self.sharded_sync is not None:
            self.sharded_sync.refresh()          # K-BC: layer-sharded merge, multimem.st into every rank's arena
        else:
            refresh_sampler_arena(self)
        if self.rollout_dtype == "fp8":
            self.quantize_arena()
        self._weights_version = ver

    # ------------------------------------------------------------------------------------------
    # KV cache
    # ------------------------------------------------------------------------------------------
    def _ensure_kv(self, blocks_needed: int):
        cfg = self.cfg
        if self.num_blocks >= blocks_needed:
            return
        per_block = 2 * cfg.num_hidden_layers * cfg.num_key_value_heads * self.PAGE * cfg.head_dim * (1 if self.kv_dtype == "fp8" else 2)
        self.k_cache, self.v_cache = [], []          # the old pool is dropped first so the allocator can reuse it
        self._graphs.clear()
        self._state_pool.clear()
        free, _total = torch.cuda.mem_get_info(self.device)
        budget = int(free * self.kv_memory_fraction) if self.kv_cache_gb is None else int(self.kv_cache_gb * 2**30)
        can = budget // per_block
        n = int(min(max(blocks_needed, 64), can))
        if n < 64:
            raise RuntimeError("not enough free HBM for a KV cache")
        shape = (n, cfg.num_key_value_heads, self.PAGE, cfg.head_dim)
        kv_dtype = torch.uint8 if self.kv_dtype == "fp8" else torch.bfloat16
        self.k_scale, self.v_scale = [], []
        for _ in range(cfg.num_hidden_layers):
            self.k_cache.append(torch.zeros(shape, dtype=kv_dtype, device=self.device))
            self.v_cache.append(torch.zeros(shape, dtype=kv_dtype, device=self.device))
            if self.kv_dtype == "fp8":     # e4m3 pages + one fp32 scale per (token, kv head)
                self.k_scale.append(torch.ones(shape[:3], dtype=torch.float32, device=self.device))
                self.v_scale.append(torch.ones(shape[:3], dtype=torch.float32, device=self.device))
        self.num_blocks = n

    # ------------------------------------------------------------------------------------------
    # model math on the fused arena
    # ------------------------------------------------------------------------------------------
    def _embed(self, ids):
        return self.lm.model.embed_tokens.weight[ids.long()]

    def _layer_prefill(self, li, x, res, cos, sin, cu, max_len, slot, src):
        cfg, lw = self.cfg, self.layers[li]
        D, Hq, Hkv = cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads
        h, res = (native.rmsnorm(x, lw.ln1, cfg.rms_norm_eps), x) if res is None else native.add_rmsnorm(x, res, lw.ln1, cfg.rms_norm_eps)
        qkv = self._lin(lw, "wqkv", h, lw.bqkv)
        T = qkv.shape[0]
        q = qkv[:, :Hq * D].view(T, Hq, D)
        k = qkv[:, Hq * D:(Hq + Hkv) * D].view(T, Hkv, D)
        v = qkv[:, (Hq + Hkv) * D:].view(T, Hkv, D)
        native.ext().rope(q, cos, sin, 1.0, True)
        native.ext().rope(k, cos, sin, 1.0, True)
        self._kv_write(li, k, v, slot, src)
        att = attention_varlen(q, k, v, cu, max_len, causal=True)
        o = self._lin(lw, "wo", att.reshape(T, Hq * D))
        h, res = native.add_rmsnorm(o, res, lw.ln2, cfg.rms_norm_eps)
        return self._lin(lw, "wdown", self._mlp_act(lw, h)), res

    def _kv_write(self, li, k, v, slot, src):
        if self.kv_dtype == "fp8":
            native._count()
            native.ext().kv_cache_write_fp8(k, v, self.k_cache[li], self.v_cache[li], self.k_scale[li], self.v_scale[li], slot, src)
        else:
            native.kv_cache_write(k, v, self.k_cache[li], self.v_cache[li], slot, src)

    def _mlp_act(self, lw, h):
        if lw.q8 is not None:
            native._count(2)
            hq, hs = native.ext().quant_rows_e4m3(h)
            if lw.wgu_i is not None:
                wq, ws = lw.q8["wgu_i"]
                return native.ext().gemm_tc_fp8(hq, hs, wq, ws, None, True)
            wq, ws = lw.q8["wgu"]
            return native.ext().swiglu(native.ext().gemm_tc_fp8(hq, hs, wq, ws, None, False))
        if lw.wgu_i is not None:
            native._count()
            return native.ext().gemm_tc_swiglu(h, lw.wgu_i, None)
        native._count(1)
        return native.ext().swiglu(_gemm.gemm(h, lw.wgu, split_k=0))

    def _layer_decode(self, li, x, res, cos, sin, st):
        cfg, lw = self.cfg, self.layers[li]
        D, Hq, Hkv = cfg.head_dim, cfg.num_attention_heads, cfg.num_key_value_heads
        if res is None:
            h, res = native.ext().rmsnorm(x, lw.ln1, cfg.rms_norm_eps, None, False)[0], x
        else:
            h, res = native.add_rmsnorm(x, res, lw.ln1, cfg.rms_norm_eps)
        qkv = self._lin(lw, "wqkv", h, lw.bqkv)
        S = qkv.shape[0]
        q = qkv[:, :Hq * D].view(S, Hq, D)
        k = qkv[:, Hq * D:(Hq + Hkv) * D].view(S, Hkv, D)
        v = qkv[:, (Hq + Hkv) * D:].view(S, Hkv, D)
        # RoPE(q), RoPE(k) and the page write of (k, v) are one launch (csrc/rope_kv.cu)
        native._count()
        fp8 = self.kv_dtype == "fp8"
        native.ext().rope_kv_write(qkv, cos, sin, self.k_cache[li], self.v_cache[li], self.k_scale[li] if fp8 else None,
                                   self.v_scale[li] if fp8 else None, st["slot"], Hq, Hkv)
        if self.kv_dtype == "fp8":
            native._count()
            att = native.ext().paged_decode_fp8(q, self.k_cache[li], self.v_cache[li], self.k_scale[li], self.v_scale[li],
                                                st["block_tables"], st["ctx_lens"], 1.0 / math.sqrt(D), st["splits"])
        else:
            att = native.paged_decode(q, self.k_cache[li], self.v_cache[li], st["block_tables"], st["ctx_lens"],
                                      1.0 / math.sqrt(D), st["splits"])
        o = self._lin(lw, "wo", att.view(S, Hq * D))
        h, res = native.add_rmsnorm(o, res, lw.ln2, cfg.rms_norm_eps)
        native._count(2)
        return self._lin(lw, "wdown", self._mlp_act(lw, h)), res

    def _final_logits(self, x, res):
        cfg = self.cfg
        h, _ = native.add_rmsnorm(x, res, self.lm.model.norm.weight, cfg.rms_norm_eps)
        return _gemm.gemm(h, self.lm.lm_head.weight, split_k=0)

    # ------------------------------------------------------------------------------------------
    # prefill
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _prefill(self, sched, groups: List[int], prompts: Dict[int, Sequence[int]]):
        """Run the prompts of ``groups`` through the model, writing KV pages; returns last-token logits
        [len(groups), V]."""
        cfg, dev = self.cfg, self.device
        ids, pos, cu, slot, src = [], [], [0], [], []
        off = 0
        for g in groups:
            p = list(prompts[g])
            ids += p
            pos += list(range(len(p)))
            cu.append(cu[-1] + len(p))
            tok, sl = sched.prefill_slots(g)
            src += [off + t for t in tok]
            slot += sl
            off += len(p)
        ids_t = torch.tensor(ids, dtype=torch.long, device=dev)
        pos_t = torch.tensor(pos, dtype=torch.long, device=dev)
        cu_t = torch.tensor(cu, dtype=torch.int32, device=dev)
        slot_t = torch.tensor(slot, dtype=torch.int32, device=dev)
        src_t = torch.tensor(src, dtype=torch.int32, device=dev)
        max_len = max(b - a for a, b in zip(cu[:-1], cu[1:]))
        cos, sin = ref.rope_cos_sin(pos_t, cfg.head_dim, cfg.rope_theta)
        x, res = self._embed(ids_t), None
        for li in range(cfg.num_hidden_layers):
            x, res = self._layer_prefill(li, x, res, cos, sin, cu_t, max_len, slot_t, src_t)
        last = cu_t[1:].long() - 1
        self.stats["prefill_tokens"] += len(ids)
        return self._final_logits(x[last].contiguous(), res[last].contiguous())

    # ------------------------------------------------------------------------------------------
    # decode
    # ------------------------------------------------------------------------------------------
    def _alloc_state(self, S: int, max_blocks: int, max_tokens: int, pad_id: int = 0):
        """Batch state of ``S`` rows.  States are pooled per shape: the decode CUDA graph of a bucket is captured
        against these exact tensors, so a batch that shrinks (compaction) or grows (admission) back into a bucket it
        has used before replays the existing graph instead of capturing a new one."""
        dev = self.device
        key = (S, max_blocks, max_tokens)
        st = self._state_pool.get(key)
        if st is None:
            i32 = dict(dtype=torch.int32, device=dev)
            st = {
                "S": S, "tokens": torch.zeros(S, **i32), "positions": torch.zeros(S, **i32),
                "ctx_lens": torch.ones(S, **i32), "block_tables": torch.zeros(S, max_blocks, **i32),
                "slot": torch.zeros(S, **i32), "finished": torch.ones(S, dtype=torch.bool, device=dev),
                "gen_count": torch.zeros(S, **i32), "row_ids": torch.zeros(S, **i32),
                "out": torch.full((S, max_tokens + 1), pad_id, dtype=torch.int32, device=dev),
                "rows": torch.arange(S, device=dev), "splits": 1,
            }
            if len(self._state_pool) >= 8:
                old = next(iter(self._state_pool))
                self._state_pool.pop(old)
                self._graphs = {k: v for k, v in self._graphs.items() if v[1] is not None and (k[0], k[1], k[7]) != old}
            self._state_pool[key] = st
        else:
            for k, fill in (("tokens", 0), ("positions", 0), ("ctx_lens", 1), ("slot", 0), ("gen_count", 0), ("row_ids", 0)):
                st[k].fill_(fill)
            st["finished"].fill_(True)
            st["out"].fill_(pad_id)
        return st

    def _decode_step(self, st, temperature, top_p, seed, eos_id, pad_id, max_tokens):
        """One token for every row of the batch; pure device work (graph-capturable)."""
        cfg = self.cfg
        PAGE = self.PAGE
        pos = st["positions"]
        # slot of the token being fed (position = ctx_len - 1)
        page = torch.div(pos, PAGE, rounding_mode="floor").long()
        blk = st["block_tables"].gather(1, page[:, None]).squeeze(1)
        st["slot"].copy_(blk * PAGE + (pos % PAGE))
        cos, sin = ref.rope_cos_sin(pos, cfg.head_dim, cfg.rope_theta)
        x, res = self._embed(st["tokens"]), None
        for li in range(cfg.num_hidden_layers):
            x, res = self._layer_decode(li, x, res, cos, sin, st)
        logits = self._final_logits(x, res)
        tok = native.sample(logits, temperature, top_p, seed, 0, st["row_ids"], st["gen_count"])
        fin = st["finished"]
        tok = torch.where(fin, torch.full_like(tok, pad_id), tok)
        col = st["gen_count"].clamp(max=max_tokens).long()
        st["out"][st["rows"], col] = tok
        newly = (tok == eos_id) if eos_id is not None else torch.zeros_like(fin)
        live = ~fin
        st["gen_count"].add_(live.int())
        fin_new = fin | newly | (st["gen_count"] >= max_tokens)
        adv = (~fin_new).int()
        st["tokens"].copy_(tok)
        st["positions"].add_(adv)
        st["ctx_lens"].add_(adv)
        st["finished"].copy_(fin_new)

    def _run_decode(self, st, steps, temperature, top_p, seed, eos_id, pad_id, max_tokens):
        """Run up to ``steps`` decode iterations through a CUDA graph (one graph per batch bucket and
        sampling configuration)."""
        key = (st["S"], st["block_tables"].shape[1], float(temperature), float(top_p), int(seed), eos_id, pad_id,
               max_tokens, st["splits"])
        if not self.use_cuda_graph:
            for _ in range(steps):
                self._decode_step(st, temperature, top_p, seed, eos_id, pad_id, max_tokens)
            return
        entry = self._graphs.get(key)
        if entry is None or entry[1] is not st:
            # warm-up on a side stream (lazy inits, cudaFuncSetAttribute) then capture
            s = torch.cuda.Stream(self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            snap = {k: v.clone() for k, v in st.items() if isinstance(v, torch.Tensor)}
            with torch.cuda.stream(s):
                self._decode_step(st, temperature, top_p, seed, eos_id, pad_id, max_tokens)
            torch.cuda.current_stream(self.device).wait_stream(s)
            for k, v in snap.items():
                st[k].copy_(v)
            g = torch.cuda.CUDAGraph()
            n0 = native.launches()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._decode_step(st, temperature, top_p, seed, eos_id, pad_id, max_tokens)
            self._kernels_per_step = native.launches() - n0
            for k, v in snap.items():
                st[k].copy_(v)
            # keep at most a handful of graphs alive
            if len(self._graphs) > 6:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = entry = (g, st)
        g = entry[0]
        for _ in range(steps):
            g.replay()
        native._count(steps * getattr(self, "_kernels_per_step", 0))
        self.stats["graph_replays"] += steps

    # ------------------------------------------------------------------------------------------
    # public API
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, prompts: Sequence[Sequence[int]], n: int, temperature: float, top_p: float, max_tokens: int,
                 eos_id: Optional[int], pad_id: int, seed: int) -> torch.Tensor:
        from .. import _C
        dev, PAGE = self.device, self.PAGE
        was_training = self.lm.training
        self.lm.eval()
        P = len(prompts)
        total = P * n
        out = torch.full((total, max_tokens), pad_id, dtype=torch.long, device=dev)
        max_prompt = max(len(p) for p in prompts)
        per_seq_blocks = (max_prompt + max_tokens + PAGE - 1) // PAGE + 1
        need = sum((len(p) // PAGE) + n * ((len(p) + max_tokens + PAGE - 1) // PAGE - len(p) // PAGE) for p in prompts)
        self._ensure_kv(need + 8)
        sched = _C.Scheduler(self.num_blocks - 1, PAGE, self.max_num_seqs, True)   # last page = scratch for padding rows
        gid_to_prompt = {}
        for i, p in enumerate(prompts):
            gid = sched.add_request(len(p), n, max_tokens)
            gid_to_prompt[gid] = p
        scratch_block = self.num_blocks - 1
        running: List[int] = []          # seq ids in batch-row order
        st = None
        compact = False
        while True:
            admitted = sched.admit()
            if admitted or compact:
                compact = False
                # ---- prefill the newly admitted groups in token-budget chunks ----
                new_seq, first_tok_logits_rows, chunk, tok_count = [], [], [], 0
                logits_parts = []
                for g in admitted + [None]:
                    if g is not None:
                        chunk.append(g)
                        tok_count += len(gid_to_prompt[g])
                    if chunk and (g is None or tok_count >= self.prefill_token_budget):
                        logits_parts.append((list(chunk), self._prefill(sched, chunk, gid_to_prompt)))
                        chunk, tok_count = [], 0
                # ---- (re)build the batch state: surviving rows + new rows ----
                keep_rows = []
                if st is not None:
                    fin = st["finished"][:len(running)].tolist()
                    keep_rows = [i for i, f in enumerate(fin) if not f]
                    self._flush_rows(st, running, out, max_tokens, n)
                    sched.finish([running[i] for i, f in enumerate(fin) if f])
                new_running = [running[i] for i in keep_rows]
                old_state = st
                old_keep = {k: old_state[k][torch.tensor(keep_rows, device=dev)] for k in
                            ("tokens", "positions", "ctx_lens", "finished", "gen_count", "out")} if keep_rows else None
                for groups, _ in logits_parts:
                    for g in groups:
                        new_running += sched.group_seqs(g)
                S_real = len(new_running)
                if S_real == 0:
                    if sched.num_waiting() > 0:
                        raise RuntimeError("KV pool too small: a waiting request group cannot be admitted into an empty batch")
                    break
                S = max(128, (S_real + 127) // 128 * 128) if self.use_cuda_graph else S_real
                st = self._alloc_state(S, per_seq_blocks, max_tokens, pad_id)
                st["block_tables"].fill_(scratch_block)
                bt = torch.from_numpy(sched.block_tables(new_running, per_seq_blocks, scratch_block))     # built by the C++ scheduler
                st["block_tables"][:S_real].copy_(bt.to(dev))
                st["row_ids"][:S_real].copy_(torch.tensor(new_running, dtype=torch.int32))
                nk = len(keep_rows)
                if nk:                               # (gathered before the pooled state was reset: it may be the same tensors)
                    for k, v in old_keep.items():
                        st[k][:nk].copy_(v)
                # first token of every new sequence: sample from the prompt's last-position logits
                r0 = nk
                for groups, logits in logits_parts:
                    rep = torch.arange(len(groups), device=dev).repeat_interleave(n)
                    cnt = len(groups) * n
                    rows = slice(r0, r0 + cnt)
                    lg = logits[rep]
                    tok = native.sample(lg, temperature, top_p, seed, 0, st["row_ids"][rows], st["gen_count"][rows])
                    plen = torch.tensor([len(gid_to_prompt[g]) for g in groups], dtype=torch.int32, device=dev).repeat_interleave(n)
                    st["out"][rows, 0] = tok
                    st["tokens"][rows] = tok
                    st["positions"][rows] = plen
                    st["ctx_lens"][rows] = plen + 1
                    st["gen_count"][rows] = 1
                    done = (tok == eos_id) if eos_id is not None else torch.zeros(cnt, dtype=torch.bool, device=dev)
                    if max_tokens <= 1:
                        done = torch.ones_like(done)
                    st["finished"][rows] = done
                    r0 += cnt
                running = new_running
                avg_ctx = float(st["ctx_lens"][:S_real].float().mean()) + max_tokens / 2
                st["splits"] = 1 if S_real * self.cfg.num_key_value_heads >= 296 else int(min(8, max(1, math.ceil(
                    296 / max(1, S_real * self.cfg.num_key_value_heads)))))
            if st is None:
                break
            # ---- decode until the next sync point ----
            self._run_decode(st, self.sync_every, temperature, top_p, seed, eos_id, pad_id, max_tokens)
            self.stats["decode_steps"] += self.sync_every
            self.stats["row_steps"] += self.sync_every * st["S"]
            fin = st["finished"][:len(running)]
            n_fin = int(fin.sum().item())                     # the only host sync of the decode loop
            if n_fin == len(running) and sched.num_waiting() == 0:
                self._flush_rows(st, running, out, max_tokens, n)
                sched.finish(running)
                break
            if n_fin > 0:
                # Finished rows still cost a GEMM row and a (short) KV read every step.  Release their pages and
                # rebuild the batch without them when that lets a waiting group in, or when the survivors fit a
                # smaller graph bucket (the reference's vLLM does the same every step; here every sync point).
                live = len(running) - n_fin
                bucket = max(128, (live + 127) // 128 * 128) if self.use_cuda_graph else live
                if sched.num_waiting() > 0 or (self.enable_compaction and bucket < st["S"]):
                    done_rows = fin.nonzero().squeeze(1).tolist()
                    self._flush_rows(st, running, out, max_tokens, n)
                    sched.finish([running[i] for i in done_rows])
                    compact = True
                    self.stats["compactions"] += 1
        self.stats["decode_tokens"] += int((out != pad_id).sum().item())
        self.lm.train(was_training)
        return out

    def _flush_rows(self, st, running, out, max_tokens, n):
        """Copy generated ids of the batch rows into the [P*n, max_tokens] result (seq id == output row)."""
        if not running:
            return
        rows = torch.tensor(running, device=self.device)
        out[rows] = st["out"][:len(running), :max_tokens].long()
