import os

import pytest
import torch

from nanorlhf_b200.config import RLConfig
from nanorlhf_b200.models.lora import LoraConfig, LoraLinear, PeftModel, get_peft_model
from nanorlhf_b200.models.qwen2 import (Qwen2Config, Qwen2ForCausalLM, Qwen2ForSequenceClassification, forward,
                                        pack_padded, response_logprobs)


def test_config_defaults_and_overrides():
    a = RLConfig()
    assert (a.kl_coef, a.temperature, a.learning_rate, a.response_length) == (0.01, 0.9, 6e-6, 1500)
    assert (a.per_device_train_batch_size, a.gradient_accumulation_steps, a.num_mini_batches) == (4, 8, 16)
    assert a.lr_scheduler_type == "cosine_with_min_lr" and a.lr_scheduler_kwargs == {"min_lr_rate": 0.1}
    assert a.lora_r == 64 and a.lora_alpha == 16 and a.modules_to_save == ["embed_tokens", "lm_head", "score"]
    a.apply_overrides(["--kl_coef=0.05", "--response-length", "77", "--use_lora=false", "--modules_to_save=[\"lm_head\"]"],
                      env={"NANORLHF_TOP_P": "0.5"})
    assert a.kl_coef == 0.05 and a.response_length == 77 and a.use_lora is False and a.top_p == 0.5
    assert a.modules_to_save == ["lm_head"]
    with pytest.raises(ValueError):
        a.apply_overrides(["--no_such_flag=1"], env={})


def _tiny(dtype=torch.float32):
    return Qwen2ForCausalLM.from_config(Qwen2Config.tiny(vocab_size=97), dtype, seed=0)


def test_lora_zero_init_merge_and_roundtrip(tmp_path):
    base = _tiny()
    ref_out = None
    ids = torch.randint(0, 90, (12,))
    cu = torch.tensor([0, 5, 12], dtype=torch.int32)
    pos = torch.tensor([0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 6])
    ref_out = base(ids, cu, pos, 7)
    m = get_peft_model(base, LoraConfig(r=4, lora_alpha=8, modules_to_save=["embed_tokens", "lm_head"]))
    assert torch.allclose(m(ids, cu, pos, 7), ref_out, atol=1e-6)                 # B = 0 at init
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert any("lora_A" in n for n in names) and any("embed_tokens" in n for n in names) and any("lm_head" in n for n in names)
    assert m.base_model.lm_head.weight is not m.base_model.model.embed_tokens.weight   # un-tied copies
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, LoraLinear):
                mod.lora_B.weight.normal_(0, 0.05)
    out = m(ids, cu, pos, 7)
    sd = m.adapter_state_dict()
    assert "base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight" in sd
    assert "base_model.model.lm_head.weight" in sd
    m.save_pretrained(str(tmp_path))
    assert os.path.exists(tmp_path / "adapter_model.safetensors") and os.path.exists(tmp_path / "adapter_config.json")
    m2 = PeftModel.from_pretrained(_tiny(), str(tmp_path))
    assert torch.allclose(m2(ids, cu, pos, 7), out, atol=1e-6)
    merged = m.merge_and_unload()
    assert torch.allclose(merged(ids, cu, pos, 7), out, atol=1e-5)


def test_packed_forward_equals_padded_semantics():
    torch.manual_seed(0)
    m = _tiny().eval()
    pad = 96
    qr = torch.randint(0, 90, (3, 10))
    qr[0, :3] = pad
    qr[1, 8:] = pad
    ctx = 4
    logits = forward(m, qr, pad)[0]
    # row 2 has no padding: compare with a plain single-sequence forward
    ids, cu, pos, mx, flat = pack_padded(qr[2:3], pad)
    assert torch.allclose(logits[2], m(ids, cu, pos, mx), atol=1e-5)
    # left padding must not change the real tokens' outputs (position ids restart after the pads)
    ids0, cu0, pos0, mx0, _ = pack_padded(qr[0:1], pad)
    assert pos0.tolist() == list(range(7))
    lp, ent = response_logprobs(m, qr, ctx, pad, 0.9, want_entropy=True)
    z = logits[:, ctx - 1:-1] / 0.9
    want = torch.log_softmax(z, -1).gather(2, qr[:, ctx:].clamp(max=95).unsqueeze(-1)).squeeze(-1)
    real = (qr[:, ctx:] != pad)
    assert torch.allclose(lp[real], want[real], atol=1e-4)
    assert (lp[~real] == 1.0).all()
    vm = Qwen2ForSequenceClassification.from_causal_lm(m)
    out = response_logprobs(m, qr, ctx, pad, 0.9, value_model=vm)
    assert out[2].shape == lp.shape


def test_bucket_padded_plan_gives_the_same_logprobs_and_grads():
    """trainer/graphed.py pads the packed plan with a dummy sequence and dump-row entries so that the micro-step
    has a static shape; neither may change the log-probs, the entropies or the gradients."""
    from nanorlhf_b200.models.qwen2 import build_logprob_plan, planned_response_logprobs
    from nanorlhf_b200.trainer.graphed import GraphedMicroStep
    torch.manual_seed(0)
    m = _tiny()
    pad, ctx = 96, 4
    qr = torch.randint(0, 90, (3, 12))
    qr[0, :3] = pad
    qr[1, 9:] = pad
    B, T_r = 3, 12 - ctx

    def run(plan):
        m.zero_grad()
        lp, ent, _ = planned_response_logprobs(m, plan, B, T_r, 0.9, True, max_seqlen=12)
        real = qr[:, ctx:] != pad
        (lp[real] ** 2).sum().backward()
        return lp.detach(), ent.detach(), torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone()

    plan = build_logprob_plan(qr, ctx, pad)
    T, R = plan["ids"].numel(), plan["src"].numel()
    lp0, ent0, g0 = run(plan)
    for T_b, R_b in ((T + 5, R + 7), (T, R)):                 # with and without a dummy sequence
        padded = GraphedMicroStep._pad_plan(plan, B, pad, T_b, R_b)
        padded["max_seqlen"] = 12
        lp1, ent1, g1 = run(padded)
        assert torch.allclose(lp0, lp1, atol=1e-5) and torch.allclose(ent0, ent1, atol=1e-5)
        assert torch.allclose(g0, g1, atol=1e-5)


def test_save_load_pretrained_roundtrip(tmp_path):
    m = _tiny()
    m.save_pretrained(str(tmp_path))
    m2 = Qwen2ForCausalLM.from_pretrained(str(tmp_path), torch.float32)
    ids = torch.randint(0, 90, (6,))
    cu = torch.tensor([0, 6], dtype=torch.int32)
    assert torch.allclose(m(ids, cu, torch.arange(6), 6), m2(ids, cu, torch.arange(6), 6), atol=1e-6)


def test_grouped_lora_projections_equal_independent_ones():
    """q/k/v (and gate/up) go through ``lora_group_forward``: same outputs and gradients as calling each LoraLinear on its own
    (on CUDA the group shares its rank-r launches, ops/gemm.py; here the dispatcher's per-projection path is exercised)."""
    import torch
    from nanorlhf_b200.models.lora import LoraLinear, lora_group_forward
    torch.manual_seed(0)
    mods = [LoraLinear(torch.nn.Linear(32, n, bias=(n != 16)), r=4, alpha=8) for n in (48, 16, 16)]
    for m in mods:
        torch.nn.init.normal_(m.lora_B.weight, std=0.1)
    x = torch.randn(5, 7, 32, requires_grad=True)
    ys = lora_group_forward(x, mods)
    sum(y.square().sum() for y in ys).backward()
    g_group = [x.grad.clone()] + [m.lora_A.weight.grad.clone() for m in mods] + [m.lora_B.weight.grad.clone() for m in mods]
    x.grad = None
    for m in mods:
        m.zero_grad()
    ys2 = [m(x) for m in mods]
    sum(y.square().sum() for y in ys2).backward()
    g_ind = [x.grad] + [m.lora_A.weight.grad for m in mods] + [m.lora_B.weight.grad for m in mods]
    for a, b in zip(ys, ys2):
        assert torch.allclose(a, b, atol=1e-6)
    for a, b in zip(g_group, g_ind):
        assert torch.allclose(a, b, atol=1e-5)
    # a plain nn.Linear in the list (no adapter on that projection) falls back to independent calls
    mixed = [mods[0], torch.nn.Linear(32, 8)]
    out = lora_group_forward(x, mixed)
    assert out[1].shape == (5, 7, 8) and torch.allclose(out[0], mods[0](x))
