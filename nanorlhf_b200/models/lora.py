"""LoRA adapters + ``modules_to_save`` with peft-compatible on-disk naming.

The reference wraps the policy with ``peft.get_peft_model(policy, LoraConfig(r=64, alpha=16,
target_modules=[q,k,v,o,gate,up,down]_proj, modules_to_save=[embed_tokens, lm_head, score]))``
(/root/reference/GRPO/grpo.py:228-243) and merges with ``merge_and_unload`` on the CPU before
every rollout (/root/reference/GRPO/grpo_trainer.py:131-138).  peft is not a dependency here.
Semantics owned by this file:

* ``y = W x + (alpha/r) * B (A x)``; A ~ kaiming-uniform, B = 0 (so step 0 equals the base model);
* ``modules_to_save`` modules become fully trainable (tied embeddings are un-tied by copying);
* adapters are saved as ``adapter_model.safetensors`` + ``adapter_config.json`` with peft's key
  names (``base_model.model.model.layers.N.self_attn.q_proj.lora_A.weight`` ...);
* merging is a GPU op (``merged_weight``) consumed by the sampler weight refresh (K-BC) -- the
  merged matrices never touch the host or the disk.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import asdict, dataclass, field
from typing import Dict, Iterable, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

DEFAULT_TARGETS = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]


@dataclass
class LoraConfig:
    r: int = 64
    lora_alpha: int = 16
    target_modules: List[str] = field(default_factory=lambda: list(DEFAULT_TARGETS))
    lora_dropout: float = 0.0
    bias: str = "none"
    task_type: str = "CAUSAL_LM"
    modules_to_save: Optional[List[str]] = None
    base_model_name_or_path: str = ""
    peft_type: str = "LORA"

    @property
    def scaling(self) -> float:
        return self.lora_alpha / self.r


class LoraLinear(nn.Module):
    """Frozen base ``nn.Linear`` + trainable rank-r update."""

    def __init__(self, base: nn.Linear, r: int, alpha: int, dropout: float = 0.0):
        super().__init__()
        self.base_layer = base
        self.r, self.scaling = r, alpha / r
        dev, dt = base.weight.device, base.weight.dtype
        self.lora_A = nn.Linear(base.in_features, r, bias=False, device=dev, dtype=dt)
        self.lora_B = nn.Linear(r, base.out_features, bias=False, device=dev, dtype=dt)
        nn.init.kaiming_uniform_(self.lora_A.weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B.weight)
        self.lora_dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        base.weight.requires_grad_(False)
        if base.bias is not None:
            base.bias.requires_grad_(False)

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    def forward(self, x):
        if isinstance(self.lora_dropout, nn.Identity) and not self.base_layer.weight.requires_grad:
            # one dual-source-K tcgen05 GEMM per direction on CUDA (ops/gemm.py); PyTorch oracle elsewhere
            return ops.lora_linear(x, self.base_layer.weight, self.base_layer.bias, self.lora_A.weight,
                                   self.lora_B.weight, self.scaling)
        y = self.base_layer(x)
        return y + self.lora_B(self.lora_A(self.lora_dropout(x))) * self.scaling

    @torch.no_grad()
    def merged_weight(self) -> torch.Tensor:
        """W + (alpha/r) B A, in the base dtype (fp32 accumulate)."""
        w = self.base_layer.weight
        delta = (self.lora_B.weight.float() @ self.lora_A.weight.float()) * self.scaling
        return (w.float() + delta).to(w.dtype)


class PeftModel(nn.Module):
    """Thin wrapper giving the adapter-aware API the trainers expect (``peft_type``, save/merge)."""

    def __init__(self, base: nn.Module, config: LoraConfig):
        super().__init__()
        self.base_model = base
        self.peft_config = config
        self.peft_type = "LORA"
        self.config = base.config
        self.name_or_path = getattr(base, "name_or_path", "")

    # delegate the model API
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("base_model"), name)

    def forward(self, *a, **kw):
        return self.base_model(*a, **kw)

    def trainable_parameters(self):
        return [p for p in self.parameters() if p.requires_grad]

    def print_trainable_parameters(self):
        t = sum(p.numel() for p in self.parameters() if p.requires_grad)
        seen, a = set(), 0
        for p in self.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                a += p.numel()
        print(f"trainable params: {t:,} || all params: {a:,} || trainable%: {100 * t / max(a, 1):.4f}")
        return t, a

    # ---- adapter state dict in peft's on-disk naming ----------------------------------------
    def adapter_state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for name, mod in self.base_model.named_modules():
            if isinstance(mod, LoraLinear):
                out[f"base_model.model.{name}.lora_A.weight"] = mod.lora_A.weight
                out[f"base_model.model.{name}.lora_B.weight"] = mod.lora_B.weight
        for name in _saved_module_names(self.base_model, self.peft_config.modules_to_save):
            mod = self.base_model.get_submodule(name)
            for pn, p in mod.named_parameters(recurse=False):
                out[f"base_model.model.{name}.{pn}"] = p
        return out

    def load_adapter_state_dict(self, sd: Dict[str, torch.Tensor]):
        own = self.adapter_state_dict()
        missing = [k for k in own if k not in sd]
        if missing:
            raise RuntimeError(f"adapter checkpoint is missing {missing[:4]}")
        with torch.no_grad():
            for k, p in own.items():
                p.copy_(sd[k].to(p.device, p.dtype))
        # a resident sampler merged the previous adapter: make it re-merge before the next rollout
        self.base_model._nrl_version = getattr(self.base_model, "_nrl_version", 0) + 1

    def save_pretrained(self, path: str, **_):
        from .hf_io import save_state_dict
        os.makedirs(path, exist_ok=True)
        cfg = asdict(self.peft_config)
        cfg["base_model_name_or_path"] = self.name_or_path
        with open(os.path.join(path, "adapter_config.json"), "w") as f:
            json.dump(cfg, f, indent=2)
        save_state_dict(self.adapter_state_dict(), os.path.join(path, "adapter_model.safetensors"))

    @classmethod
    def from_pretrained(cls, base: nn.Module, path: str):
        from .hf_io import load_state_dict
        with open(os.path.join(path, "adapter_config.json")) as f:
            raw = json.load(f)
        cfg = LoraConfig(**{k: v for k, v in raw.items() if k in LoraConfig.__dataclass_fields__})
        model = get_peft_model(base, cfg)
        model.load_adapter_state_dict(load_state_dict(os.path.join(path, "adapter_model.safetensors")))
        return model

    @torch.no_grad()
    def merge_and_unload(self) -> nn.Module:
        """Fold every adapter into its base weight and return the plain model (peft API parity)."""
        for name, mod in list(self.base_model.named_modules()):
            if isinstance(mod, LoraLinear):
                mod.base_layer.weight.copy_(mod.merged_weight())
                parent, leaf = _parent_and_leaf(self.base_model, name)
                setattr(parent, leaf, mod.base_layer)
        return self.base_model


def _parent_and_leaf(root: nn.Module, dotted: str):
    parts = dotted.split(".")
    parent = root.get_submodule(".".join(parts[:-1])) if len(parts) > 1 else root
    return parent, parts[-1]


def _saved_module_names(model: nn.Module, modules_to_save: Optional[Iterable[str]]) -> List[str]:
    if not modules_to_save:
        return []
    wanted = set(modules_to_save)
    return [n for n, _ in model.named_modules() if n and n.split(".")[-1] in wanted]


def get_peft_model(model: nn.Module, config: LoraConfig) -> PeftModel:
    """Freeze the base model, inject LoRA into ``target_modules``, un-freeze ``modules_to_save``."""
    for p in model.parameters():
        p.requires_grad_(False)
    targets = set(config.target_modules)
    for name, mod in list(model.named_modules()):
        if isinstance(mod, nn.Linear) and name.split(".")[-1] in targets:
            parent, leaf = _parent_and_leaf(model, name)
            setattr(parent, leaf, LoraLinear(mod, config.r, config.lora_alpha, config.lora_dropout))
    saved = _saved_module_names(model, config.modules_to_save)
    if saved:
        if "lm_head" in [s.split(".")[-1] for s in saved] and hasattr(model, "untie_weights"):
            model.untie_weights()   # peft makes independent trainable copies of tied embed/lm_head
        for name in saved:
            for p in model.get_submodule(name).parameters(recurse=False):
                p.requires_grad_(True)
    return PeftModel(model, config)


def lora_group_forward(x, mods):
    """``[m(x) for m in mods]`` for projections that share their input; when all of them are plain LoRA layers with the
    same rank / scaling their rank-r projections run as one launch (ops.lora_linear_group)."""
    if (len(mods) > 1 and all(isinstance(m, LoraLinear) for m in mods)
            and all(isinstance(m.lora_dropout, nn.Identity) and not m.base_layer.weight.requires_grad for m in mods)
            and len({(m.r, m.scaling) for m in mods}) == 1):
        return ops.lora_linear_group(x, [(m.base_layer.weight, m.base_layer.bias, m.lora_A.weight, m.lora_B.weight) for m in mods],
                                     mods[0].scaling)
    return [m(x) for m in mods]


def iter_lora_layers(model: nn.Module):
    for name, mod in model.named_modules():
        if isinstance(mod, LoraLinear):
            yield name, mod
