"""safetensors state-dict I/O (HF checkpoint format library; SURVEY.md App. C compatibility)."""
from __future__ import annotations

import torch


def save_state_dict(sd: dict, path: str, metadata: dict | None = None) -> None:
    from safetensors.torch import save_file
    # safetensors refuses aliased storage; clone shared tensors, keep everything contiguous on CPU
    seen, out = set(), {}
    for k, v in sd.items():
        t = v.detach()
        if t.device.type != "cpu":
            t = t.cpu()
        ptr = t.untyped_storage().data_ptr() if t.numel() else 0
        if ptr in seen:
            t = t.clone()
        seen.add(ptr)
        out[k] = t.contiguous()
    md = {"format": "pt"}
    md.update(metadata or {})
    save_file(out, path, metadata=md)


def load_state_dict(path: str, device: str = "cpu") -> dict:
    from safetensors.torch import load_file
    return load_file(path, device=device)
