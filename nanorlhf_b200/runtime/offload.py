"""K-OFF: role-based HBM <-> pinned-host tiering on a side stream.

The reference swaps whole models with synchronous, pageable, tensor-by-tensor ``.to('cpu')`` /
``.to(device)`` calls four times per update and walks the optimizer state dict doing the same
(/root/reference/GRPO/grpo_trainer.py:124,164,168-172,475,525,622,625,728; GRPO/grpo.py:164,195),
followed by ``torch.cuda.empty_cache()`` storms.  On a 180 GB B200 every role of the 1.5B/7B
configs fits resident, so the default residency is ``resident`` (all calls below are no-ops).
``host`` residency is kept as a first-class capability (BASELINE.json config 5):

* each role owns ONE flat device buffer (params are views) and ONE pinned host mirror, so a swap is
  a single ``cudaMemcpyAsync`` at PCIe/C2C line rate instead of hundreds of small pageable copies;
* copies run on a dedicated side stream; the compute stream only waits on an event at the point
  the role is first used (``fetch``) -- never a device-wide sync, never ``empty_cache``;
* read-only roles (ref, reward) are *clean*: evicting them is free (drop the device buffer; the
  host mirror is already current); dirty roles (optimizer moments, policy) copy back on evict.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


class _Role:
    def __init__(self, name: str, tensors: List[torch.Tensor], residency: str, dirty: bool, rebinder):
        self.name, self.residency, self.dirty = name, residency, dirty
        self.rebinder = rebinder            # callable(list of device views or None)
        self.shapes = [t.shape for t in tensors]
        self.numels = [t.numel() for t in tensors]
        self.dtype = tensors[0].dtype if tensors else torch.float32
        self.device = tensors[0].device if tensors else torch.device("cpu")
        self.total = sum(self.numels)
        self.on_device = True
        self.host: Optional[torch.Tensor] = None
        self.dev: Optional[torch.Tensor] = None
        self.ready_event = None

    def views(self, flat: torch.Tensor) -> List[torch.Tensor]:
        out, off = [], 0
        for n, s in zip(self.numels, self.shapes):
            out.append(flat[off:off + n].view(s))
            off += n
        return out


class TieringEngine:
    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self.enabled = self.device.type == "cuda"
        self.side = torch.cuda.Stream(self.device) if self.enabled else None
        self.roles: Dict[str, _Role] = {}
        self.bytes_h2d = 0
        self.bytes_d2h = 0

    # ---- registration ------------------------------------------------------------------------
    def register(self, name: str, module: torch.nn.Module, residency: str = "resident", dirty: bool = False):
        """Register a module's parameters+buffers as one role.  Must be on ``self.device`` already
        (or on the CPU, in which case it is moved once)."""
        if residency not in ("resident", "host"):
            raise ValueError(f"unknown residency {residency!r}")
        module.to(self.device)
        if not self.enabled or residency == "resident":
            self.roles[name] = _Role(name, [], "resident", dirty, None)
            return
        seen, tensors = set(), []
        for t in list(module.parameters()) + list(module.buffers()):
            if id(t) not in seen and t.numel() > 0:
                seen.add(id(t))
                tensors.append(t)
        by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
        for t in tensors:
            by_dtype.setdefault(t.dtype, []).append(t)
        # one sub-role per dtype keeps every flat buffer homogeneous
        for i, (dt, ts) in enumerate(by_dtype.items()):
            def rebinder(views, ts=ts):
                for t, v in zip(ts, views):
                    t.data = v
            self._make_role(f"{name}#{i}" if len(by_dtype) > 1 else name, ts, residency, dirty, rebinder)
        self.roles.setdefault(name, _Role(name, [], "group", dirty, None))
        self.roles[name].members = [k for k in self.roles if k == name or k.startswith(name + "#")]

    def register_optimizer(self, name: str, optimizer, residency: str = "resident"):
        if not self.enabled or residency == "resident" or not hasattr(optimizer, "state_tensors"):
            self.roles[name] = _Role(name, [], "resident", True, None)
            return
        self._lazy_opt = (name, optimizer)
        self.roles[name] = _Role(name, [], "lazy", True, None)

    def _materialise_optimizer(self, name: str):
        _, optimizer = self._lazy_opt
        st = optimizer.state_tensors()
        if not st:
            return False
        keys = list(st)
        members = []
        for k in keys:
            def rebinder(views, k=k):
                optimizer.set_state_tensor(k, views[0])
            self._make_role(f"{name}#{k}", [st[k]], "host", True, rebinder)
            members.append(f"{name}#{k}")
        grp = _Role(name, [], "group", True, None)
        grp.members = members
        self.roles[name] = grp
        return True

    def _make_role(self, name, tensors, residency, dirty, rebinder):
        r = _Role(name, tensors, residency, dirty, rebinder)
        flat = torch.empty(r.total, dtype=r.dtype, device=self.device)
        with torch.no_grad():
            for v, t in zip(r.views(flat), tensors):
                v.copy_(t.detach())
        r.dev = flat
        rebinder(r.views(flat))
        r.host = torch.empty(r.total, dtype=r.dtype, pin_memory=True)
        # the host mirror is made current once, at registration, on the side stream
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.side):
            r.host.copy_(flat, non_blocking=True)
        self.bytes_d2h += flat.numel() * flat.element_size()
        r.ready_event = self.side.record_event()
        self.roles[name] = r
        return r

    # ---- movement ------------------------------------------------------------------------------
    def _members(self, name: str) -> List[_Role]:
        r = self.roles.get(name)
        if r is None:
            raise KeyError(f"unknown tiering role {name!r}")
        if r.residency == "lazy":
            if not self._materialise_optimizer(name):
                return []
            r = self.roles[name]
        if r.residency == "group":
            return [self.roles[m] for m in r.members if self.roles[m].residency == "host"]
        return [r] if r.residency == "host" else []

    def evict(self, name: str):
        """Role leaves HBM (async).  No-op for resident roles."""
        for r in self._members(name):
            if not r.on_device:
                continue
            cur = torch.cuda.current_stream(self.device)
            if r.dirty:
                self.side.wait_stream(cur)
                with torch.cuda.stream(self.side):
                    r.host.copy_(r.dev, non_blocking=True)
                r.dev.record_stream(self.side)
                self.bytes_d2h += r.total * r.dev.element_size()
            else:
                r.dev.record_stream(cur)
            r.rebinder([torch.empty(0, dtype=r.dtype, device=self.device) for _ in r.numels])
            r.dev = None
            r.on_device = False

    def prefetch(self, name: str):
        """Start the H2D copy on the side stream without making the compute stream wait."""
        for r in self._members(name):
            if r.on_device or r.dev is not None:
                continue
            with torch.cuda.stream(self.side):
                r.dev = torch.empty(r.total, dtype=r.dtype, device=self.device)
                r.dev.copy_(r.host, non_blocking=True)
                r.ready_event = self.side.record_event()
            self.bytes_h2d += r.total * r.dev.element_size()

    def fetch(self, name: str):
        """Role must be usable by kernels subsequently enqueued on the current stream."""
        self.prefetch(name)
        for r in self._members(name):
            if r.on_device:
                continue
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(r.ready_event)
            r.dev.record_stream(cur)
            r.rebinder(r.views(r.dev))
            r.on_device = True

    def stats(self):
        return {"offload/h2d_gb": self.bytes_h2d / 2**30, "offload/d2h_gb": self.bytes_d2h / 2**30}
