"""Fused optimizer boundary: gradient-norm clipping + AdamW for one model in two multi-tensor CUDA
launches (csrc/optim.cu), with torch.optim.AdamW's state layout so `optimizer.state_dict()` / `load_state_dict()` and
the reference checkpoint format (train_hybrid.py:596-606, 502-515) are unchanged."""
import math

import numpy as np
import torch

from . import _capi
from ._capi import check

_CHUNK = 8192


class ClipAdamW(torch.optim.AdamW):
    """torch.optim.AdamW whose step() also performs clip_grad_norm_(params, max_grad_norm) (train_hybrid.py:913-915)
    on the device, without a host sync. Parameters whose grad is None are skipped, exactly like torch."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.max_grad_norm = float(max_grad_norm)
        self._host = None
        self.last_grad_norm_sq = None

    @torch.no_grad()
    def step(self, closure=None):
        lib = _capi.lib()
        stream = _capi.raw_stream()
        rows, chunks, steps, touched = [], [], [], []
        t = None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad
                if not g.is_contiguous() or g.dtype != torch.float32:
                    g = g.contiguous().float()
                    p.grad = g
                assert p.is_contiguous() and p.dtype == torch.float32
                idx = len(rows)
                rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                             p.numel()))
                chunks += [(idx, c) for c in range((p.numel() + _CHUNK - 1) // _CHUNK)]
                steps.append(st["step"])
                touched.append(p)
        if not rows:
            return None
        group = self.param_groups[0]          # the trainer uses a single group per model
        dev = steps[0].device
        torch._foreach_add_(steps, 1.0)
        # every live tensor of a model advances together: take t from a python-side counter (no device read)
        self._t = getattr(self, "_t", 0) + 1
        t = self._t
        tab = np.array(rows, dtype=np.int64)                       # 5 x int64 = 40 bytes per tensor
        ch = np.array(chunks, dtype=np.int32)
        tab_d = torch.from_numpy(tab).pin_memory().to(dev, non_blocking=True)
        ch_d = torch.from_numpy(ch).pin_memory().to(dev, non_blocking=True)
        norm2 = torch.empty(1024, device=dev, dtype=torch.float32)     # per-block partials of sum g^2
        check(lib.lun_multi_grad_sumsq(tab_d.data_ptr(), ch_d.data_ptr(), len(chunks), norm2.data_ptr(), stream),
              "lun_multi_grad_sumsq")
        b1, b2 = group["betas"]
        check(lib.lun_multi_clip_adamw(tab_d.data_ptr(), ch_d.data_ptr(), len(chunks), norm2.data_ptr(),
                                       self.max_grad_norm, float(group["lr"]), b1, b2, group["eps"],
                                       group["weight_decay"], 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), stream),
              "lun_multi_clip_adamw")
        # the kernels wrote the parameters through raw pointers: tell autograd they changed in place, so every cache
        # keyed on `param._version` (the packed bf16 kernel operands of the drop-in modules) is rebuilt next forward
        torch.autograd.graph.increment_version(touched)
        self.last_grad_norm_sq = norm2
        self._keep = (tab_d, ch_d)                                 # keep the tables alive until the kernels ran
        return None

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        ts = [float(s["step"]) for s in self.state.values() if "step" in s]
        self._t = int(max(ts)) if ts else 0
