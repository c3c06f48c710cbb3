"""Fused optimizer boundary: gradient-norm clipping + AdamW for one model in two multi-tensor CUDA
launches (csrc/optim.cu), with torch.optim.AdamW's state layout so `optimizer.state_dict()` / `load_state_dict()` and
the reference checkpoint format (train_hybrid.py:596-606, 502-515) are unchanged."""
import math

import numpy as np
import torch

from . import _capi
from ._capi import check

_CHUNK = 8192
# one 72-byte row per tensor, mirrored by `struct OptTensor` in csrc/optim.cu
_ROW = np.dtype([("p", "<i8"), ("g", "<i8"), ("m", "<i8"), ("v", "<i8"), ("n", "<i8"), ("lr", "<f4"), ("b1", "<f4"),
                 ("b2", "<f4"), ("eps", "<f4"), ("wd", "<f4"), ("bc1", "<f4"), ("bc2", "<f4"), ("pad", "<f4")])
assert _ROW.itemsize == 72


class ClipAdamW(torch.optim.AdamW):
    """torch.optim.AdamW whose step() also performs clip_grad_norm_(all params, max_grad_norm) (train_hybrid.py:913-915)
    on the device, without a host sync. Parameters whose grad is None are skipped, exactly like torch. Every param
    group keeps its own lr / betas / eps / weight_decay and every tensor its own step count (state['step'], mirrored
    on the host so bias correction needs no device read); the clip norm is taken over all groups together, as
    `clip_grad_norm_(model.parameters())` does. `grad_scale` multiplies every gradient before the norm and the update
    (data parallel: 1 / world on all-reduced sums). amsgrad / maximize / capturable / foreach variants are rejected."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.max_grad_norm = float(max_grad_norm)
        self.grad_scale = 1.0
        self.last_grad_norm_sq = None
        self._steps = {}                       # id(param) -> int step count (host mirror of state['step'])
        self._keep = None

    def _host_step(self, p, st):
        t = self._steps.get(id(p))
        if t is None:                          # state came from load_state_dict / add_param_group: one device read
            t = self._steps[id(p)] = int(float(st["step"]))
        return t

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _capi.lib()
        stream = _capi.raw_stream()
        rows, chunks, steps, touched = [], [], [], []
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize"):
                raise _capi.LunarisB200Error("ClipAdamW implements plain AdamW: amsgrad / maximize are not supported")
            if torch.is_tensor(group["lr"]):
                raise _capi.LunarisB200Error("ClipAdamW needs a python-float lr (tensor lr / capturable not supported)")
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    self._steps[id(p)] = 0
                g = p.grad
                if g.is_sparse:
                    raise _capi.LunarisB200Error("ClipAdamW does not support sparse gradients")
                if not g.is_contiguous() or g.dtype != torch.float32:
                    g = g.contiguous().float()
                    p.grad = g
                if not (p.is_contiguous() and p.dtype == torch.float32 and p.is_cuda):
                    raise _capi.LunarisB200Error("ClipAdamW updates contiguous fp32 CUDA parameters only")
                t = self._host_step(p, st) + 1
                self._steps[id(p)] = t
                idx = len(rows)
                rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                             p.numel(), group["lr"], b1, b2, group["eps"], group["weight_decay"],
                             1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), 0.0))
                chunks += [(idx, c) for c in range((p.numel() + _CHUNK - 1) // _CHUNK)]
                steps.append(st["step"])
                touched.append(p)
        if not rows:
            return loss
        dev = touched[0].device
        torch._foreach_add_(steps, 1.0)
        tab = np.array(rows, dtype=_ROW)
        ch = np.array(chunks, dtype=np.int32)
        tab_d = torch.from_numpy(tab.view(np.uint8)).pin_memory().to(dev, non_blocking=True)
        ch_d = torch.from_numpy(ch).pin_memory().to(dev, non_blocking=True)
        norm2 = torch.empty(1024, device=dev, dtype=torch.float32)     # per-block partials of sum (g * scale)^2
        check(lib.lun_multi_grad_sumsq(tab_d.data_ptr(), ch_d.data_ptr(), len(chunks), float(self.grad_scale),
                                       norm2.data_ptr(), stream), "lun_multi_grad_sumsq")
        check(lib.lun_multi_clip_adamw(tab_d.data_ptr(), ch_d.data_ptr(), len(chunks), norm2.data_ptr(),
                                       self.max_grad_norm, float(self.grad_scale), stream), "lun_multi_clip_adamw")
        # the kernels wrote the parameters through raw pointers: tell autograd they changed in place, so every cache
        # keyed on `param._version` (the packed bf16 kernel operands of the drop-in modules) is rebuilt next forward
        torch.autograd.graph.increment_version(touched)
        self.last_grad_norm_sq = norm2
        self._keep = (tab_d, ch_d)                                 # keep the tables alive until the kernels ran
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._steps = {}                                           # re-read lazily from the restored state['step']
