"""Host-side helpers shared by the two drop-in modules: the kernel-operand cache, the autograd guard of the
inference-only entry points and the dropout trace used by the mask-injection parity tests."""
import weakref

import torch

from ._capi import LunarisB200Error

# ------------------------------------------------------------------------------------------------ operand cache
# bf16 kernel-layout shadows of fp32 parameters. Entries are keyed by the owner's id and guarded by a weak reference
# whose callback drops them, so they die with the model (a WeakKeyDictionary cannot hold tensors: its key comparison
# falls back to `==`, which is elementwise). An entry is valid for one (version counter, storage address, device) of
# every source tensor: optimizer steps bump the version, `.to()` / `p.data = w` change the address.
_cache = {}


def _stamp(tensors):
    return tuple((t._version, t.data_ptr(), str(t.device)) for t in tensors)


def cached(owner, kind, sources, build):
    """build() memoised per (owner, kind) while every tensor of `sources` is unchanged. `owner` is any weak-referenceable
    object whose lifetime bounds the entry (a Parameter or a Module)."""
    key = id(owner)
    ent = _cache.get(key)
    if ent is None or ent[0]() is not owner:              # first use, or a recycled id
        ent = (weakref.ref(owner, lambda _, key=key: _cache.pop(key, None)), {})
        _cache[key] = ent
    slot = ent[1]
    stamp = _stamp(sources)
    hit = slot.get(kind)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    val = build()
    slot[kind] = (stamp, val)
    return val


def cache_entries():
    """Number of live owners (tests: entries disappear when the model is garbage collected)."""
    return len(_cache)


# ------------------------------------------------------------------------------------------------ autograd guard
def require_no_grad(what, *tensors):
    """The sub-module forwards below the two fused training entry points are inference-only: they return tensors
    without a grad_fn. The reference back-propagates through them, so refuse instead of silently cutting the graph."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise LunarisB200Error(
            f"{what} is inference-only in lunaris_orion_b200 (its output carries no autograd graph) but an input "
            "requires grad. Differentiable entry points: LunarisCoreVAE.forward, LunarMoETeacher.forward (training "
            "mode), SelfAttention2d.forward, vae_losses. Call this one under torch.no_grad() / on detached inputs.")


# ------------------------------------------------------------------------------------------------ dropout trace
# tests/test_dropout_parity_gpu.py rebuilds every dropout mask of a step from what the step drew: the seeds of the
# counter-RNG elementwise dropouts and the Dropout2d / head keep-masks. None in production (one `is None` test).
_trace = None


def set_dropout_trace(sink):
    """sink: list-like with .append, or None to switch the trace off. Entries: (kind, tag, value)."""
    global _trace
    _trace = sink


def trace(kind, tag, value):
    if _trace is not None:
        _trace.append((kind, tag, value))


# ------------------------------------------------------------------------------------------------ epsilon source
# LunarisCoreVAE.forward draws its reparameterisation noise with torch.randn on the input's device. The golden
# fixtures were produced by the reference on the CPU generator, whose stream differs from the CUDA generator's for
# the same seed: parity tests install a source that returns the reference's exact draws. None in production.
_eps_source = None


def set_eps_source(fn):
    """fn(batch, latent_dim, device) -> fp32 tensor [batch, latent_dim] on `device`, or None to restore torch.randn."""
    global _eps_source
    _eps_source = fn


def draw_eps(batch, latent_dim, device):
    if _eps_source is not None:
        return _eps_source(batch, latent_dim, device).to(device=device, dtype=torch.float32).contiguous()
    return torch.randn(batch, latent_dim, device=device, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------ zero pool
# A forward or backward pass needs ~200 small zero-initialised fp32 buffers (BatchNorm / GroupNorm sums, pooling
# accumulators, reduction targets of the fused epilogues). Inside `with zero_pool(device):` they are slices of ONE block
# that a single memset clears, instead of one fill launch each; the block is sized from the largest pass seen so far.
_pool = None
_pool_need = {}


class zero_pool:
    def __init__(self, device):
        self.key = str(device)
        self.device = device

    def __enter__(self):
        global _pool
        self.prev = _pool
        self.buf = torch.zeros(_pool_need.get(self.key, 1 << 16), device=self.device, dtype=torch.float32)
        self.off = self.total = 0
        _pool = self
        return self

    def __exit__(self, *exc):
        global _pool
        _pool = self.prev
        if self.total > _pool_need.get(self.key, 0):
            _pool_need[self.key] = self.total + (self.total >> 3)
        return False


def zeros(*shape, device):
    """fp32 zeros of `shape`: a 256-byte aligned slice of the active pool, or a plain torch.zeros outside one."""
    p = _pool
    n = 1
    for d in shape:
        n *= int(d)
    if p is None or str(device) != p.key or n > (1 << 20):
        return torch.zeros(*shape, device=device, dtype=torch.float32)
    n_al = (n + 63) & ~63
    p.total += n_al
    if p.off + n_al > p.buf.numel():
        return torch.zeros(*shape, device=device, dtype=torch.float32)
    v = p.buf[p.off:p.off + n].view(*shape)
    p.off += n_al
    return v
