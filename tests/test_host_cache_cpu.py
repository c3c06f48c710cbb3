"""lunaris_orion_b200._host.cached: the kernel-operand cache keyed by parameter identity (tensors cannot be keys of a
WeakKeyDictionary: key comparison falls back to the elementwise `==`), invalidated by in-place updates and by storage
moves, and emptied when the owner dies."""
import gc

import torch

from lunaris_orion_b200 import _host


def test_cache_hits_invalidates_and_dies_with_the_owner():
    p = torch.nn.Parameter(torch.randn(64))
    q = torch.nn.Parameter(torch.randn(64))
    builds = [0]

    def build():
        builds[0] += 1
        return p.detach().clone()
    n0 = _host.cache_entries()
    a = _host.cached(p, "f32", (p,), build)
    assert _host.cached(p, "f32", (p,), build) is a and builds[0] == 1
    assert _host.cached(q, "f32", (q,), lambda: "q") == "q"          # a second multi-element tensor as a key
    assert _host.cached(p, "other", (p,), lambda: "x") == "x" and builds[0] == 1
    with torch.no_grad():
        p.add_(1.0)                                                   # optimizer-style in-place update: version bump
    assert torch.equal(_host.cached(p, "f32", (p,), build), p.detach()) and builds[0] == 2
    p.data = torch.randn(64)                                          # storage swap without a version bump
    assert torch.equal(_host.cached(p, "f32", (p,), build), p.detach()) and builds[0] == 3
    m = torch.nn.Linear(4, 4)
    assert _host.cached(m, "fold", (m.weight, m.bias), lambda: 7) == 7
    assert _host.cache_entries() == n0 + 3
    del p, q, m, a
    gc.collect()
    assert _host.cache_entries() == n0
