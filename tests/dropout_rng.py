"""TEST INFRASTRUCTURE: numpy restatement of the stateless counter RNG the elementwise dropouts of the CUDA path use
(lunaris_orion_b200/csrc/elem_common.cuh: hash32 / drop_key / drop_keep1), and the mapping from what a step drew
(lunaris_orion_b200._host dropout trace: seeds of the counter-RNG dropouts, Dropout2d and head keep-masks) to the
`masks` dictionaries of oracle/restatement.py (reference dropout sites lunar_evaluator.py:97,212,225,245,252,359,371).
"""
import numpy as np
import torch

M32 = np.uint64(0xFFFFFFFF)


def hash32(x):
    """lowbias32 on uint32 values held in uint64 arrays."""
    x = x & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & M32
    x ^= x >> np.uint64(16)
    return x


def drop_key(seed, idx8):
    seed = int(seed)
    lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    idx8 = idx8.astype(np.uint64)
    a = ((idx8 & M32) * np.uint64(4)) & M32
    b = (((idx8 >> np.uint64(30)) & M32) * np.uint64(0x9E3779B9)) & M32
    c = (hi * np.uint64(0x85EBCA6B)) & M32
    return lo ^ a ^ b ^ c


def thresh16(p):
    return int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)) if p > 0 else 0


def keep(seed, idx, p):
    """Boolean keep decision of logical element `idx` (uint64 array) for dropout probability p."""
    idx = np.asarray(idx, dtype=np.uint64)
    h = hash32((drop_key(seed, idx >> np.uint64(3)) + ((idx >> np.uint64(1)) & np.uint64(3))) & M32)
    u = np.where(idx & np.uint64(1), h >> np.uint64(16), h & np.uint64(0xFFFF))
    return u >= np.uint64(thresh16(p))


def keep_range(seed, n, p):
    return keep(seed, np.arange(n, dtype=np.uint64), p)


def nhwc_mask_as_nchw(seed, B, H, W, C, p):
    """Keep-mask of an elementwise dropout over an NHWC [B,H*W,C] tensor (element index = linear NHWC offset), returned
    as the NCHW multiplier the reference's nn.Dropout applies: keep / (1 - p)."""
    k = keep_range(seed, B * H * W * C, p).reshape(B, H, W, C)
    return torch.from_numpy(k.astype(np.float32)).permute(0, 3, 1, 2).contiguous() / (1.0 - p)


def attention_mask(seed, B, heads, N, p, chunk=32):
    """attn_drop multiplier [B,heads,nc,32,32] for oracle.restatement.local_attention. The CUDA path evaluates only the
    rows that survive the reference's chunk-index scatter: survivor i < nc is row 0 of chunk i, survivors nc-1 .. nc+30
    are the rows of the last chunk; its RNG index is (((b*nq + i)*8 + head) << 5) | key. Rows that do not survive never
    reach an output or a gradient, their mask entries stay 1."""
    nc = N // chunk
    nq = nc + chunk - 1
    b, i, h, t = np.meshgrid(np.arange(B), np.arange(nq), np.arange(heads), np.arange(chunk), indexing="ij")
    idx = ((((b * nq + i) * heads + h).astype(np.uint64)) << np.uint64(5)) | t.astype(np.uint64)
    k = torch.from_numpy(keep(seed, idx, p).astype(np.float32)) / (1.0 - p)           # [B,nq,heads,32]
    m = torch.ones(B, heads, nc, chunk, chunk)
    m[:, :, :nc, 0, :] = k[:, :nc].permute(0, 2, 1, 3)
    m[:, :, nc - 1, :, :] = k[:, nc - 1:].permute(0, 2, 1, 3)
    return m


def oracle_masks(events, B, H, W, feat, p):
    """events: one Teacher forward's slice of the dropout trace -> masks dict for restatement.teacher_forward."""
    masks = {}
    N = H * W
    for kind, tag, val in events:
        if kind == "seed" and tag == "fe_drop":
            masks[tag] = nhwc_mask_as_nchw(val, B, H, W, 192, p)
        elif kind == "seed" and tag.endswith(".proj_drop"):
            masks[tag] = nhwc_mask_as_nchw(val, B, H, W, feat, p)
        elif kind == "seed" and tag.endswith(".attn_drop"):
            masks[tag] = attention_mask(val, B, 8, N, p)
        elif kind == "seed" and isinstance(val, tuple):        # head dropouts: (seed, B, hidden), index = b * hidden + j
            seed, nb, hidden = val
            k = keep_range(seed, nb * hidden, p).reshape(nb, hidden)
            masks[tag] = torch.from_numpy(k.astype(np.float32)) / (1.0 - p)
        elif kind == "mask2d":
            masks[tag] = (val.detach().cpu().float() / (1.0 - p)).view(B, -1, 1, 1)
        elif kind == "mask":
            masks[tag] = val.detach().cpu().float()
        else:
            raise KeyError((kind, tag))
    return masks
