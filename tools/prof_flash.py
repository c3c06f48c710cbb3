"""Driver for an ncu capture of the SelfAttention2d flash kernels: one full wave (148 CTAs): C=512, N=74*128, B=1."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lunaris_orion_b200 import _capi
dev = torch.device("cuda:0")
lib = _capi.lib()
C, N, B = 512, 74 * 128, 1
qk = (torch.randn(B * N, 128, device=dev) * 0.3).to(torch.bfloat16)
v = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
x = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
dy = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
y = torch.empty_like(x); o = torch.empty_like(x); dv = torch.empty_like(x); dqk = torch.empty_like(qk)
lse = torch.empty(B * N, device=dev); dsum = torch.randn(B * N, device=dev) * 0.01
gm = torch.full((1,), 0.5, device=dev)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: ctypes.c_void_p(t.data_ptr())
def run():
    _capi.check(lib.lun_flash_attn2d_bf16(P(qk), P(v), P(x), P(y), P(gm), B, N, C, P(o), P(lse), st), "fwd")
    _capi.check(lib.lun_flash_attn2d_dv_bf16(P(qk), P(dy), P(lse), P(gm), P(dv), B, N, C, st), "dv")
    _capi.check(lib.lun_flash_attn2d_dqk_bf16(P(qk), P(v), P(dy), P(lse), P(dsum), P(gm), P(dqk), B, N, C, st), "dqk")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print("fwd + dV + dQ/dK ms", e0.elapsed_time(e1) / 5)
