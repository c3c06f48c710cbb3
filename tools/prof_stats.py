"""channel_stats (per-channel sum / sum of squares of a bf16 [P, C] tensor) at step shapes, CUDA events."""
import os, sys, ctypes, torch
sys.path.insert(0, os.getcwd())
from lunaris_orion_b200 import _capi
lib=_capi.lib(); dev=torch.device("cuda:0")
for (B,C) in ((64,512),(64,128),(16,256)):
    x=torch.randn(B,16384,C,device=dev).to(torch.bfloat16); st=torch.zeros(2*C,device=dev)
    s=torch.cuda.current_stream().cuda_stream
    f=lambda: lib.lun_channel_stats_bf16(x.data_ptr(), B*16384, C, st.data_ptr(), s)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/20
    print(f"channel_stats B={B} C={C}: {ms*1e3:.1f} us {x.numel()*2/1e9/ms:.2f} TB/s")
