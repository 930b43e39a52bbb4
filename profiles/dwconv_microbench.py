"""dwconv7 + LayerNorm micro-benchmark (ConvNeXt front half, B=64 x 500 frames x 1536 channels).
Algorithmic bytes: fp32 in (196.6 MB) + fp16 hi plane out (98.3 MB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unified_audio_b200 import ops
B, T, C = 64, 500, 1536
x = torch.randn(B, T, C, device="cuda")
w = torch.randn(C, 7, device="cuda") * 0.3
b = torch.randn(C, device="cuda") * 0.1
lw = torch.rand(C, device="cuda") + 0.5
lb = torch.randn(C, device="cuda") * 0.1
out = ops.Planes.zeros((B * T, C), False, "cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): ops.dwconv7_ln(x, w, b, lw, lb, B, T, C, out)
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.dwconv7_ln(x, w, b, lw, lb, B, T, C, out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"dwconv7_ln {ms*1e3:.1f} us  {(B*T*C*6)/ms/1e6:.0f} GB/s algorithmic")
