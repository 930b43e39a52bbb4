"""LSTM kernel micro-benchmark (B=64, T=500, H=1536).  QB_LSTM_PROF=1 prints per-phase cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unified_audio_b200 import ops
B, T, H = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 500, 1536
xp = torch.randn(B, T, 4 * H, device="cuda")
whh = ops.Planes.from_f32((torch.rand(4 * H, H, device="cuda") * 2 - 1) / H ** 0.5, False)
out = ops.Planes.zeros((B, T, H), False, "cuda")
ws = torch.zeros(ops.lstm_workspace_bytes(B, H), dtype=torch.uint8, device="cuda")
for _ in range(2): ops.lstm(xp, whh, B, T, H, out, ws)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(3): ops.lstm(xp, whh, B, T, H, out, ws)
e1.record(); torch.cuda.synchronize()
print("lstm ms", e0.elapsed_time(e1) / 3, "us/step", e0.elapsed_time(e1) / 3 / T * 1e3)

U = ops.lstm_tc_units(H)
wp = ops.lstm_tc_permute(whh.hi.float(), U)
ws2 = torch.zeros(ops.lstm_tc_workspace_bytes(B, H), dtype=torch.uint8, device="cuda")
for _ in range(2): ops.lstm_tc(xp, wp, U, B, T, H, out, ws2)
torch.cuda.synchronize(); e0.record()
for _ in range(3): ops.lstm_tc(xp, wp, U, B, T, H, out, ws2)
e1.record(); torch.cuda.synchronize()
print("lstm_tc ms", e0.elapsed_time(e1) / 3, "us/step", e0.elapsed_time(e1) / 3 / T * 1e3)
