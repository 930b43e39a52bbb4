"""tcgen05 attention (csrc/attention_umma.cu) against the kernels it replaces, at the shapes of the callers.  CUDA-event timing
(L2 flushed between launches); also the target of the ncu capture in att_profile.sh."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from unified_audio_b200 import ops

dev = torch.device("cuda:0")
SHAPES = [("h15 aggregator (T+G=382, 8 x 64, split)", 64, 382, 8, 64, True),
          ("h15 bottleneck (T=250, 8 x 128, split)", 64, 250, 8, 128, True),
          ("h2 codec transformer (T=500, 24 x 64, single)", 64, 500, 24, 64, False),
          ("h2 accurate / hubert (T=500, 12 x 64, split)", 64, 500, 12, 64, True),
          ("h1.5 decoder transformer (T=500, 8 x 128, single)", 64, 500, 8, 128, False)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
out = []
for name, B, L, H, D, split in SHAPES:
    g = torch.Generator(device=dev).manual_seed(1)
    qkv = torch.randn(B * L, 3 * H * D, device=dev, generator=g)
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
    fr = torch.arange(L).float()[:, None] * inv[None]
    emb = torch.cat([fr, fr], -1)
    cos, sin = emb.cos().to(dev).contiguous(), emb.sin().to(dev).contiguous()
    o = ops.Planes.zeros((B * L, H * D), split, dev)
    ws = torch.zeros(ops.attention_umma_workspace_bytes(B, L, H, D, split), dtype=torch.uint8, device=dev)
    t_new = timed(lambda: ops.attention_umma(qkv, B, L, H, D, cos, sin, o, ws))
    flops = 4.0 * B * H * L * L * D
    row = dict(shape=name, umma_ms=t_new, umma_tflops_algorithmic=flops / t_new / 1e9, passes=3 if split else 1)
    if not quick:
        o2 = ops.Planes.zeros((B * L, H * D), True, dev)
        row["simt_fp32_ms"] = timed(lambda: ops.attention_hd(qkv, B, L, H, D, cos, sin, o2), 3)
        if D == 64 and not split:
            ws2 = torch.zeros(ops.attention_tc_workspace_bytes(B, L, H), dtype=torch.uint8, device=dev)
            row["mma_sync_ms"] = timed(lambda: ops.attention_tc(qkv, B, L, H, cos, sin, o2, ws2))
    out.append(row)
    print(json.dumps(row))
