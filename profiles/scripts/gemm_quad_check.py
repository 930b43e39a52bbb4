"""4-CTA multicast GEMM (QB_GEMM_QUAD=1) vs the pair kernel: bit-equality of the outputs, then timing.
Run as two processes (the variant switch is read once per process):  python gemm_quad_check.py dump|check|time"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from unified_audio_b200 import ops

def run_case(M, N, K, act, f32out, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = ops.Planes(torch.randn(M, K, device="cuda", generator=g).half(), None)
    w = ops.Planes((torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).half(), None)
    bias = torch.randn(N, device="cuda", generator=g)
    if f32out:
        o = torch.randn(M, N, device="cuda", generator=g)
        ops.gemm(a, w, N, a_batch=1, a_rows_per_batch=M, a_ld=K, m_per_batch=M, bias=bias, gamma=bias, residual=ops.rowmap(o, N, M, 0),
                 out_f32=ops.rowmap(o, N, M, 0))
        return o
    op = ops.Planes.zeros((M, N), False, "cuda")
    ops.gemm(a, w, N, a_batch=1, a_rows_per_batch=M, a_ld=K, m_per_batch=M, bias=bias, act=act, out_planes=op, out_planes_map=(N, M, 0))
    return op.hi

CASES = [(32000, 4608, 1536, ops.ACT_GELU, False), (32000, 1536, 4608, ops.ACT_NONE, True), (4608, 512, 256, ops.ACT_NONE, True),
         (5000, 1536, 1536, ops.ACT_NONE, False), (16000, 4608, 1536, ops.ACT_GELU, False)]
mode = sys.argv[1]
print("variant:", ops.gemm_kernel_name(32000, 4608, False))
if mode in ("dump", "check"):
    outs = [run_case(*c, seed=10 + i).float().cpu() for i, c in enumerate(CASES)]
    torch.cuda.synchronize()
    path = "/tmp/gemm_quad_ref.pt"
    if mode == "dump":
        torch.save(outs, path)
    else:
        ref = torch.load(path)
        for c, a, b in zip(CASES, outs, ref):
            print(c[:3], "bit-identical" if torch.equal(a, b) else f"DIFF max {float((a - b).abs().max()):.3e} (nan: {int(torch.isnan(a).sum())})")
else:
    sys.argv = [sys.argv[0], "10"]
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gemm_microbench as gm
    for M in (32000, 16000):
        gm.bench(f"pwconv1 gelu->f16 M={M}", M, 4608, 1536, False, ops.ACT_GELU, 10)
        gm.bench(f"pwconv2 gamma+res->f32 M={M}", M, 1536, 4608, False, ops.ACT_NONE, 10, residual=True, f32out=True)
        gm.bench(f"qkv-like 4608x1536 M={M}", M, 4608, 1536, False, ops.ACT_NONE, 10, f32out=True)
