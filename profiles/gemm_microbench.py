"""Micro-benchmark / ncu target for the tcgen05 GEMM at the path's dominant shapes.
usage: python profiles/gemm_microbench.py [reps]   (under ncu: -k regex:gemm_tc -s 4 -c 4)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unified_audio_b200 import ops

def bench(name, M, N, K, split, act, reps, residual=False, f32out=False):
    dev = "cuda"
    a = ops.Planes(torch.randn(M, K, device=dev).half(), torch.randn(M, K, device=dev).half() * 1e-3 if split else None)
    w = ops.Planes((torch.randn(N, K, device=dev) * K ** -0.5).half(), (torch.randn(N, K, device=dev) * 1e-4).half() if split else None)
    bias = torch.randn(N, device=dev)
    n_out = N // 2 if act == ops.ACT_SWIGLU else N
    kw = {}
    if f32out:
        o = torch.zeros(M, n_out, device=dev)
        kw["out_f32"] = ops.rowmap(o, n_out, M, 0)
        if residual:
            kw["residual"] = ops.rowmap(o, n_out, M, 0)
            kw["gamma"] = bias
    else:
        op = ops.Planes.zeros((M, n_out), split, dev)
        kw["out_planes"] = op; kw["out_planes_map"] = (n_out, M, 0)
    def run():
        ops.gemm(a, w, N, a_batch=1, a_rows_per_batch=M, a_ld=K, m_per_batch=M, bias=bias, act=act, **kw)
    for _ in range(2): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * M * N * K / ms / 1e9
    r = dict(name=name, M=M, N=N, K=K, split=split, ms=round(ms, 4), algorithmic_tflops=round(tf, 1), issued_tflops=round(tf * (3 if split else 1), 1))
    print(json.dumps(r), flush=True)
    return r

if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    M = 32000
    bench("pwconv1 gelu->f16", M, 4608, 1536, False, ops.ACT_GELU, reps)
    bench("pwconv2 gamma+res->f32", M, 1536, 4608, False, ops.ACT_NONE, reps, residual=True, f32out=True)
    bench("plain ->f16 (no act)", M, 4608, 1536, False, ops.ACT_NONE, reps)
    bench("mlp w13 swiglu split", M, 8192, 1536, True, ops.ACT_SWIGLU, reps)
    bench("mlp w2 split res->f32", M, 1536, 4096, True, ops.ACT_NONE, reps, residual=True, f32out=True)
