"""A/B of LLM_SFT.generate's concurrent lanes (QB_LM_LANES x QB_LM_CHUNK): UniSE SR B = 32 / 256, TSE B = 16, greedy, 283 steps.
Prints ms per generate, tokens/s and whether the tokens equal the serial walk's.  python profiles/scripts/lm_lanes_ab.py [quick]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m = bench.build_lm(dev)
m.lane_att_unroll = m.att_unroll        # one summation order everywhere: the token comparison below is exact
T = 250
out = []


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


cases = [("se", 32, [(1, 32), (2, 16), (4, 8), (2, 8)]), ("tse", 16, [(1, 16), (2, 8), (4, 4)]),
         ("se", 256, [(1, 32), (2, 32), (4, 32), (8, 32), (8, 16), (4, 16)])]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    cases = [("se", 32, [(1, 32), (2, 16)]), ("se", 256, [(1, 32), (4, 32)])]
for task, B, combos in cases:
    g = torch.Generator().manual_seed(3000)
    mix = torch.randn(B, T, 768, generator=g).to(dev)
    enr = torch.randn(B, T, 768, generator=g).to(dev) if task == "tse" else None
    ref = None
    for lanes, chunk in combos:
        m.lanes, m.chunk = lanes, chunk
        m._gen_state, m._lane_views = {}, None
        torch.cuda.empty_cache()
        try:
            f = lambda: m.generate(task, enr, enr, mix, mix, do_sample=False)
            got = f()
            f()
            ms = timed(f, 3 if B <= 32 else 2)
            same = None
            if ref is None:
                ref = got
            else:
                same = bool(torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]))
            r = dict(task=task, B=B, lanes=lanes, chunk=chunk, ms=round(ms, 2), tokens_per_s=round(B * 283 / ms * 1e3), same_tokens=same)
        except Exception as e:
            r = dict(task=task, B=B, lanes=lanes, chunk=chunk, error=repr(e))
            torch.cuda.synchronize()
        print(json.dumps(r), flush=True)
        out.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/lm_lanes_ab.json", "w"), indent=1)
