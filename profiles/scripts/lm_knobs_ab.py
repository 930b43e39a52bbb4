"""One timing line for the decode knobs set in the environment (QB_LM_ATT_U, QB_LM_GRAPH_STEPS, QB_LM_LANES, ...): UniSE SR B = 32,
TSE B = 16 (single chain each) and SR B = 256 (lanes), greedy, 283 steps.  python profiles/scripts/lm_knobs_ab.py <tag>"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m = bench.build_lm(dev)
T = 250
res = dict(tag=sys.argv[1] if len(sys.argv) > 1 else "", env={k: v for k, v in os.environ.items() if k.startswith("QB_LM")})
toks = {}
for task, B, reps in (("se", 32, 4), ("tse", 16, 3), ("se", 256, 2)):
    g = torch.Generator().manual_seed(3000)
    mix = torch.randn(B, T, 768, generator=g).to(dev)
    enr = torch.randn(B, T, 768, generator=g).to(dev) if task == "tse" else None
    f = lambda: m.generate(task, enr, enr, mix, mix, do_sample=False)
    out = f()
    f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[f"{task}{B}_ms"] = round(ms, 2)
    res[f"{task}{B}_tok_s"] = round(B * 283 / ms * 1e3)
    toks[f"{task}{B}"] = int(torch.cat(out, 1).sum())          # checksum of the tokens (compare across knob settings)
res["token_checksums"] = toks
print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/lm_knobs_ab.jsonl", "a").write(json.dumps(res) + "\n")
