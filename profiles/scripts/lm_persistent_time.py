"""time LLM_SFT.generate (SR B=32, 283 steps) with the per-kernel graph path vs the persistent kernel"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
m = bench.build_lm(torch.device("cuda", 0))
mix = torch.randn(32, 250, 768, device="cuda")
for kern in ("tc", "persistent"):
    m.decode_kernel = kern
    for _ in range(2):
        out = m.generate("se", None, None, mix, mix, do_sample=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        out = m.generate("se", None, None, mix, mix, do_sample=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{kern}: {ms:.2f} ms per generate -> {32 * 283 / ms:.1f} k tokens/s", flush=True)
    if kern == "tc": ref = [t.clone() for t in out]
print("identical:", all(torch.equal(a, b) for a, b in zip(ref, out)))
