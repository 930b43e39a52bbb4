"""Pin oracle/adaptive.py against the reference's own static methods (build container only).

python -m oracle.make_golden_adaptive
FlexiCodec (HCodec-1.5/adaptive/modeling_flexicodec_new.py) imports funasr / dac / easydict / audiotools at module level;
they are stubbed (none is touched by the static methods used here).  Output: tests/golden/adaptive_alignment.npz.
"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = "/root/reference/QuarkAudio-HCodec/HCodec-1.5"


def load_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
    stub("funasr", AutoModel=object)
    stub("easydict", EasyDict=dict)
    stub("dac"); stub("dac.nn"); stub("dac.model")
    stub("dac.nn.layers", WNConv1d=lambda *a, **k: nn.utils.weight_norm(nn.Conv1d(*a, **k)), Snake1d=nn.Identity)
    stub("dac.model.base", CodecMixin=object)
    stub("audiotools", AudioSignal=object)
    stub("audiotools.ml", BaseModel=nn.Module)
    sys.path.insert(0, BASE)
    from adaptive.modeling_flexicodec_new import FlexiCodec
    return FlexiCodec


def main():
    from oracle import adaptive as oa
    Flexi = load_reference()
    g = torch.Generator().manual_seed(4)
    B, T, D = 3, 64, 32
    base = torch.randn(B, T // 4, D, generator=g).repeat_interleave(4, 1)          # runs of similar frames
    h = base + 0.35 * torch.randn(B, T, D, generator=g)
    out = {}
    for thr in (0.6, 0.85):
        ref_a, ref_sim, ref_n = Flexi._perform_similarity_alignment_vectorized(h, x_lens=torch.full((B,), T), current_threshold=thr,
                                                                             max_tokens_per_group=8)
        a, sim, n = oa.similarity_alignment(h, thr, 8)
        assert torch.equal(a, ref_a) and torch.equal(n, ref_n) and torch.allclose(sim, ref_sim)
        lens = oa.token_lengths(a)
        grouped = torch.randn(B, 16, a.shape[1], generator=g) * (lens > 0)[:, None]
        assert torch.equal(oa.deaggregate(grouped, a), Flexi.deaggregate_features(grouped, a, is_channel_last=False))
        assert torch.equal(oa.deaggregate_by_lengths(grouped, lens), Flexi._deaggregate_features_from_token_lengths(grouped, lens))
        out[f"align_{thr}"] = ref_a.numpy()
        print(f"threshold {thr}: groups per item {ref_n.tolist()} of {T} frames")
    codes = torch.randint(0, 1024, (B, 4, a.shape[1]), generator=g)
    packed = oa.inject_lengths(codes, lens.clamp(min=1), 1024)
    plain, l2 = oa.extract_lengths(packed, 1024)
    assert torch.equal(plain, codes) and torch.equal(l2, lens.clamp(min=1))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "adaptive_alignment.npz"), h=h.numpy(), **out)


if __name__ == "__main__":
    main()
