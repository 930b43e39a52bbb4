"""H-Codec-1.5 adaptive frame-rate primitives on the device against the oracle (oracle/adaptive.py, pinned exact
against the reference's FlexiCodec static methods by tests/golden/adaptive_alignment.npz)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_similarity_alignment_and_length_packing(lib):
    from oracle import adaptive as oa
    from unified_audio_b200 import adaptive as ga
    g = torch.Generator().manual_seed(3)
    for B, T, D, thr, cap in ((3, 50, 64, 0.6, 8), (2, 200, 512, 0.3, 4), (1, 2, 16, 0.9, 8), (4, 33, 128, -2.0, 3), (2, 40, 32, 2.0, 8)):
        base = torch.randn(B, T, D, generator=g)
        h = base.clone()
        for t in range(1, T):                                 # correlated frames so that both outcomes occur
            h[:, t] = 0.7 * h[:, t - 1] + 0.7 * base[:, t]
        align, sim, ng = oa.similarity_alignment(h, thr, cap)
        a2, s2, n2, lens = ga.similarity_alignment(h.cuda(), thr, cap)
        torch.cuda.synchronize()
        assert float((s2.cpu() - sim).abs().max()) < 1e-5
        safe = bool(((sim - thr).abs() > 1e-5).all())           # a similarity within float noise of the threshold may flip a boundary
        if safe:
            assert torch.equal(n2.cpu(), ng) and torch.equal(a2.cpu(), align)
            assert torch.equal(lens.cpu(), oa.token_lengths(align))
        G = align.shape[1]
        codes = torch.randint(0, 1024, (B, 4, G), generator=g)
        ol = oa.token_lengths(align).clamp(min=1)
        packed = ga.inject_lengths(codes.cuda(), ol.cuda(), 1024)
        assert torch.equal(packed.cpu(), oa.inject_lengths(codes, ol, 1024))
        plain, ln = ga.extract_lengths(packed, 1024)
        op, oln = oa.extract_lengths(oa.inject_lengths(codes, ol, 1024), 1024)
        assert torch.equal(plain.cpu(), op) and torch.equal(ln.cpu(), oln) and torch.equal(plain.cpu(), codes)
        feats = torch.randn(B, 24, G, generator=g)
        tl = oa.token_lengths(align)
        assert torch.equal(ga.deaggregate_by_lengths(feats.cuda(), tl.cuda()).cpu(), oa.deaggregate_by_lengths(feats, tl))
        assert torch.equal(ga.deaggregate_by_lengths(codes.cuda(), tl.cuda()).cpu(), oa.deaggregate_by_lengths(codes, tl))
        assert torch.equal(ga.deaggregate(feats.cuda(), align.cuda()).cpu(), oa.deaggregate(feats, align))
        print(f"[adaptive B={B} T={T} thr={thr} cap={cap}] tokens per clip {ng.tolist()}, compression {T / float(ng.float().mean()):.2f}x")


def test_alignment_against_reference_fixture(lib):
    """the committed fixture holds the alignment matrices of the reference's own FlexiCodec._perform_similarity_alignment_vectorized"""
    from oracle import adaptive as oa
    from unified_audio_b200 import adaptive as ga
    z = np.load(os.path.join(GOLD, "adaptive_alignment.npz"))
    h = torch.from_numpy(z["h"])
    for thr in (0.6, 0.85):
        a, s, n, lens = ga.similarity_alignment(h.cuda(), thr, 8)
        torch.cuda.synchronize()
        _, sim, _ = oa.similarity_alignment(h, thr, 8)
        assert float((s.cpu() - sim).abs().max()) < 1e-5
        if bool(((sim - thr).abs() > 1e-5).all()):
            assert torch.equal(a.cpu(), torch.from_numpy(z[f"align_{thr}"]))
        assert int(lens.max()) <= 8 and bool((lens.sum(1) == h.shape[1]).all())
