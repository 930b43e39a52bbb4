"""TEST INFRASTRUCTURE - see oracle/__init__.py.

Shared parity helpers for the -m gpu tests, bench.py's `parity` key and __graft_entry__.smoke().

RVQ indices are the north-star's bit-exact gate.  The RVQ kernel itself is exact (tests assert
equality with the oracle on identical inputs).  End to end, the embedding entering the quantiser is
a float tensor that may differ from the reference's by the float tolerance (1e-3 relative), so an
index can legitimately differ only where the reference's own decision is closer than that
perturbation can reach.  `audit_codes` checks exactly that, rigorously:

  for a token whose embedding differs by delta, with identical indices on layers < q, the residual
  entering layer q differs by the same delta (the residual loop subtracts identical code vectors).
  Every Euclidean distance is 1-Lipschitz in its argument, so if the device picked j' where the oracle
  picked j0,   dist(r, e_j') - dist(r, e_j0) <= 2 * |delta|   (r = the oracle's fp64 residual).

A differing index that violates this bound is a real bug and fails the test.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def audit_codes(got, want, x_got_rows, x_ref_rows, codebooks, slack=1e-5):
    """got / want: int64 [B, nq, N] (device or cpu); x_*_rows: [B*N, D] embeddings entering the
    quantiser on our path / the oracle's; codebooks [nq, K, D].
    Returns dict(tokens, tokens_differing, index_match_rate, first_layer_match_rate, worst_excess, worst_gap,
    worst_reach, explained) where `explained` is True iff every first divergence satisfies the Lipschitz bound."""
    B, nq, N = want.shape
    got_r = got.cpu().transpose(1, 2).reshape(B * N, nq)
    want_r = want.cpu().transpose(1, 2).reshape(B * N, nq)
    xr = x_ref_rows.double().cpu()
    delta = (x_got_rows.double().cpu() - xr).norm(dim=-1)              # [M]
    cb = codebooks.double().cpu()
    first_bad = torch.full((B * N,), nq, dtype=torch.long)
    for q in range(nq - 1, -1, -1):
        first_bad[got_r[:, q] != want_r[:, q]] = q
    bad = (first_bad < nq).nonzero().flatten()
    worst_excess, worst_gap, worst_reach = None, None, None
    if len(bad):
        r = xr[bad].clone()
        qb = first_bad[bad]
        gaps = torch.zeros(len(bad), dtype=torch.float64)
        scale = torch.zeros(len(bad), dtype=torch.float64)
        for q in range(nq):
            sel = (qb == q).nonzero().flatten()
            if len(sel):
                rq = r[sel]
                dg = (rq - cb[q][got_r[bad[sel], q]]).norm(dim=-1)
                dw = (rq - cb[q][want_r[bad[sel], q]]).norm(dim=-1)
                gaps[sel] = dg - dw
                scale[sel] = rq.norm(dim=-1) + cb[q][want_r[bad[sel], q]].norm(dim=-1)
            r = r - cb[q][want_r[bad, q]]
        reach = 2.0 * delta[bad] + slack * scale
        excess = gaps - reach
        k = int(excess.argmax())
        worst_excess, worst_gap, worst_reach = float(excess[k]), float(gaps[k]), float(reach[k])
    # all nq decisions of a token count as matching only up to (excluding) its first divergence
    matched = int(first_bad.sum())
    return dict(tokens=B * N, tokens_differing=int(len(bad)), token_match_rate=1.0 - len(bad) / (B * N),
                index_match_rate=matched / float(B * N * nq), first_layer_match_rate=float((first_bad > 0).float().mean()),
                worst_excess=worst_excess, worst_gap=worst_gap, worst_reach=worst_reach,
                max_embedding_delta=float(delta.max()), explained=bool(len(bad) == 0 or worst_excess <= 0.0))  # worst_* are None when nothing differs


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def feat_tap_error(a, b, verbose=True):
    """-> max of: (i) max-norm relative error of the magnitude |S| = exp(log-mag tap) (the suite's yardstick, applied to
    the quantity the DFT computes); (ii) max-rel of the log-magnitude over the well-conditioned bins (|S| >= 1e-2 of the
    largest); (iii) phase error on the circle weighted by |S| / max|S|.  log|S| of a bin 1e-4..1e-5 of the maximum
    amplifies the DFT's ~1e-6 absolute rounding by 1/|S| (the reference's own fp32 FFT is 1.7e-4 off fp64 there), so the raw
    log max-rel is printed with the magnitude of the bin it occurs at, not asserted."""
    nf = b.shape[1] // 2
    la, lb = a[:, :nf].double(), b[:, :nf].double()
    ma, mb = la.exp(), lb.exp()
    e_mag = float((ma - mb).abs().max() / mb.max())
    wc = mb >= 1e-2 * mb.max()
    e_log = float((la - lb).abs()[wc].max() / lb.abs().max())
    d = (a[:, nf:].double() - b[:, nf:].double() + 1) % 2 - 1
    e_ph = float((d.abs() * (mb / mb.max())).max())
    if verbose:
        i = int((la - lb).abs().argmax())
        print(f"    enc.feat: |S| max-rel {e_mag:.2e}; log|S| max-rel on well-conditioned bins {e_log:.2e}; weighted phase {e_ph:.2e}; "
              f"raw log max-rel {rel(la, lb):.2e} at a bin of magnitude {float(mb.flatten()[i]):.2e} (max {float(mb.max()):.2e}); "
              f"bins at the 1e-5 clip floor: {int((mb <= 1.0001e-5).sum())}")
    return max(e_mag, e_log, e_ph)


def phase_wrap_clips(a, b):
    """STFT feature taps [B, 2*nf, F] (ours, oracle) -> (bool [B]: clips where some phase channel differs by a full turn,
    number of wrapped bins, worst on-circle phase difference at the wrapped bins).

    The reference's feature angle(S)/pi (vq/codec_encoder.py:70) has a branch cut at +-1: a bin whose spectrum has re < 0 and
    |im| within the float noise of the FFT (about 1e-6 of the frame's largest magnitude - the reference's own fp32 FFT is that
    far from exact arithmetic) lands on +1 or -1 by the sign of a rounding error.  About one bin in a million does; any
    implementation whose FFT is not bit-identical to the reference's (the reference itself on cuFFT vs its CPU FFT included)
    flips some of them, and a flipped channel moves the embed convolution's input by 2.0, i.e. the whole clip by ~1e-2
    downstream.  Such clips are checked ON THE CIRCLE (the feature is right, the function is discontinuous) and excluded
    from the 1e-3 assertions - the count is reported."""
    nf = b.shape[1] // 2
    d = a[:, nf:].double() - b[:, nf:].double()
    wrapped = d.abs() > 1.0
    circ = ((d + 1) % 2 - 1).abs()
    worst = float(circ[wrapped].max()) if bool(wrapped.any()) else 0.0
    return wrapped.flatten(1).any(1), int(wrapped.sum()), worst
