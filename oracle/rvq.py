"""Oracle restatement of `vector_quantize_pytorch.ResidualVQ` (inference only).

TEST INFRASTRUCTURE - see oracle/__init__.py.  PARITY UNPINNED for this file:
the library (`vector-quantize-pytorch==1.22.15`, pinned at
/root/reference/requirements.txt:54 and HCodec-2.0/requirements.txt:54) is not
vendored in the reference and cannot be installed offline, so no output of the
real package exists to check against.  What this file follows:

  * call sites / IO contract: HCodec-2.0/vq/codec.py:58-61 (input [b,t,d] ->
    (quantized [b,t,d], codes [b,t,nq], commit_loss [nq])), :81-82, :94-95
    (`get_output_from_indices([b,t,nq]) -> [b,t,d]`);
  * ctor kwargs: HCodec-2.0/conf/large_12.5hz_config.yaml:22-29;
  * arithmetic template that IS in the reference tree:
    HCodec-2.0/vq/core_vq.py:223-238 (distance / arg-max of negated distance /
    embedding lookup) and :394-412 (residual loop, indices stacked);
  * published upstream behaviour (lucidrains v1.2x): `x.float()`; Euclidean
    codebook buffer `embed[1,K,D]`; `dist = -sqrt(clamp(|x|^2+|e|^2-2x.e, 0))`;
    eval-mode arg-max (no gumbel noise), first index wins ties; eval-mode
    `quantize` is the raw codebook row; project_in/out are identities because
    codebook_dim == dim; quantize_dropout only acts in training;
    `get_output_from_indices` treats index -1 as "dropped" (zero contribution)
    and sums the per-layer rows over the quantiser axis.

State-dict layout mirrors upstream: `layers.{i}._codebook.{initted, cluster_size,
embed_avg, embed}`.

Partial pin (oracle/make_golden_rvq.py): the reference's IN-TREE residual VQ of the same family,
`HCodec-2.0/vq/core_vq.py` `ResidualVectorQuantization` (lines 223-238, 394-412), loaded with the same seeded
codebooks, returns exactly the indices and de-quantised vectors of `rvq_encode` / `rvq_decode` (150 rows x 4 layers and
80 rows x 16 layers x 4096 codes: 100 % identical); its outputs are committed as tests/golden/rvq_intree.npz.
"""
from __future__ import annotations

import torch
from torch import nn


class _EuclideanCodebook(nn.Module):
    def __init__(self, dim: int, codebook_size: int):
        super().__init__()
        self.register_buffer("initted", torch.tensor([True]))
        self.register_buffer("cluster_size", torch.ones(1, codebook_size))
        self.register_buffer("embed_avg", torch.zeros(1, codebook_size, dim))
        self.register_buffer("embed", torch.zeros(1, codebook_size, dim))


class _VectorQuantize(nn.Module):
    def __init__(self, dim: int, codebook_size: int):
        super().__init__()
        self._codebook = _EuclideanCodebook(dim, codebook_size)


def nearest_code(residual: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """Library-faithful nearest-code search for one layer.

    residual [M,D], embed [K,D] (same dtype).  Mirrors upstream `cdist` +
    eval-mode arg-max (first index on ties); cf. core_vq.py:223-231 for the
    in-tree template without the sqrt/clamp.
    """
    x2 = (residual * residual).sum(-1, keepdim=True)          # [M,1]
    y2 = (embed * embed).sum(-1)[None, :]                     # [1,K]
    xy = residual @ embed.t() * -2.0                          # [M,K]
    dist = -(x2 + y2 + xy).clamp(min=0).sqrt()
    return dist.argmax(dim=-1)


def rvq_encode(x: torch.Tensor, codebooks: torch.Tensor, dtype=torch.float32):
    """x [M,D], codebooks [nq,K,D] -> (indices [M,nq] int64, quantized [M,D]).

    Residual loop of core_vq.py:394-405: residual -= quantized (same dtype).
    """
    r = x.to(dtype)
    cb = codebooks.to(dtype)
    out = torch.zeros_like(r)
    idx = []
    for q in range(cb.shape[0]):
        i = nearest_code(r, cb[q])
        e = cb[q][i]
        r = r - e
        out = out + e
        idx.append(i)
    return torch.stack(idx, dim=-1), out


def rvq_decode(indices: torch.Tensor, codebooks: torch.Tensor) -> torch.Tensor:
    """indices [M,nq] -> sum_q codebooks[q][idx_q], q summed 0..nq-1 in order;
    -1 means dropped (contributes zero)."""
    M, nq = indices.shape
    out = torch.zeros(M, codebooks.shape[-1], dtype=codebooks.dtype)
    for q in range(nq):
        i = indices[:, q]
        e = codebooks[q][i.clamp(min=0)]
        e = torch.where((i < 0)[:, None], torch.zeros_like(e), e)
        out = out + e
    return out


def rvq_margin_audit(x: torch.Tensor, codebooks: torch.Tensor, indices: torch.Tensor):
    """fp64 re-evaluation of every decision ALONG THE GIVEN index path.

    Returns (true_idx [M,nq] int64, rel_margin [M,nq] float64): at each layer
    the fp64 arg-min over squared distances given the fp64 residual that
    follows `indices`, and (d2_second - d2_best) / max(d2_best, tiny).
    A decision is 'numerically safe' when rel_margin >> fp32 epsilon.
    """
    r = x.double()
    cb = codebooks.double()
    true_idx, margins = [], []
    for q in range(cb.shape[0]):
        d2 = (r * r).sum(-1, keepdim=True) + (cb[q] * cb[q]).sum(-1)[None] - 2.0 * (r @ cb[q].t())
        best2 = torch.topk(d2, 2, dim=-1, largest=False)
        true_idx.append(best2.indices[:, 0])
        margins.append((best2.values[:, 1] - best2.values[:, 0]) / best2.values[:, 0].clamp(min=1e-300))
        r = r - cb[q][indices[:, q]]
    return torch.stack(true_idx, -1), torch.stack(margins, -1)


class ResidualVQ(nn.Module):
    """Drop-in (inference) for vector_quantize_pytorch.ResidualVQ as the
    reference constructs it (codec.py:37-43)."""

    def __init__(self, *, dim, codebook_size, num_quantizers, decay=0.99, kmeans_init=False,
                 kmeans_iters=10, quantize_dropout=False, **unused):
        super().__init__()
        self.dim, self.codebook_size, self.num_quantizers = dim, codebook_size, num_quantizers
        self.layers = nn.ModuleList([_VectorQuantize(dim, codebook_size) for _ in range(num_quantizers)])

    def codebooks(self) -> torch.Tensor:
        return torch.stack([l._codebook.embed[0] for l in self.layers], 0)

    @torch.no_grad()
    def forward(self, x):
        assert not self.training, "oracle ResidualVQ restates eval-mode only"
        for l in self.layers:
            assert bool(l._codebook.initted.item()), "codebook not initialised (kmeans trap, SURVEY 7.7)"
        b, t, d = x.shape
        idx, quant = rvq_encode(x.float().reshape(b * t, d), self.codebooks())
        losses = torch.zeros(1, self.num_quantizers)
        return quant.reshape(b, t, d), idx.reshape(b, t, -1), losses

    @torch.no_grad()
    def get_output_from_indices(self, indices):
        b, t, nq = indices.shape
        return rvq_decode(indices.reshape(b * t, nq), self.codebooks()).reshape(b, t, -1)
