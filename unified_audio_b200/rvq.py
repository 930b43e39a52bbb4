"""ResidualVQ with the surface the reference uses (third-party `vector_quantize_pytorch.ResidualVQ`,
constructed at QuarkAudio-HCodec/HCodec-2.0/vq/codec.py:37-43, called at :81-82 and :94-95), running on
libquark_b200's fused tensor-core search + exact re-rank (csrc/rvq.cu).  Inference only."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class _EuclideanCodebook(nn.Module):
    def __init__(self, dim, codebook_size):
        super().__init__()
        self.register_buffer("initted", torch.tensor([True]))
        self.register_buffer("cluster_size", torch.ones(1, codebook_size))
        self.register_buffer("embed_avg", torch.zeros(1, codebook_size, dim))
        self.register_buffer("embed", torch.zeros(1, codebook_size, dim))


class _VectorQuantize(nn.Module):
    def __init__(self, dim, codebook_size):
        super().__init__()
        self._codebook = _EuclideanCodebook(dim, codebook_size)


class ResidualVQ(nn.Module):
    """ctor kwargs as in HCodec-2.0/conf/large_12.5hz_config.yaml:22-29 (decay / kmeans_* /
    quantize_dropout only matter in training and are accepted and ignored)."""

    def __init__(self, *, dim, codebook_size, num_quantizers, decay=0.99, kmeans_init=False, kmeans_iters=10,
                 quantize_dropout=False, **unused):
        super().__init__()
        self.dim, self.codebook_size, self.num_quantizers = dim, codebook_size, num_quantizers
        self.layers = nn.ModuleList([_VectorQuantize(dim, codebook_size) for _ in range(num_quantizers)])
        self._prep = None
        self.register_load_state_dict_post_hook(lambda m, k: setattr(m, "_prep", None))
        self.eval()

    def _apply(self, fn, *a, **k):
        self._prep = None
        return super()._apply(fn, *a, **k)

    def set_codebooks(self, cb: torch.Tensor):
        """cb [nq, K, D]"""
        for i, l in enumerate(self.layers):
            l._codebook.embed.copy_(cb[i][None])
            l._codebook.embed_avg.copy_(cb[i][None])
            l._codebook.initted.fill_(True)
        self._prep = None

    def _prepare(self):
        if self._prep is None:
            for l in self.layers:
                if not bool(l._codebook.initted.item()):
                    raise RuntimeError("ResidualVQ codebook not initialised (k-means init only happens in training)")
            cb = torch.stack([l._codebook.embed[0] for l in self.layers], 0).float().contiguous()
            e2 = (cb.double() ** 2).sum(-1)                                        # [nq, K]
            consts = torch.cat([(-0.5 * e2).float().reshape(-1),
                                torch.full((self.codebook_size,), -2.0, device=cb.device)]).contiguous()
            self._prep = dict(cb=cb, planes=ops.Planes.from_f32(cb, True), consts=consts, e2max=float(e2.max()))
        return self._prep

    @torch.no_grad()
    def encode_rows(self, x: torch.Tensor, want_quantized=True):
        """x [M, D] fp32 cuda -> (indices [M, nq] int64, quantized [M, D] or None)"""
        p = self._prepare()
        M, D = x.shape
        x = x.float().contiguous()
        idx = torch.empty(M, self.num_quantizers, dtype=torch.int64, device=x.device)
        quant = torch.empty(M, D, device=x.device) if want_quantized else None
        ws = torch.empty(ops.rvq_workspace_bytes(M, D, self.codebook_size), dtype=torch.uint8, device=x.device)
        ops.rvq_encode(x, p["cb"], p["planes"], p["consts"], p["e2max"], M, D, self.codebook_size, self.num_quantizers,
                       idx, quant, ws)
        return idx, quant

    @torch.no_grad()
    def forward(self, x):
        """[b, t, d] -> (quantized [b,t,d], indices [b,t,nq], commit_loss [1,nq]) as codec.py:58-61 documents."""
        if self.training:
            raise RuntimeError("unified_audio_b200.ResidualVQ is inference-only (call .eval())")
        b, t, d = x.shape
        idx, quant = self.encode_rows(x.reshape(b * t, d))
        return quant.reshape(b, t, d), idx.reshape(b, t, -1), torch.zeros(1, self.num_quantizers, device=x.device)

    @torch.no_grad()
    def decode_rows(self, idx: torch.Tensor, out: torch.Tensor = None, out_ld=None, col_off=0):
        p = self._prepare()
        M = idx.shape[0]
        if out is None:
            out = torch.empty(M, self.dim, device=idx.device)
            out_ld = self.dim
        ops.rvq_decode(idx.contiguous(), p["cb"], M, self.dim, self.codebook_size, self.num_quantizers, out, out_ld,
                       col_off)
        return out

    @torch.no_grad()
    def get_output_from_indices(self, indices):
        b, t, nq = indices.shape
        return self.decode_rows(indices.reshape(b * t, nq).long()).reshape(b, t, self.dim)
