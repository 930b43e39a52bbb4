"""CPU restatement of the H-Codec-1.5 adaptive frame-rate codec: encode / decode of
QuarkAudio-HCodec/HCodec-1.5/vq/codec_adaptive.py:150-207.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Paths relative to /root/reference/QuarkAudio-HCodec/HCodec-1.5/.

H-Codec-1.5 = the H-Codec-1.0 modules (vq/encoder_modules, vq/codec_decoder.py, vq/semantic_module.py are byte-identical to
HCodec-1.0's - checked with diff) at the widths of conf/config_adaptive_v3.yaml, plus
  * similarity alignment of consecutive 25 Hz frames (oracle/adaptive.py);
  * two `QueryTokenAggregator`s (adaptive/model_blocks/mimi/transformer.py:701-826): one query token per group, initialised to
    the group mean + a learned embedding, interleaved behind its group's last frame; 32 pre-norm transformer layers
    (LayerNorm eps 1e-5, bias-free in_proj / out_proj / linear1 / linear2, exact GELU, LayerScale, interleaved RoPE) over the
    T + G sequence; the outputs at the query positions are the tokens.  `causal=False` makes `attn_bias` None
    (transformer.py:403-415): attention is FULL over the T + G positions - the `context` argument is inert, and the padded
    query slots of shorter items (all equal to the learned embedding) are attended like any other position;
  * a 32-layer `ProjectedTransformer` bottleneck (d_model 1024, 8 heads of 128) between de-aggregation and the decoder;
  * token lengths packed into the indices (oracle/adaptive.py).
Pinned by oracle/make_golden_h15.py against the reference's own modules (tests/golden/h15_small.npz, h15_pinning_report.json).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import adaptive as ad
from . import hcodec1 as h1
from . import hcodec2 as h2
from .rvq import rvq_encode

AGG = dict(dim=512, heads=8, layers=32, ff=2048)
H15 = dict(h1.H1, ratios=[8, 5, 4, 2], dec_dim=1024, sem_in=1024, sem_ch=1024,
           agg=AGG, bottleneck=dict(dim=1024, heads=8, layers=32, ff=2048), threshold=0.6, max_group=8, layer_scale=0.01)


def h15_shallow():
    """the shipped widths (the reference hard-codes 8 heads in both LSTM-transformers, vq/encoder_modules/seanet.py:165-172 and
    vq/codec_decoder.py:41-48, so head_dim 64 / 128 needs dimension 512 / dec_dim 1024) with fewer layers - fast fixtures"""
    return dict(H15, dec_layers=2, dec_inter=768, agg=dict(AGG, layers=3), bottleneck=dict(H15["bottleneck"], layers=2))


# --------------------------------------------------------------------------- parameter layout
def mimi_specs(out, prefix, t):
    """state_dict of ProjectedTransformer(d_model == input_dimension == output_dimensions[0]): no projections
    (transformer.py:855-866), layers of StreamingTransformerLayer (transformer.py:459-548)."""
    d, ff = t["dim"], t["ff"]
    for i in range(t["layers"]):
        p = f"{prefix}transformer.layers.{i}."
        out[p + "self_attn.in_proj_weight"] = ((3 * d, d), "w")
        out[p + "self_attn.out_proj.weight"] = ((d, d), "w")
        out[p + "norm1.weight"] = ((d,), "nw")
        out[p + "norm1.bias"] = ((d,), "nb")
        out[p + "norm2.weight"] = ((d,), "nw")
        out[p + "norm2.bias"] = ((d,), "nb")
        out[p + "linear1.weight"] = ((ff, d), "w")
        out[p + "linear2.weight"] = ((d, ff), "w")
        out[p + "layer_scale_1.scale"] = ((d,), "ls")
        out[p + "layer_scale_2.scale"] = ((d,), "ls")


def param_specs(c):
    out = h1.param_specs(c)
    for name in ("semantic_aggregator", "acoustic_aggregator"):
        out[f"{name}.query_embedding"] = ((1, c["agg"]["dim"], 1), "q")
        mimi_specs(out, f"{name}.transformer.", c["agg"])
    mimi_specs(out, "bottleneck_transformer.", c["bottleneck"])
    return out


def make_state_dict(c, seed=0):
    """Seeded weights.  LayerScale is drawn around 0.35 rather than the 0.01 initial value so that the 96 transformer layers
    move the signal by O(1) - a trained checkpoint's scales are not small, and parity on a near-identity stack would test nothing."""
    return h1.make_state_dict(c, seed, param_specs(c))


# --------------------------------------------------------------------------- mimi transformer
def rope_interleaved(q, k, max_period=10000.0):
    """module/rope.py:12-70, offset 0, [B,H,T,D]: pairs (2i, 2i+1) rotated by t * max_period^(-2i/D), computed in fp32"""
    B, H, T, D = q.shape
    freqs = torch.exp(torch.arange(D // 2, dtype=torch.float32) * (-math.log(max_period) * 2 / D))
    ang = torch.arange(T, dtype=torch.float32).view(1, 1, T, 1) * freqs
    cr, sr = torch.cos(ang), torch.sin(ang)

    def rot(x):
        x = x.view(B, H, T, D // 2, 2)
        r, i = x[..., 0].float(), x[..., 1].float()
        return torch.stack([r * cr - i * sr, r * sr + i * cr], -1).view(B, H, T, D)

    return rot(q), rot(k)


def mimi_layer(sd, p, x, heads):
    """StreamingTransformerLayer.forward (transformer.py:553-596), no streaming state, causal False -> no mask."""
    B, T, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    qkv = F.linear(h, sd[p + "self_attn.in_proj_weight"]).view(B, T, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k = rope_interleaved(qkv[0], qkv[1])
    a = F.scaled_dot_product_attention(q, k, qkv[2]).transpose(1, 2).reshape(B, T, C)
    x = x + sd[p + "layer_scale_1.scale"] * F.linear(a, sd[p + "self_attn.out_proj.weight"])
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    u = F.linear(F.gelu(F.linear(h, sd[p + "linear1.weight"])), sd[p + "linear2.weight"])
    return x + sd[p + "layer_scale_2.scale"] * u


def projected_transformer(sd, prefix, x, t, taps=None, tap_name=None):
    """ProjectedTransformer.forward with conv_layout=True (transformer.py:867-880): [B,C,T] -> [B,C,T]"""
    x = x.transpose(1, 2)
    for i in range(t["layers"]):
        x = mimi_layer(sd, f"{prefix}transformer.layers.{i}.", x, t["heads"])
        if taps is not None and tap_name and i in (0, t["layers"] - 1):
            taps[f"{tap_name}.layer{i}"] = x.transpose(1, 2)
    return x.transpose(1, 2)


def interleave_plan(align, n_groups):
    """transformer.py:760-805 restated as index arithmetic.  align [B,G,T] 0/1 with contiguous groups, all T frames valid.
    Returns (frame_pos [B,T], query_pos [B,G], group_mask [B,G]): interleaved position of frame t = t + (groups ended before t),
    of the query of group g = last_frame(g) + g + 1; padded groups fill the tail T + n_b ... T + G - 1 (their contents are
    identical, so the unstable argsort of the reference cannot matter)."""
    B, G, T = align.shape
    gmask = torch.arange(G)[None] < n_groups[:, None]
    seg = align.argmax(1)                                                       # frame -> group
    frame_pos = torch.arange(T)[None] + seg
    last = (align * torch.arange(T)).amax(2).long()
    query_pos = last + torch.arange(G)[None] + 1
    query_pos = torch.where(gmask, query_pos, T + torch.arange(G)[None])        # padded g: behind the T + n_b real positions
    return frame_pos, query_pos, gmask


def query_token_aggregator(sd, prefix, feats, align, n_groups, t, taps=None, tap_name=None):
    """QueryTokenAggregator.forward (transformer.py:740-826): feats [B,D,T], align [B,G,T] -> tokens [B,D,G]"""
    B, D, T = feats.shape
    G = align.shape[1]
    counts = align.sum(2).clamp(min=1)
    queries = (torch.einsum("bgt,bdt->bgd", align, feats) / counts[..., None]) + sd[prefix + "query_embedding"][:, :, 0][:, None]
    fpos, qpos, gmask = interleave_plan(align, n_groups)
    seq = torch.zeros(B, T + G, D)
    bi = torch.arange(B)[:, None]
    seq[bi, fpos] = feats.transpose(1, 2)
    seq[bi, qpos] = queries
    if taps is not None and tap_name:
        taps[f"{tap_name}.interleaved"] = seq.transpose(1, 2)
    out = projected_transformer(sd, prefix + "transformer.", seq.transpose(1, 2), t, taps, tap_name).transpose(1, 2)
    tok = out[bi, qpos] * gmask[..., None]
    return tok.transpose(1, 2)


# --------------------------------------------------------------------------- codec
def _rvq(sd, name, e):
    B, D, N = e.shape
    idx, _ = rvq_encode(e.transpose(1, 2).reshape(B * N, D).float(), h2._codebooks(sd, name).float())
    return idx.reshape(B, N, -1).transpose(1, 2)


@torch.no_grad()
def codec_encode(sd, c, x, feat, taps=None, threshold=None):
    """codec_adaptive.py:150-183: x [B,1,T] @16 kHz, feat [B,sem_in,T/320] -> length-packed codes [B,nq,G] x2"""
    emb = h1.seanet_encoder(sd, c, x, taps)
    sem = h1.semantic_encoder(sd, c, feat, taps)
    if taps is not None: taps["sem.out"] = sem
    thr = c["threshold"] if threshold is None else threshold
    align, _, n = ad.similarity_alignment(sem.transpose(1, 2), thr, c["max_group"])
    if taps is not None:
        taps["align"], taps["n_groups"] = align, n
    sem_tok = query_token_aggregator(sd, "semantic_aggregator.", sem, align, n, c["agg"], taps, "sem_agg")
    ac_tok = query_token_aggregator(sd, "acoustic_aggregator.", emb, align, n, c["agg"], taps, "ac_agg")
    if taps is not None:
        taps["sem_agg.out"], taps["ac_agg.out"] = sem_tok, ac_tok
    lens = ad.token_lengths(align)
    ac = ad.inject_lengths(_rvq(sd, "quantizer", ac_tok), lens, c["codebook_size"])
    sc = ad.inject_lengths(_rvq(sd, "semantic_quantizer", sem_tok), lens, c["codebook_size"])
    return ac, sc


@torch.no_grad()
def codec_decode(sd, c, ac, sc, taps=None):
    """codec_adaptive.py:186-207 with token_lengths=None: lengths come out of the indices"""
    ac, lens = ad.extract_lengths(ac, c["codebook_size"])
    sc, lens = ad.extract_lengths(sc, c["codebook_size"])
    ac = ad.deaggregate_by_lengths(ac, lens)
    sc = ad.deaggregate_by_lengths(sc, lens)
    z = h2.codec_dequantize(sd, ac, sc)
    if taps is not None: taps["dec.z"] = z
    z = projected_transformer(sd, "bottleneck_transformer.", z, c["bottleneck"], taps, "bottleneck")
    if taps is not None: taps["bottleneck.out"] = z
    return h1.decoder(sd, c, z, taps)
