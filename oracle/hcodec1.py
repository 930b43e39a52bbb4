"""CPU restatement of the H-Codec-1.0 encode / decode hot path (BASELINE.json configs[0]).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Paths relative to /root/reference/QuarkAudio-HCodec/HCodec-1.0/.
H-Codec-1.0 hard-codes its hyper-parameters (vq/codec.py:30-136): SEANet encoder (ratios 2,4,5,8, 32 base
filters, weight-normed reflect-padded convs, conv shortcuts, ELU) -> Transformer(512, 8 heads, LSTM) -> ELU ->
conv k4 s2; 2 x ResidualVQ(4 x 1024 x 512); decoder: sub-pixel x2 up-sampler -> ResNet/Transformer(768, 8 heads of
96) -> 12 ConvNeXt(768/2304) -> ISTFT(1280, hop 320).  Transformer / ConvNeXt / ResNet / ISTFT / semantic encoder
code is identical to H-Codec-2.0's (verified by diff) and is shared with oracle/hcodec2.py.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import hcodec2 as h2
from .rvq import rvq_decode, rvq_encode

RATIOS = [2, 4, 5, 8]          # SEANetEncoder reverses ratios=[8,5,4,2] (seanet.py:111)
H1 = dict(n_filters=32, dimension=512, dec_dim=768, dec_inter=2304, dec_layers=12, n_fft=1280, hop=320, nq=4,
          codebook_size=1024, sem_in=768, sem_ch=768, sem_strides=[2, 1], tf_layers=2, heads=8)


def h1_small():
    """reduced-width variant with the same topology (for fast fixtures)"""
    return dict(H1, n_filters=16, dimension=256, dec_dim=384, dec_inter=768, dec_layers=2, nq=3, codebook_size=128,
                sem_ch=256, heads=4)


# --------------------------------------------------------------------------- parameter layout
def param_specs(c):
    out = OrderedDict()

    def wn_conv(p, cout, cin, k):
        out[p + "conv.conv.bias"] = ((cout,), "b")
        out[p + "conv.conv.weight_g"] = ((cout, 1, 1), "g")
        out[p + "conv.conv.weight_v"] = ((cout, cin, k), "w")

    nf, dim = c["n_filters"], c["dimension"]
    wn_conv("encoder.model.0.", nf, 1, 7)
    mult, idx = 1, 1
    for r in c.get("ratios", RATIOS):
        ch = mult * nf
        wn_conv(f"encoder.model.{idx}.block.1.", ch // 2, ch, 3)
        wn_conv(f"encoder.model.{idx}.block.3.", ch, ch // 2, 1)
        wn_conv(f"encoder.model.{idx}.shortcut.", ch, ch, 1)
        wn_conv(f"encoder.model.{idx + 2}.", ch * 2, ch, 2 * r)
        mult *= 2
        idx += 3
    assert mult * nf == dim
    _tf(out, f"encoder.model.{idx + 1}.", dim, dim * 4, c["tf_layers"])
    wn_conv(f"encoder.model.{idx + 4}.", dim, dim, 4)
    dd, di = c["dec_dim"], c["dec_inter"]
    out["decoder.embed.up.weight"] = ((dd * 2, 2 * dim, 1), "w")
    out["decoder.embed.up.bias"] = ((dd * 2,), "b")
    out["decoder.embed.dw.weight"] = ((dd, 1, 5), "w")
    out["decoder.embed.dw.bias"] = ((dd,), "b")
    out["decoder.norm.weight"] = ((dd,), "nw")
    out["decoder.norm.bias"] = ((dd,), "nb")
    for i in range(c["dec_layers"]):
        p = f"decoder.post_net.{i}."
        out[p + "gamma"] = ((dd,), "gamma")
        out[p + "dwconv.conv.weight"] = ((dd, 1, 7), "w")
        out[p + "dwconv.conv.bias"] = ((dd,), "b")
        out[p + "norm.weight"] = ((dd,), "nw")
        out[p + "norm.bias"] = ((dd,), "nb")
        out[p + "pwconv1.linear.weight"] = ((di, dd), "w")
        out[p + "pwconv1.linear.bias"] = ((di,), "b")
        out[p + "pwconv2.linear.weight"] = ((dd, di), "w")
        out[p + "pwconv2.linear.bias"] = ((dd,), "b")
    out["decoder.final_layer_norm.weight"] = ((dd,), "nw")
    out["decoder.final_layer_norm.bias"] = ((dd,), "nb")
    for i in (0, 1, 5, 6):
        p = f"decoder.prior_net.{i}."
        for j in (1, 2):
            out[p + f"norm{j}.weight"] = ((dd,), "nw")
            out[p + f"norm{j}.bias"] = ((dd,), "nb")
            out[p + f"conv{j}.conv.weight"] = ((dd, dd, 3), "w")
            out[p + f"conv{j}.conv.bias"] = ((dd,), "b")
    _tf(out, "decoder.prior_net.3.", dd, dd * 4, c["tf_layers"])
    out["decoder.prior_net.7.weight"] = ((dd,), "nw")
    out["decoder.prior_net.7.bias"] = ((dd,), "nb")
    out["decoder.head.out.weight"] = ((c["n_fft"] + 2, dd), "w")
    out["decoder.head.out.bias"] = ((c["n_fft"] + 2,), "b")
    out["decoder.head.istft.window"] = ((c["n_fft"],), "hann")
    for name in ("quantizer", "semantic_quantizer"):
        for i in range(c["nq"]):
            p = f"{name}.layers.{i}._codebook."
            out[p + "initted"] = ((1,), "true")
            out[p + "cluster_size"] = ((1, c["codebook_size"]), "ones")
            out[p + "embed_avg"] = ((1, c["codebook_size"], dim), "embed")
            out[p + "embed"] = ((1, c["codebook_size"], dim), "embed")
    sc = c["sem_ch"]
    out["semantic_encoder.conv.conv.weight"] = ((sc, c["sem_in"], 3), "w")
    for i, st in enumerate(c["sem_strides"]):
        p = f"semantic_encoder.conv_blocks.{i}."
        for u in (0, 1):
            out[p + f"res_units.{u}.conv1.conv.weight"] = ((sc, sc, 3), "w")
            out[p + f"res_units.{u}.conv2.weight"] = ((sc, sc, 1), "w")
        k = 3 if st == 1 else 2 * st
        out[p + "conv.conv.weight"] = ((sc, sc, k), "w")
        out[p + "conv.conv.bias"] = ((sc,), "b")
    out["semantic_encoder.conv2.conv.weight"] = ((dim, sc, 3), "w")
    return out


def _tf(out, prefix, dim, inter, layers):
    for i in range(layers):
        p = f"{prefix}layers.{i}."
        for n in ("weight_ih_l0", "weight_hh_l0"):
            out[p + "self_attn.rnn." + n] = ((4 * dim, dim), "lstm")
        for n in ("bias_ih_l0", "bias_hh_l0"):
            out[p + "self_attn.rnn." + n] = ((4 * dim,), "lstm")
        for n in "qkv":
            out[p + f"self_attn.{n}_proj.weight"] = ((dim, dim), "w")
            out[p + f"self_attn.{n}_proj.bias"] = ((dim,), "b")
        out[p + "self_attn.o_proj.weight"] = ((dim, dim), "w")
        out[p + "mlp.w1.weight"] = ((inter, dim), "w")
        out[p + "mlp.w2.weight"] = ((dim, inter), "w")
        out[p + "mlp.w3.weight"] = ((inter, dim), "w")
        out[p + "input_layernorm.weight"] = ((dim,), "nw")
        out[p + "post_attention_layernorm.weight"] = ((dim,), "nw")


def make_state_dict(c, seed=0, specs=None):
    sd = OrderedDict()
    for name, (shape, kind) in (specs or param_specs(c)).items():
        key = name.replace("embed_avg", "embed")
        g = torch.Generator()
        g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2**63 - 1))
        if kind == "w":
            fan = 1
            for s_ in shape[1:]:
                fan *= s_
            sd[name] = torch.randn(shape, generator=g) * fan ** -0.5
        elif kind == "g":
            sd[name] = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif kind == "b":
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "nw":
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "nb":
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "gamma":
            sd[name] = (1.0 / c["dec_layers"]) * (1 + 0.2 * torch.randn(shape, generator=g))
        elif kind == "lstm":
            k = shape[-1] ** -0.5 if len(shape) == 2 else (shape[0] // 4) ** -0.5
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * k
        elif kind == "hann":
            sd[name] = torch.hann_window(shape[0])
        elif kind == "true":
            sd[name] = torch.tensor([True])
        elif kind == "ones":
            sd[name] = torch.ones(shape)
        elif kind == "embed":
            layer = int(name.split(".layers.")[1].split(".")[0])
            sd[name] = torch.randn(shape, generator=g) * (0.35 * 0.85 ** layer)
        elif kind == "ls":            # LayerScale of the H-Codec-1.5 transformers (oracle/hcodec15.py)
            sd[name] = 0.35 * (1 + 0.2 * torch.randn(shape, generator=g))
        elif kind == "q":
            sd[name] = 0.5 * torch.randn(shape, generator=g)
        else:
            raise KeyError(kind)
    return sd


# --------------------------------------------------------------------------- forward
def wn_weight(sd, p):
    """old-style torch.nn.utils.weight_norm (encoder_modules/conv.py:25-28): w = g * v / ||v||_(1,2)"""
    v = sd[p + "conv.conv.weight_v"]
    return sd[p + "conv.conv.weight_g"] * v / v.flatten(1).norm(dim=1)[:, None, None]


def sconv1d(sd, p, x, stride=1):
    """encoder_modules/conv.py:175-211 non-causal: reflect pad (left = total - total//2, right = total//2 + extra)."""
    w = wn_weight(sd, p)
    k = w.shape[-1]
    total = k - stride
    T = x.shape[-1]
    n_frames = (T - k + total) / stride + 1
    import math
    extra = (math.ceil(n_frames) - 1) * stride + (k - total) - T
    right = total // 2
    left = total - right
    if left or right or extra:
        assert T > max(left, right + extra), "short inputs need the zero-extension path of pad1d (conv.py:88-94)"
        x = F.pad(x, (left, right + extra), mode="reflect")
    return F.conv1d(x, w, sd[p + "conv.conv.bias"], stride=stride)


def seanet_encoder(sd, c, x, taps=None, aten_lstm=True):
    """encoder_modules/seanet.py:121-208 as instantiated at vq/codec.py:30-35.  x [B,1,T] -> [B,dimension,T/640]."""
    h = sconv1d(sd, "encoder.model.0.", x)
    idx = 1
    for r in c.get("ratios", RATIOS):
        p = f"encoder.model.{idx}."
        y = sconv1d(sd, p + "block.1.", F.elu(h))
        y = sconv1d(sd, p + "block.3.", F.elu(y))
        h = sconv1d(sd, p + "shortcut.", h) + y
        h = sconv1d(sd, f"encoder.model.{idx + 2}.", F.elu(h), stride=r)
        if taps is not None: taps[f"enc.down{r}"] = h
        idx += 3
    h = h2.transformer(sd, f"encoder.model.{idx + 1}.", h.transpose(1, 2), c["tf_layers"], aten_lstm=aten_lstm,
                       num_heads=c["heads"]).transpose(1, 2)
    if taps is not None: taps["enc.tf"] = h
    h = sconv1d(sd, f"encoder.model.{idx + 4}.", F.elu(h), stride=2)
    if taps is not None: taps["enc.out"] = h
    return h


def semantic_encoder(sd, c, feat, taps=None):
    return h2.semantic_encoder_forward(sd, dict(strides=c["sem_strides"]), feat, taps)


def decoder(sd, c, z, taps=None, aten_lstm=True):
    """vq/codec_decoder.py:54-66; sub-pixel up-sampler vq/conv.py:60-93 (stride 2, kernel 5)."""
    p = "decoder."
    dd = c["dec_dim"]
    h = F.conv1d(z, sd[p + "embed.up.weight"], sd[p + "embed.up.bias"])
    h = h.unflatten(1, (2, dd)).permute(0, 2, 3, 1).flatten(-2, -1)              # (B, D, T*2)
    h = F.conv1d(F.pad(h, (2, 2)), sd[p + "embed.dw.weight"], sd[p + "embed.dw.bias"], groups=dd)
    if taps is not None: taps["dec.embed"] = h
    h = h2.resnet_block(sd, p + "prior_net.0.", h)
    h = h2.resnet_block(sd, p + "prior_net.1.", h)
    h = h2.transformer(sd, p + "prior_net.3.", h.transpose(1, 2), c["tf_layers"], aten_lstm=aten_lstm,
                       num_heads=c["heads"]).transpose(1, 2)
    if taps is not None: taps["dec.tf"] = h
    h = h2.resnet_block(sd, p + "prior_net.5.", h)
    h = h2.resnet_block(sd, p + "prior_net.6.", h)
    h = F.group_norm(h, 32, sd[p + "prior_net.7.weight"], sd[p + "prior_net.7.bias"], 1e-6)
    h = h2.layer_norm_c(sd, p + "norm.", h.transpose(1, 2)).transpose(1, 2)
    for i in range(c["dec_layers"]):
        h = h2.convnext_block(sd, f"{p}post_net.{i}.", h)
    if taps is not None: taps["dec.post"] = h
    h = h2.layer_norm_c(sd, p + "final_layer_norm.", h.transpose(1, 2))
    return h2.istft_head(sd, dict(n_fft=c["n_fft"], hop_length=c["hop"]), h)


@torch.no_grad()
def codec_encode(sd, c, x, feat, taps=None, aten_lstm=True):
    """vq/codec.py:167-176: x [B,1,T] (16 kHz), feat [B,768,T/320] -> codes [B,nq,N] x2."""
    emb = seanet_encoder(sd, c, x, taps, aten_lstm)
    sem = semantic_encoder(sd, c, feat, taps)
    out = []
    for e, name in ((emb, "quantizer"), (sem, "semantic_quantizer")):
        B, D, N = e.shape
        idx, _ = rvq_encode(e.transpose(1, 2).reshape(B * N, D).float(), h2._codebooks(sd, name).float())
        out.append(idx.reshape(B, N, -1).transpose(1, 2))
    return out[0], out[1]


@torch.no_grad()
def codec_decode(sd, c, ac, sc, taps=None, aten_lstm=True):
    """vq/codec.py:179-188."""
    return decoder(sd, c, h2.codec_dequantize(sd, ac, sc), taps, aten_lstm)
