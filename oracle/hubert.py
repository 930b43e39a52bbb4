"""CPU restatement of the SSL feature front end of H-Codec-2.0's tokenizer - groundwork for round 2.

TEST INFRASTRUCTURE - see oracle/__init__.py.  No CUDA path exists for this row yet; this file and its fixture are the
parity anchor it will be built against.

    HCodecTokenizer.extract_ssl_features (QuarkAudio-HCodec/HCodec-2.0/audio_tokenizer.py:47-61):
        wavs 48 kHz -> Resample(48k -> 16k) -> pad 160 each side -> HuBERT-base (AutoModel "bosonai/hubert_base",
        output_hidden_states=True) -> mean over the 13 hidden states -> sign(x) * |x| ** 0.3        [B, T50, 768]
    UniSE uses the same recipe with WavLM-base-plus and without the magnitude compression (U/model/model.py:38-51).

HuBERT-base is `transformers.HubertModel` (third-party; the architecture is reproduced here from its published
definition, hubert-base-ls960 configuration): 7 bias-free strided convs (k 10,3,3,3,3,2,2 / s 5,2,2,2,2,2,2; GroupNorm
with one group per channel after the first; GELU) -> LayerNorm -> Linear 512 -> 768 -> + GELU(weight-normed grouped conv k=128,
16 groups, trailing sample removed) -> LayerNorm -> 12 post-LN encoder layers (12 heads x 64, FFN 3072, GELU).
Pinning: oracle/make_golden_hubert.py compares against `transformers.HubertModel` itself (random weights of the same seed,
weights are not available offline) and against `torchaudio.transforms.Resample`.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .weights import _gen

HUBERT_BASE = dict(conv_dim=[512] * 7, conv_kernel=[10, 3, 3, 3, 3, 2, 2], conv_stride=[5, 2, 2, 2, 2, 2, 2], hidden=768,
                   layers=12, heads=12, ffn=3072, pos_k=128, pos_groups=16, eps=1e-5)


def hubert_small():
    return dict(conv_dim=[32] * 7, conv_kernel=[10, 3, 3, 3, 3, 2, 2], conv_stride=[5, 2, 2, 2, 2, 2, 2], hidden=64,
                layers=2, heads=4, ffn=128, pos_k=16, pos_groups=4, eps=1e-5)


def param_specs(c):
    """transformers.HubertModel state-dict keys -> (shape, kind)"""
    out = OrderedDict()
    cin = 1
    for i, (co, k) in enumerate(zip(c["conv_dim"], c["conv_kernel"])):
        out[f"feature_extractor.conv_layers.{i}.conv.weight"] = ((co, cin, k), "w")
        if i == 0:
            out["feature_extractor.conv_layers.0.layer_norm.weight"] = ((co,), "nw")
            out["feature_extractor.conv_layers.0.layer_norm.bias"] = ((co,), "nb")
        cin = co
    H = c["hidden"]
    out["feature_projection.layer_norm.weight"] = ((cin,), "nw"); out["feature_projection.layer_norm.bias"] = ((cin,), "nb")
    out["feature_projection.projection.weight"] = ((H, cin), "w"); out["feature_projection.projection.bias"] = ((H,), "b")
    out["encoder.pos_conv_embed.conv.bias"] = ((H,), "b")
    out["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = ((1, 1, c["pos_k"]), "g")
    out["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = ((H, H // c["pos_groups"], c["pos_k"]), "w")
    out["encoder.layer_norm.weight"] = ((H,), "nw"); out["encoder.layer_norm.bias"] = ((H,), "nb")
    for i in range(c["layers"]):
        p = f"encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            out[p + f"attention.{n}.weight"] = ((H, H), "w"); out[p + f"attention.{n}.bias"] = ((H,), "b")
        out[p + "layer_norm.weight"] = ((H,), "nw"); out[p + "layer_norm.bias"] = ((H,), "nb")
        out[p + "feed_forward.intermediate_dense.weight"] = ((c["ffn"], H), "w")
        out[p + "feed_forward.intermediate_dense.bias"] = ((c["ffn"],), "b")
        out[p + "feed_forward.output_dense.weight"] = ((H, c["ffn"]), "w")
        out[p + "feed_forward.output_dense.bias"] = ((H,), "b")
        out[p + "final_layer_norm.weight"] = ((H,), "nw"); out[p + "final_layer_norm.bias"] = ((H,), "nb")
    return out


def make_state_dict(c, seed=0):
    sd = OrderedDict()
    for name, (shape, kind) in param_specs(c).items():
        g = _gen(seed, name)
        if kind == "w":
            fan = 1
            for v in shape[1:]:
                fan *= v
            sd[name] = torch.randn(shape, generator=g) * (1.5 / fan) ** 0.5
        elif kind == "b":
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "nw":
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "nb":
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "g":
            sd[name] = torch.zeros(shape)
    v = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = v.pow(2).sum((0, 1), keepdim=True).sqrt() * 0.5
    return sd


# --------------------------------------------------------------------------- resampling (torchaudio.transforms.Resample)
def resample_kernel(orig, new, lowpass_filter_width=6, rolloff=0.99):
    """torchaudio.functional._get_sinc_resample_kernel (sinc_interp_hann): [new/g, 1, k], width"""
    g = math.gcd(orig, new)
    orig, new = orig // g, new // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None, None] / new + idx
    t = (t * base).clamp(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels = (kernels * window * scale).float()
    return kernels, width, orig, new


def resample(wav, orig=48000, new=16000):
    """torchaudio.functional._apply_sinc_resample_kernel: wav [B,T] -> [B, ceil(T * new / orig)]"""
    k, width, o, n = resample_kernel(orig, new)
    B, T = wav.shape
    x = F.pad(wav, (width, width + o))
    y = F.conv1d(x[:, None], k, stride=o)                  # [B, n, frames]
    y = y.transpose(1, 2).reshape(B, -1)
    return y[:, : math.ceil(n * T / o)]


# --------------------------------------------------------------------------- HuBERT
def feature_encoder(sd, c, wav):
    """modeling_hubert.HubertFeatureEncoder (feat_extract_norm='group', conv_bias=False): [B,T] -> [B, C, T']"""
    x = wav[:, None]
    for i, s in enumerate(c["conv_stride"]):
        x = F.conv1d(x, sd[f"feature_extractor.conv_layers.{i}.conv.weight"], stride=s)
        if i == 0:
            C = x.shape[1]
            x = F.group_norm(x, C, sd["feature_extractor.conv_layers.0.layer_norm.weight"],
                             sd["feature_extractor.conv_layers.0.layer_norm.bias"], 1e-5)
        x = F.gelu(x)
    return x


def pos_conv_weight(sd):
    """weight_norm parametrization with dim=2: w = g * v / ||v||, norm over (out, in) per kernel tap"""
    g = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    v = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    return v * (g / v.pow(2).sum((0, 1), keepdim=True).sqrt())


def encoder_layer(sd, p, c, x):
    """HubertEncoderLayer (post-LN): x [B,T,H]"""
    B, T, H = x.shape
    h, d = c["heads"], H // c["heads"]
    lin = lambda n, t: F.linear(t, sd[p + f"{n}.weight"], sd[p + f"{n}.bias"])
    q = lin("attention.q_proj", x).view(B, T, h, d).transpose(1, 2)
    k = lin("attention.k_proj", x).view(B, T, h, d).transpose(1, 2)
    v = lin("attention.v_proj", x).view(B, T, h, d).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, -1) @ v
    x = x + lin("attention.out_proj", a.transpose(1, 2).reshape(B, T, H))
    x = F.layer_norm(x, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], c["eps"])
    f = lin("feed_forward.output_dense", F.gelu(lin("feed_forward.intermediate_dense", x)))
    return F.layer_norm(x + f, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], c["eps"])


@torch.no_grad()
def hubert_hidden_states(sd, c, wav):
    """HubertModel(wav, output_hidden_states=True).hidden_states: 1 + layers tensors [B, T', H]"""
    feats = feature_encoder(sd, c, wav).transpose(1, 2)
    Cc = feats.shape[-1]
    x = F.layer_norm(feats, (Cc,), sd["feature_projection.layer_norm.weight"], sd["feature_projection.layer_norm.bias"], c["eps"])
    x = F.linear(x, sd["feature_projection.projection.weight"], sd["feature_projection.projection.bias"])
    pos = F.conv1d(x.transpose(1, 2), pos_conv_weight(sd), sd["encoder.pos_conv_embed.conv.bias"], padding=c["pos_k"] // 2,
                   groups=c["pos_groups"])
    if c["pos_k"] % 2 == 0:
        pos = pos[:, :, :-1]                                # HubertSamePadLayer
    x = x + F.gelu(pos).transpose(1, 2)
    H = x.shape[-1]
    x = F.layer_norm(x, (H,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], c["eps"])
    hs = [x]
    for i in range(c["layers"]):
        x = encoder_layer(sd, f"encoder.layers.{i}.", c, x)
        hs.append(x)
    return hs


@torch.no_grad()
def extract_ssl_features(sd, c, wav48k, compress=True):
    """audio_tokenizer.py:47-61 (compress=True) / U/model/model.py:38-51 (16 kHz input, compress=False)"""
    w = resample(wav48k) if compress else wav48k
    w = F.pad(w, (160, 160))
    mix = torch.stack(hubert_hidden_states(sd, c, w), 1).mean(1)
    if compress:
        mix = ((mix > 0).float() * 2 - 1) * mix.abs() ** 0.3
    return mix


# --------------------------------------------------------------------------- WavLM-base-plus (UniSE, U/model/model.py:30,38-51)
# Same feature encoder / projection / positional conv / post-LN layers as HuBERT-base; the attention adds a gated relative
# position bias (transformers.models.wavlm.modeling_wavlm.WavLMAttention): a bucketed embedding [320, heads] owned by layer 0
# and shared by all layers, scaled per query by a gate computed from the layer input.
WAVLM_BASE_PLUS = dict(HUBERT_BASE, num_buckets=320, max_distance=800)


def wavlm_small():
    return dict(hubert_small(), num_buckets=32, max_distance=80)


def wavlm_param_specs(c):
    out = OrderedDict()
    for k, v in param_specs(c).items():
        out[k] = v
    h, d = c["heads"], c["hidden"] // c["heads"]
    for i in range(c["layers"]):
        p = f"encoder.layers.{i}.attention."
        out[p + "gru_rel_pos_const"] = ((1, h, 1, 1), "nw")
        out[p + "gru_rel_pos_linear.weight"] = ((8, d), "w"); out[p + "gru_rel_pos_linear.bias"] = ((8,), "b")
    out["encoder.layers.0.attention.rel_attn_embed.weight"] = ((c["num_buckets"], h), "emb")
    return out


def wavlm_make_state_dict(c, seed=0):
    sd = make_state_dict(c, seed)
    for name, (shape, kind) in wavlm_param_specs(c).items():
        if name in sd:
            continue
        g = _gen(seed, name)
        if kind == "w":
            sd[name] = torch.randn(shape, generator=g) * (1.5 / shape[-1]) ** 0.5
        elif kind == "b":
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "nw":
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "emb":
            sd[name] = 0.5 * torch.randn(shape, generator=g)
    return sd


def wavlm_position_bias(sd, c, T):
    """WavLMAttention.compute_bias / _relative_positions_bucket -> [heads, T, T]"""
    rel = torch.arange(T)[None, :] - torch.arange(T)[:, None]
    nb = c["num_buckets"] // 2
    bucket = (rel > 0).long() * nb
    a = rel.abs()
    max_exact = nb // 2
    large = torch.log(a.float() / max_exact) / math.log(c["max_distance"] / max_exact) * (nb - max_exact)
    large = torch.min((max_exact + large).long(), torch.full_like(a, nb - 1))
    bucket = bucket + torch.where(a < max_exact, a, large)
    return F.embedding(bucket, sd["encoder.layers.0.attention.rel_attn_embed.weight"]).permute(2, 0, 1)


def wavlm_encoder_layer(sd, p, c, x, pos_bias):
    B, T, H = x.shape
    h, d = c["heads"], H // c["heads"]
    lin = lambda n, t: F.linear(t, sd[p + f"{n}.weight"], sd[p + f"{n}.bias"])
    xh = x.view(B, T, h, d).permute(0, 2, 1, 3)                                        # gate from the layer INPUT, per head
    proj = lin("attention.gru_rel_pos_linear", xh).view(B, h, T, 2, 4).sum(-1)
    ga, gb = torch.sigmoid(proj).chunk(2, dim=-1)
    gate = ga * (gb * sd[p + "attention.gru_rel_pos_const"] - 1.0) + 2.0                # [B,h,T,1]
    bias = gate * pos_bias[None]                                                       # [B,h,T,T]
    q = lin("attention.q_proj", x).view(B, T, h, d).transpose(1, 2)
    k = lin("attention.k_proj", x).view(B, T, h, d).transpose(1, 2)
    v = lin("attention.v_proj", x).view(B, T, h, d).transpose(1, 2)
    a = torch.softmax((q * d ** -0.5) @ k.transpose(-1, -2) + bias, -1) @ v            # F.multi_head_attention_forward
    x = x + lin("attention.out_proj", a.transpose(1, 2).reshape(B, T, H))
    x = F.layer_norm(x, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], c["eps"])
    f = lin("feed_forward.output_dense", F.gelu(lin("feed_forward.intermediate_dense", x)))
    return F.layer_norm(x + f, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], c["eps"])


@torch.no_grad()
def wavlm_hidden_states(sd, c, wav):
    feats = feature_encoder(sd, c, wav).transpose(1, 2)
    Cc = feats.shape[-1]
    x = F.layer_norm(feats, (Cc,), sd["feature_projection.layer_norm.weight"], sd["feature_projection.layer_norm.bias"], c["eps"])
    x = F.linear(x, sd["feature_projection.projection.weight"], sd["feature_projection.projection.bias"])
    pos = F.conv1d(x.transpose(1, 2), pos_conv_weight(sd), sd["encoder.pos_conv_embed.conv.bias"], padding=c["pos_k"] // 2,
                   groups=c["pos_groups"])
    if c["pos_k"] % 2 == 0:
        pos = pos[:, :, :-1]
    x = x + F.gelu(pos).transpose(1, 2)
    H = x.shape[-1]
    x = F.layer_norm(x, (H,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], c["eps"])
    pb = wavlm_position_bias(sd, c, x.shape[1])
    hs = [x]
    for i in range(c["layers"]):
        x = wavlm_encoder_layer(sd, f"encoder.layers.{i}.", c, x, pb)
        hs.append(x)
    return hs


@torch.no_grad()
def extract_semantic_features(sd, c, wav16k):
    """U/model/model.py:38-51: pad 160 each side, mean of the 13 WavLM hidden states (no compression)"""
    return torch.stack(wavlm_hidden_states(sd, c, F.pad(wav16k, (160, 160))), 1).mean(1)
