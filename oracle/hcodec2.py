"""CPU restatement of the H-Codec-2.0 encode / decode hot path.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Functional PyTorch-CPU code over
a flat state-dict whose keys are the reference's (oracle/weights.py).  Works in
fp32 (the reference's arithmetic) or fp64 ("truth" for error budgets): the
dtype follows the state-dict.  Each function cites the reference lines it
restates; oracle/make_golden.py pins it against the reference's own modules.

Paths are relative to /root/reference/QuarkAudio-HCodec/HCodec-2.0/.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .rvq import rvq_decode, rvq_encode


def _codebooks(sd, name):
    i, cbs = 0, []
    while f"{name}.layers.{i}._codebook.embed" in sd:
        cbs.append(sd[f"{name}.layers.{i}._codebook.embed"][0])
        i += 1
    return torch.stack(cbs, 0)


# --------------------------------------------------------------------------- blocks
def conv1d_same(sd, p, x, stride=1):
    """vq/conv.py:35-57 - zero-pad k//2 both sides (non-causal), then nn.Conv1d."""
    w = sd[p + "conv.weight"]
    k = w.shape[-1]
    x = F.pad(x, (k // 2, k // 2))
    return F.conv1d(x, w, sd.get(p + "conv.bias"), stride=stride)


def layer_norm_c(sd, p, x_btc, eps=1e-6):
    return F.layer_norm(x_btc, (x_btc.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps)


def convnext_block(sd, p, x):
    """vq/conv.py:200-213.  x [B,C,T]."""
    w = sd[p + "dwconv.conv.weight"]
    h = F.conv1d(F.pad(x, (3, 3)), w, sd[p + "dwconv.conv.bias"], groups=w.shape[0])
    h = h.transpose(1, 2)
    h = layer_norm_c(sd, p + "norm.", h)
    h = F.linear(h, sd[p + "pwconv1.linear.weight"], sd[p + "pwconv1.linear.bias"])
    h = F.gelu(h)                                   # nn.GELU() = exact erf form
    h = F.linear(h, sd[p + "pwconv2.linear.weight"], sd[p + "pwconv2.linear.bias"])
    h = sd[p + "gamma"] * h
    return x + h.transpose(1, 2)


def lstm_layer(sd, p, x_btc):
    """nn.LSTM(H,H,1,batch_first=True) (encoder_modules/transformer.py:115,133),
    zero initial state, gate order i,f,g,o; explicit recurrence."""
    w_ih, w_hh = sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"]
    b = sd[p + "bias_ih_l0"] + sd[p + "bias_hh_l0"]
    B, T, H = x_btc.shape
    xp = F.linear(x_btc, w_ih, b)                   # [B,T,4H]
    h = x_btc.new_zeros(B, H)
    c = x_btc.new_zeros(B, H)
    out = []
    for t in range(T):
        g = xp[:, t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, dim=-1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out, 1)


def lstm_layer_aten(sd, p, x_btc):
    """Same layer through ATen's fused LSTM - the kernel the reference calls."""
    H = x_btc.shape[-1]
    m = torch.nn.LSTM(H, H, 1, batch_first=True).to(x_btc.dtype)
    with torch.no_grad():
        for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            getattr(m, n).copy_(sd[p + n])
        return m(x_btc)[0]


def rope_tables(T, head_dim, dtype, theta=10000.0):
    """transformer.py:8-44, 71-74: inv_freq over arange(0,dim,2)/dim, emb=cat(f,f)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = torch.arange(T).float()[:, None] * inv[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def rms_norm(w, x, eps=1e-6):
    """transformer.py:77-96."""
    return F.rms_norm(x, (x.shape[-1],), w, eps)


def transformer_layer(sd, p, x, head_dim=64, aten_lstm=True, num_heads=None):
    """transformer.py:367-393 (layer), :120-182 (attention), :218-226 (MLP).  x [B,T,C]."""
    B, T, C = x.shape
    if num_heads is not None:
        head_dim = C // num_heads
    nh = C // head_dim
    h = rms_norm(sd[p + "input_layernorm.weight"], x)
    h = (lstm_layer_aten if aten_lstm else lstm_layer)(sd, p + "self_attn.rnn.", h)
    q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]).view(B, T, nh, head_dim).transpose(1, 2)
    k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"]).view(B, T, nh, head_dim).transpose(1, 2)
    v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]).view(B, T, nh, head_dim).transpose(1, 2)
    cos, sin = rope_tables(T, head_dim, x.dtype)
    q = q * cos + _rotate_half(q) * sin
    k = k * cos + _rotate_half(k) * sin
    att = torch.matmul(q, k.transpose(2, 3)) * head_dim ** -0.5
    att = F.softmax(att, dim=-1, dtype=torch.float32 if x.dtype == torch.float32 else x.dtype).to(q.dtype)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, T, C)
    x = x + F.linear(o, sd[p + "self_attn.o_proj.weight"])
    h = rms_norm(sd[p + "post_attention_layernorm.weight"], x)
    h = F.linear(F.silu(F.linear(h, sd[p + "mlp.w1.weight"])) * F.linear(h, sd[p + "mlp.w3.weight"]),
                 sd[p + "mlp.w2.weight"])
    return x + h


def transformer(sd, p, x_btc, n_layers, **kw):  # kw: aten_lstm, num_heads / head_dim
    """transformer.py:449-489, non-causal (mask None), no final norm."""
    for i in range(n_layers):
        x_btc = transformer_layer(sd, f"{p}layers.{i}.", x_btc, **kw)
    return x_btc


def resnet_block(sd, p, x):
    """vq/conv.py:265-306 (GroupNorm(32, eps 1e-6) + swish + conv k3, twice; dropout = id in eval)."""
    def gn(name, h):
        return F.group_norm(h, 32, sd[p + name + ".weight"], sd[p + name + ".bias"], 1e-6)
    h = gn("norm1", x); h = h * torch.sigmoid(h); h = conv1d_same(sd, p + "conv1.", h)
    h = gn("norm2", h); h = h * torch.sigmoid(h); h = conv1d_same(sd, p + "conv2.", h)
    return x + h


# --------------------------------------------------------------------------- encoder
def stft_features(sd, cfg, x):
    """vq/codec_encoder.py:65-71.  x [B,T] -> [B, 2*(n_fft/2+1), F]."""
    n_fft, hop = cfg["n_fft"], cfg["hop_length"]
    pad = (n_fft - hop) // 2
    x = F.pad(x, (pad, pad))
    # torchaudio Spectrogram(center=False, power=None): hann window, rfft per frame, no normalisation
    frames = x.unfold(-1, n_fft, hop) * sd["encoder.stft.window"]
    spec = torch.fft.rfft(frames, dim=-1).transpose(1, 2)          # [B, n_freq, F]
    mag = torch.log(torch.clip(spec.abs(), min=1e-5))
    phase = spec.angle() / torch.pi
    return torch.cat([mag, phase], dim=1)


def encoder_forward(sd, cfg, x, taps=None, aten_lstm=True):
    """vq/codec_encoder.py:62-79.  x [B,T] -> emb [B, dimension, N]."""
    p = "encoder."
    h = stft_features(sd, cfg, x)
    if taps is not None: taps["enc.feat"] = h
    h = conv1d_same(sd, p + "embed.", h)
    h = layer_norm_c(sd, p + "norm.", h.transpose(1, 2)).transpose(1, 2)
    if taps is not None: taps["enc.embed_norm"] = h
    for i in range(cfg["convnext_layers"]):
        h = convnext_block(sd, f"{p}prior_net.{i}.", h)
        if taps is not None and i == 0: taps["enc.convnext0"] = h
    if taps is not None: taps["enc.prior"] = h
    h = transformer(sd, p + "post_net.1.", h.transpose(1, 2), cfg["transformer_layers"], aten_lstm=aten_lstm)
    if taps is not None: taps["enc.post"] = h.transpose(1, 2)
    h = layer_norm_c(sd, p + "final_layer_norm.", h).transpose(1, 2)
    stride = int(50 / cfg["target_frame_rate"])
    h = conv1d_same(sd, p + "out.", h, stride=stride)
    if taps is not None: taps["enc.out"] = h
    return h


def semantic_encoder_forward(sd, cfg, feat, taps=None):
    """vq/semantic_module.py:196-201 (Encoder), :122-154 (EncoderBlock), :54-80 (ResidualUnit, ELU)."""
    p = "semantic_encoder."
    h = F.conv1d(feat, sd[p + "conv.conv.weight"], None, padding=1)
    for i, st in enumerate(cfg["strides"]):
        b = f"{p}conv_blocks.{i}."
        for u in (0, 1):
            y = F.conv1d(F.elu(h), sd[b + f"res_units.{u}.conv1.conv.weight"], None, padding=1)
            y = F.conv1d(F.elu(y), sd[b + f"res_units.{u}.conv2.weight"], None)
            h = h + y
        w = sd[b + "conv.conv.weight"]
        k = w.shape[-1]
        h = F.conv1d(h, w, sd[b + "conv.conv.bias"], stride=st, padding=(k - 1) // 2)
        if taps is not None: taps[f"sem.block{i}"] = h
    h = F.conv1d(h, sd[p + "conv2.conv.weight"], None, padding=1)
    if taps is not None: taps["sem.out"] = h
    return h


# --------------------------------------------------------------------------- decoder
def istft_same(window, spec, n_fft, hop):
    """vq/spectral_ops.py:33-75 ('same' padding).  spec complex [B, n_freq, F] -> [B, F*hop]."""
    pad = (n_fft - hop) // 2
    B, N, T = spec.shape
    ifft = torch.fft.irfft(spec, n_fft, dim=1, norm="backward") * window[None, :, None]
    out_size = (T - 1) * hop + n_fft
    y = F.fold(ifft, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop))[:, 0, 0, pad:-pad]
    wsq = window.square().expand(1, T, -1).transpose(1, 2)
    env = F.fold(wsq, output_size=(1, out_size), kernel_size=(1, n_fft), stride=(1, hop)).squeeze()[pad:-pad]
    assert (env > 1e-11).all()
    return y / env


def istft_head(sd, cfg, x_btc):
    """vq/heads.py:41-66."""
    p = "decoder.head."
    h = F.linear(x_btc, sd[p + "out.weight"], sd[p + "out.bias"]).transpose(1, 2)
    mag, ph = h.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    S = mag * (torch.cos(ph) + 1j * torch.sin(ph))
    return istft_same(sd[p + "istft.window"], S, cfg["n_fft"], cfg["hop_length"])


def decoder_forward(sd, cfg, z, taps=None, aten_lstm=True):
    """vq/codec_decoder.py:62-72.  z [B, input_channels, N] -> wav [B, N*hop*factor]."""
    p = "decoder."
    f = int(50 / cfg["target_frame_rate"])
    h = z.repeat_interleave(f, dim=-1)
    h = conv1d_same(sd, p + "embed.", h)
    if taps is not None: taps["dec.embed"] = h
    h = resnet_block(sd, p + "prior_net.0.", h)
    if taps is not None: taps["dec.res0"] = h
    h = resnet_block(sd, p + "prior_net.1.", h)
    h = transformer(sd, p + "prior_net.3.", h.transpose(1, 2), cfg["transformer_layers"], aten_lstm=aten_lstm).transpose(1, 2)
    if taps is not None: taps["dec.tf"] = h
    h = resnet_block(sd, p + "prior_net.5.", h)
    h = resnet_block(sd, p + "prior_net.6.", h)
    h = F.group_norm(h, 32, sd[p + "prior_net.7.weight"], sd[p + "prior_net.7.bias"], 1e-6)
    if taps is not None: taps["dec.prior"] = h
    h = layer_norm_c(sd, p + "norm.", h.transpose(1, 2)).transpose(1, 2)
    for i in range(cfg["convnext_layers"]):
        h = convnext_block(sd, f"{p}post_net.{i}.", h)
    if taps is not None: taps["dec.post"] = h
    h = layer_norm_c(sd, p + "final_layer_norm.", h.transpose(1, 2))
    if taps is not None: taps["dec.final_norm"] = h
    return istft_head(sd, cfg, h)


# --------------------------------------------------------------------------- codec
@torch.no_grad()
def codec_encode(sd, cfg, x, feat, taps=None, aten_lstm=True):
    """vq/codec.py:75-87 -> (acoustic_codes [B,nq,N], semantic_codes [B,nq,N]) int64."""
    emb = encoder_forward(sd, cfg["encoder_config"], x, taps, aten_lstm)
    sem = semantic_encoder_forward(sd, cfg["semantic_encoder_config"], feat, taps)
    out = []
    for e, name in ((emb, "quantizer"), (sem, "semantic_quantizer")):
        B, D, N = e.shape
        idx, _ = rvq_encode(e.transpose(1, 2).reshape(B * N, D).float(), _codebooks(sd, name).float())
        out.append(idx.reshape(B, N, -1).transpose(1, 2))
    return out[0], out[1]


@torch.no_grad()
def codec_dequantize(sd, acoustic_codes, semantic_codes):
    """vq/codec.py:91-97 -> z [B, 2*dim, N]."""
    zs = []
    for codes, name in ((acoustic_codes, "quantizer"), (semantic_codes, "semantic_quantizer")):
        B, nq, N = codes.shape
        cb = _codebooks(sd, name)
        zs.append(rvq_decode(codes.transpose(1, 2).reshape(B * N, nq), cb).reshape(B, N, -1).transpose(1, 2))
    return torch.cat(zs, dim=1)


@torch.no_grad()
def codec_decode(sd, cfg, acoustic_codes, semantic_codes, taps=None, aten_lstm=True):
    """vq/codec.py:89-99 -> wav [B, N*3840]."""
    z = codec_dequantize(sd, acoustic_codes, semantic_codes)
    return decoder_forward(sd, cfg["decoder_config"], z, taps, aten_lstm)


def to_dtype(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
