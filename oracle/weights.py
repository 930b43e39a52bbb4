"""Seeded synthetic state-dicts for the H-Codec-2.0 hot path.

TEST INFRASTRUCTURE - see oracle/__init__.py.  No pretrained weights ship with
the reference, so parity runs on seeded random weights.  Every
tensor is drawn from its own generator (seed = f(base_seed, crc32(name))), so a
state-dict is reproducible tensor-by-tensor on any box with the same torch
build, independent of module construction order.

Key names and shapes restate the reference modules:
  encoder.*            HCodec-2.0/vq/codec_encoder.py:30-58
  decoder.*            HCodec-2.0/vq/codec_decoder.py:30-59
  semantic_encoder.*   HCodec-2.0/vq/semantic_module.py:157-194
  quantizer.* / semantic_quantizer.*   upstream ResidualVQ buffers (oracle/rvq.py)
They are checked against the reference's own `state_dict()` by
oracle/make_golden.py (fixture tests/golden/h2_keys_*.json).
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import torch

# HCodec-2.0/conf/large_12.5hz_config.yaml
H2_FULL = dict(
    sampling_rate=48000,
    encoder_config=dict(dim=1536, intermediate_dim=4608, dimension=512, n_fft=1920, hop_length=960,
                        convnext_layers=24, transformer_layers=2, target_frame_rate=12.5, causal=False),
    decoder_config=dict(input_channels=1024, dim=1536, intermediate_dim=4608, convnext_layers=32, n_fft=1920,
                        hop_length=960, transformer_layers=2, target_frame_rate=12.5, causal=False),
    quantizer_config=dict(dim=512, codebook_size=1024, num_quantizers=16, decay=0.99, kmeans_init=True,
                          kmeans_iters=50, quantize_dropout=False),
    semantic_encoder_config=dict(input_channels=768, encode_channels=1536, out_channels=512,
                                 channel_ratios=[1, 1, 1], strides=[2, 1, 2]),
    semantic_decoder_config=dict(code_dim=512, output_channels=768, decode_channels=1536,
                                 channel_ratios=[1, 1, 1], strides=[2, 1, 2]),
)


def h2_small(dim=256, inter=768, enc_layers=2, dec_layers=3, tf_layers=1, sem_ch=256, nq=4, cb=256, qdim=128):
    """Reduced-width H-Codec-2.0 (same topology, n_fft/hop/strides unchanged)."""
    import copy
    c = copy.deepcopy(H2_FULL)
    c["encoder_config"].update(dim=dim, intermediate_dim=inter, dimension=qdim, convnext_layers=enc_layers,
                               transformer_layers=tf_layers)
    c["decoder_config"].update(input_channels=2 * qdim, dim=dim, intermediate_dim=inter,
                               convnext_layers=dec_layers, transformer_layers=tf_layers)
    c["quantizer_config"].update(dim=qdim, codebook_size=cb, num_quantizers=nq)
    c["semantic_encoder_config"].update(encode_channels=sem_ch, out_channels=qdim)
    c["semantic_decoder_config"].update(code_dim=qdim, decode_channels=sem_ch)
    return c


def _tf_shapes(prefix, dim, inter, layers, out):
    for i in range(layers):
        p = f"{prefix}layers.{i}."
        out[p + "self_attn.rnn.weight_ih_l0"] = ((4 * dim, dim), "lstm", dim)
        out[p + "self_attn.rnn.weight_hh_l0"] = ((4 * dim, dim), "lstm", dim)
        out[p + "self_attn.rnn.bias_ih_l0"] = ((4 * dim,), "lstm", dim)
        out[p + "self_attn.rnn.bias_hh_l0"] = ((4 * dim,), "lstm", dim)
        for n in "qkv":
            out[p + f"self_attn.{n}_proj.weight"] = ((dim, dim), "w", dim)
            out[p + f"self_attn.{n}_proj.bias"] = ((dim,), "b", dim)
        out[p + "self_attn.o_proj.weight"] = ((dim, dim), "w", dim)
        out[p + "mlp.w1.weight"] = ((inter, dim), "w", dim)
        out[p + "mlp.w2.weight"] = ((dim, inter), "w", inter)
        out[p + "mlp.w3.weight"] = ((inter, dim), "w", dim)
        out[p + "input_layernorm.weight"] = ((dim,), "norm_w", 0)
        out[p + "post_attention_layernorm.weight"] = ((dim,), "norm_w", 0)


def _convnext_shapes(prefix, dim, inter, n_layers, out):
    for i in range(n_layers):
        p = f"{prefix}{i}."
        out[p + "gamma"] = ((dim,), "gamma", n_layers)
        out[p + "dwconv.conv.weight"] = ((dim, 1, 7), "w", 7)
        out[p + "dwconv.conv.bias"] = ((dim,), "b", 7)
        out[p + "norm.weight"] = ((dim,), "norm_w", 0)
        out[p + "norm.bias"] = ((dim,), "norm_b", 0)
        out[p + "pwconv1.linear.weight"] = ((inter, dim), "w", dim)
        out[p + "pwconv1.linear.bias"] = ((inter,), "b", dim)
        out[p + "pwconv2.linear.weight"] = ((dim, inter), "w", inter)
        out[p + "pwconv2.linear.bias"] = ((dim,), "b", inter)


def h2_param_specs(cfg) -> "OrderedDict[str, tuple]":
    """name -> (shape, kind, fan) for every tensor on the encode/decode path."""
    out: OrderedDict = OrderedDict()
    e, d, s, q = cfg["encoder_config"], cfg["decoder_config"], cfg["semantic_encoder_config"], cfg["quantizer_config"]
    nf = e["n_fft"] // 2 + 1
    dim, inter = e["dim"], e["intermediate_dim"]
    out["encoder.stft.window"] = ((e["n_fft"],), "hann", 0)
    out["encoder.embed.conv.weight"] = ((dim, 2 * nf, 3), "w", 2 * nf * 3)
    out["encoder.embed.conv.bias"] = ((dim,), "b", 2 * nf * 3)
    out["encoder.norm.weight"] = ((dim,), "norm_w", 0)
    out["encoder.norm.bias"] = ((dim,), "norm_b", 0)
    _convnext_shapes("encoder.prior_net.", dim, inter, e["convnext_layers"], out)
    _tf_shapes("encoder.post_net.1.", dim, min(dim * 4, 4096), e["transformer_layers"], out)
    out["encoder.final_layer_norm.weight"] = ((dim,), "norm_w", 0)
    out["encoder.final_layer_norm.bias"] = ((dim,), "norm_b", 0)
    stride = int(50 / e["target_frame_rate"])
    out["encoder.out.conv.weight"] = ((e["dimension"], dim, 2 * stride + 1), "w", dim * (2 * stride + 1))
    out["encoder.out.conv.bias"] = ((e["dimension"],), "b", dim)

    dd, di = d["dim"], d["intermediate_dim"]
    f = int(50 / d["target_frame_rate"])
    out["decoder.embed.conv.weight"] = ((dd, d["input_channels"], f + 1), "w", d["input_channels"] * (f + 1))
    out["decoder.embed.conv.bias"] = ((dd,), "b", d["input_channels"])
    out["decoder.norm.weight"] = ((dd,), "norm_w", 0)
    out["decoder.norm.bias"] = ((dd,), "norm_b", 0)
    _convnext_shapes("decoder.post_net.", dd, di, d["convnext_layers"], out)
    out["decoder.final_layer_norm.weight"] = ((dd,), "norm_w", 0)
    out["decoder.final_layer_norm.bias"] = ((dd,), "norm_b", 0)
    for i in (0, 1, 5, 6):
        p = f"decoder.prior_net.{i}."
        for j in (1, 2):
            out[p + f"norm{j}.weight"] = ((dd,), "norm_w", 0)
            out[p + f"norm{j}.bias"] = ((dd,), "norm_b", 0)
            out[p + f"conv{j}.conv.weight"] = ((dd, dd, 3), "w", dd * 3)
            out[p + f"conv{j}.conv.bias"] = ((dd,), "b", dd * 3)
    _tf_shapes("decoder.prior_net.3.", dd, min(dd * 4, 4096), d["transformer_layers"], out)
    out["decoder.prior_net.7.weight"] = ((dd,), "norm_w", 0)
    out["decoder.prior_net.7.bias"] = ((dd,), "norm_b", 0)
    out["decoder.head.out.weight"] = ((d["n_fft"] + 2, dd), "w", dd)
    out["decoder.head.out.bias"] = ((d["n_fft"] + 2,), "b", dd)
    out["decoder.head.istft.window"] = ((d["n_fft"],), "hann", 0)

    ch = s["encode_channels"]
    out["semantic_encoder.conv.conv.weight"] = ((ch, s["input_channels"], 3), "w", s["input_channels"] * 3)
    cin = ch
    for i, st in enumerate(s["strides"]):
        cout = int(ch * s["channel_ratios"][i])
        p = f"semantic_encoder.conv_blocks.{i}."
        for u in (0, 1):
            out[p + f"res_units.{u}.conv1.conv.weight"] = ((cin, cin, 3), "w", cin * 3)
            out[p + f"res_units.{u}.conv2.weight"] = ((cin, cin, 1), "w", cin)
        k = 3 if st == 1 else 2 * st
        out[p + "conv.conv.weight"] = ((cout, cin, k), "w", cin * k)
        out[p + "conv.conv.bias"] = ((cout,), "b", cin * k)
        cin = cout
    out["semantic_encoder.conv2.conv.weight"] = ((s["out_channels"], cin, 3), "w", cin * 3)

    for name in ("quantizer", "semantic_quantizer"):
        for i in range(q["num_quantizers"]):
            p = f"{name}.layers.{i}._codebook."
            out[p + "initted"] = ((1,), "true", 0)
            out[p + "cluster_size"] = ((1, q["codebook_size"]), "ones", 0)
            out[p + "embed_avg"] = ((1, q["codebook_size"], q["dim"]), "embed_avg", 0)
            out[p + "embed"] = ((1, q["codebook_size"], q["dim"]), "embed", i)
    return out


def _gen(base_seed: int, name: str) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((base_seed * 1000003 + zlib.crc32(name.encode())) % (2**63 - 1))
    return g


def make_tensor(name, shape, kind, fan, base_seed, codebook_scale=1.0):
    g = _gen(base_seed, name)
    if kind == "hann":
        return torch.hann_window(shape[0])
    if kind == "true":
        return torch.tensor([True])
    if kind == "ones":
        return torch.ones(shape)
    if kind == "w":        # trained-like dense weight: N(0, 1/fan_in) (unit-gain)
        return torch.randn(shape, generator=g) * (1.0 / max(fan, 1)) ** 0.5
    if kind == "b":
        return torch.randn(shape, generator=g) * 0.05
    if kind == "lstm":     # torch default U(-1/sqrt(H), 1/sqrt(H))
        k = 1.0 / fan ** 0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * k
    if kind == "norm_w":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if kind == "norm_b":
        return 0.05 * torch.randn(shape, generator=g)
    if kind == "gamma":    # layer-scale: init value 1/n_layers (conv.py:195-199), perturbed
        return (1.0 / fan) * (1.0 + 0.2 * torch.randn(shape, generator=g))
    if kind == "embed":
        # residual energy shrinks layer by layer in a trained RVQ; mimic with a geometric scale
        return torch.randn(shape, generator=g) * (codebook_scale * 0.85 ** fan)
    raise ValueError(kind)


def make_h2_state_dict(cfg, seed=0, prefixes=None, codebook_scale=0.35):
    """Seeded state-dict for the H-Codec-2.0 encode/decode path.  `prefixes`
    optionally restricts generation (e.g. ("encoder.",))."""
    specs = h2_param_specs(cfg)
    sd = OrderedDict()
    for name, (shape, kind, fan) in specs.items():
        if prefixes is not None and not name.startswith(tuple(prefixes)):
            continue
        if kind == "embed_avg":   # cluster_size == 1  =>  embed_avg == embed
            en = name.replace("embed_avg", "embed")
            sd[name] = make_tensor(en, shape, "embed", specs[en][2], seed, codebook_scale)
        else:
            sd[name] = make_tensor(name, shape, kind, fan, seed, codebook_scale)
    return sd


def synth_inputs(cfg, batch, n_tokens, seed=2000):
    """Seeded synthetic (wav [B,T], feat [B,768,T50]) as SURVEY 8(d) prescribes."""
    hop_tok = int(cfg["sampling_rate"] / cfg["encoder_config"]["target_frame_rate"])
    T = n_tokens * hop_tok
    frames = T // cfg["encoder_config"]["hop_length"]
    g = torch.Generator(); g.manual_seed(seed)
    wav = 0.1 * torch.randn(batch, T, generator=g)
    g2 = torch.Generator(); g2.manual_seed(seed + 1)
    f = torch.randn(batch, cfg["semantic_encoder_config"]["input_channels"], frames, generator=g2)
    feat = torch.sign(f) * f.abs() ** 0.3   # mimics the compressed SSL features (H2/audio_tokenizer.py:58-60)
    return wav, feat
