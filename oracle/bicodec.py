"""CPU restatement of BiCodec `detokenize` - the decoder UniSE actually feeds its AR-LM tokens to.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Paths relative to /root/reference/QuarkAudio-UniSE/model/bicodec/.

    BiCodec.detokenize(semantic_tokens [B,T], global_tokens [B,1,32]) -> wav [B,1,T*320]      bicodec.py:182-199
      z_q      = FactorizedVectorQuantize.detokenize: codebook lookup (8192 x 8) + weight-normed 1x1 conv 8 -> 1024
                 (modules/vq/factorized_vector_quantize.py:154-167)
      d_vector = SpeakerEncoder.detokenize: residual FSQ (levels 4^6, one quantizer) index -> code -> Linear 6 -> 128,
                 flatten channel-major, Linear 4096 -> 1024            (modules/speaker/speaker_encoder.py:111-116,
                 modules/fsq/residual_fsq.py:112-156, modules/fsq/finite_scalar_quantization.py:139-162)
      x        = prenet(z_q, d_vector): Linear -> 2 x [SamplingBlock(ratio 1) + VocosBackbone(2 ConvNeXt)] ->
                 VocosBackbone(12 ConvNeXt, AdaLayerNorm conditioned on d_vector) -> Linear
                 (modules/encoder_decoder/feat_decoder.py:81-97, modules/blocks/vocos.py:28-335, modules/blocks/samper.py:75-100)
      wav      = WaveGenerator(x + d_vector[:, :, None]): conv k7 -> 4 x [Snake, weight-normed ConvTranspose1d (x8, x5, x4, x2),
                 3 residual units (Snake, dilated conv k7 d = 1/3/9, Snake, conv 1x1)] -> Snake -> conv k7 -> tanh
                 (modules/encoder_decoder/wave_generator.py:32-91, modules/blocks/layers.py:24-67)

The reference ships the code but neither `config.yaml` nor weights (they come from Spark-TTS-0.5B, U/README.md:57-74);
BICODEC_FULL restates that published configuration.  Pinning: oracle/make_golden_bicodec.py builds the reference's own
module classes with these hyper-parameters and the seeded weights below and compares (5e-6 relative on CPU fp32); the
residual-FSQ class needs the absent `einx` package, for which the pin script supplies the one gather it uses.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from .weights import _gen

BICODEC_FULL = dict(
    sample_rate=16000, hop=320,
    quantizer=dict(input_dim=1024, codebook_size=8192, codebook_dim=8),
    speaker=dict(out_dim=1024, latent_dim=128, token_num=32, fsq_levels=[4, 4, 4, 4, 4, 4], fsq_num_quantizers=1),
    prenet=dict(input_channels=1024, vocos_dim=384, vocos_intermediate_dim=2048, vocos_num_layers=12, out_channels=1024,
                condition_dim=1024, sample_ratios=[1, 1], use_tanh_at_final=False),
    decoder=dict(input_channel=1024, channels=1536, rates=[8, 5, 4, 2], kernel_sizes=[16, 11, 8, 4]),
)


def bicodec_small():
    """reduced-width variant, same topology (odd-k transposed conv kept: k=11, s=5)"""
    return dict(
        sample_rate=16000, hop=320,
        quantizer=dict(input_dim=128, codebook_size=256, codebook_dim=8),
        speaker=dict(out_dim=128, latent_dim=16, token_num=8, fsq_levels=[4, 4, 4, 4, 4, 4], fsq_num_quantizers=1),
        prenet=dict(input_channels=128, vocos_dim=64, vocos_intermediate_dim=192, vocos_num_layers=3, out_channels=128,
                    condition_dim=128, sample_ratios=[1, 1], use_tanh_at_final=False),
        decoder=dict(input_channel=128, channels=256, rates=[8, 5, 4, 2], kernel_sizes=[16, 11, 8, 4]),
    )


# --------------------------------------------------------------------------- parameter layout (reference state-dict keys)
def param_specs(c):
    """name -> (shape, kind); kinds: w (dense, fan-in scaled), b, g (weight-norm gain), nw / nb (norm affine),
    gamma (layer scale, value = fan), alpha (Snake), cb (codebook), ada_s / ada_b (AdaLayerNorm scale / shift linears)."""
    out = OrderedDict()
    q, s, p, d = c["quantizer"], c["speaker"], c["prenet"], c["decoder"]

    def wn(prefix, shape, n_out=None):          # torch.nn.utils.weight_norm, dim=0
        out[prefix + "bias"] = ((shape[0] if n_out is None else n_out,), "b")
        out[prefix + "weight_g"] = ((shape[0],) + (1,) * (len(shape) - 1), "g")
        out[prefix + "weight_v"] = (tuple(shape), "w")

    out["quantizer.codebook.weight"] = ((q["codebook_size"], q["codebook_dim"]), "cb")
    wn("quantizer.out_project.", (q["input_dim"], q["codebook_dim"], 1))
    L = len(s["fsq_levels"])
    out["speaker_encoder.quantizer.project_out.weight"] = ((s["latent_dim"], L), "w")
    out["speaker_encoder.quantizer.project_out.bias"] = ((s["latent_dim"],), "b")
    out["speaker_encoder.project.weight"] = ((s["out_dim"], s["latent_dim"] * s["token_num"]), "w")
    out["speaker_encoder.project.bias"] = ((s["out_dim"],), "b")

    dim, inter = p["vocos_dim"], p["vocos_intermediate_dim"]

    def backbone(prefix, layers, cond):
        out[prefix + "embed.weight"] = ((dim, dim, 7), "w")
        out[prefix + "embed.bias"] = ((dim,), "b")

        def norm(pp):
            if cond:
                out[pp + "scale.weight"] = ((dim, cond), "ada_s"); out[pp + "scale.bias"] = ((dim,), "nw")
                out[pp + "shift.weight"] = ((dim, cond), "ada_b"); out[pp + "shift.bias"] = ((dim,), "nb")
            else:
                out[pp + "weight"] = ((dim,), "nw"); out[pp + "bias"] = ((dim,), "nb")
        norm(prefix + "norm.")
        for i in range(layers):
            b = f"{prefix}convnext.{i}."
            out[b + "gamma"] = ((dim,), ("gamma", layers))
            out[b + "dwconv.weight"] = ((dim, 1, 7), "w"); out[b + "dwconv.bias"] = ((dim,), "b")
            norm(b + "norm.")
            out[b + "pwconv1.weight"] = ((inter, dim), "w"); out[b + "pwconv1.bias"] = ((inter,), "b")
            out[b + "pwconv2.weight"] = ((dim, inter), "w"); out[b + "pwconv2.bias"] = ((dim,), "b")
        out[prefix + "final_layer_norm.weight"] = ((dim,), "nw"); out[prefix + "final_layer_norm.bias"] = ((dim,), "nb")

    out["prenet.linear_pre.weight"] = ((dim, p["input_channels"]), "w")
    out["prenet.linear_pre.bias"] = ((dim,), "b")
    for i, r in enumerate(p["sample_ratios"]):
        if r != 1:
            raise NotImplementedError("SamplingBlock ratios other than 1 (the shipped prenet uses [1, 1])")
        backbone(f"prenet.downsample.{i}.1.", 2, None)
    backbone("prenet.vocos_backbone.", p["vocos_num_layers"], p["condition_dim"])
    out["prenet.linear.weight"] = ((p["out_channels"], dim), "w")
    out["prenet.linear.bias"] = ((p["out_channels"],), "b")

    ch = d["channels"]
    wn("decoder.model.0.", (ch, d["input_channel"], 7))
    for i, (k, r) in enumerate(zip(d["kernel_sizes"], d["rates"])):
        cin, cout = ch // 2 ** i, ch // 2 ** (i + 1)
        b = f"decoder.model.{i + 1}.block."
        out[b + "0.alpha"] = ((1, cin, 1), "alpha")
        wn(b + "1.", (cin, cout, k), cout)                   # ConvTranspose1d weight [Cin, Cout, k], gain per Cin slice
        for j in range(3):
            u = f"{b}{j + 2}.block."
            out[u + "0.alpha"] = ((1, cout, 1), "alpha")
            wn(u + "1.", (cout, cout, 7))
            out[u + "2.alpha"] = ((1, cout, 1), "alpha")
            wn(u + "3.", (cout, cout, 1))
    n = len(d["rates"])
    out[f"decoder.model.{n + 1}.alpha"] = ((1, ch // 2 ** n, 1), "alpha")
    wn(f"decoder.model.{n + 2}.", (1, ch // 2 ** n, 7))
    return out


def make_state_dict(c, seed=0):
    sd = OrderedDict()
    for name, (shape, kind) in param_specs(c).items():
        g = _gen(seed, name)
        fan = None
        if isinstance(kind, tuple):
            kind, fan = kind
        if kind == "w":
            fan_in = 1
            for v in shape[1:]:
                fan_in *= v
            if name.endswith("block.1.weight_v") and len(shape) == 3 and "decoder.model" in name and shape[0] != shape[1]:
                # transposed conv [Cin, Cout, k]: every output sample sees k/stride taps of Cin channels
                fan_in = shape[0] * 2
            t = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif kind == "b":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "g":            # filled below: gain ~ ||v|| so that the effective weight keeps its unit-gain scale
            t = torch.zeros(shape)
        elif kind == "nw":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "nb":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "gamma":
            t = (1.0 / fan) * (1.0 + 0.2 * torch.randn(shape, generator=g))
        elif kind == "alpha":
            t = 1.0 + 0.3 * torch.rand(shape, generator=g)
        elif kind == "cb":
            t = torch.randn(shape, generator=g)
        elif kind == "ada_s":        # conditioning perturbs scale ~ 1 (+ bias "nw") and shift ~ 0 by a few percent
            t = 0.1 * torch.randn(shape, generator=g) / shape[1] ** 0.5
        elif kind == "ada_b":
            t = 0.1 * torch.randn(shape, generator=g) / shape[1] ** 0.5
        else:
            raise ValueError(kind)
        sd[name] = t
    for name in list(sd):            # weight-norm gains: ||v|| per dim-0 slice, perturbed
        if name.endswith("weight_g"):
            v = sd[name[:-1] + "v"]
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(sd[name].shape)
            gain = 1.0
            if ".block.3." in name:          # residual-unit output conv: a trained decoder keeps the branch well below the trunk
                gain = 0.3
            elif name == f"decoder.model.{len(c['decoder']['rates']) + 2}.weight_g":
                gain = 0.1                   # pre-tanh waveform of O(0.2), not saturated
            sd[name] = gain * nrm * (1.0 + 0.1 * torch.randn(nrm.shape, generator=_gen(seed, name)))
    return sd


def synth_tokens(c, batch, frames, seed=0):
    g = torch.Generator(); g.manual_seed(seed)
    sem = torch.randint(0, c["quantizer"]["codebook_size"], (batch, frames), generator=g)
    n_codes = 1
    for v in c["speaker"]["fsq_levels"]:
        n_codes *= v
    glob = torch.randint(0, n_codes, (batch, c["speaker"]["fsq_num_quantizers"], c["speaker"]["token_num"]), generator=g)
    return sem, glob


# --------------------------------------------------------------------------- forward
def wn_weight(sd, p):
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, norm over all dims but the first (layers.py:24-29)."""
    v, g = sd[p + "weight_v"], sd[p + "weight_g"]
    return v * (g / v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape))


def snake(x, alpha):
    """layers.py:33-38"""
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def fvq_detokenize(sd, sem):
    """factorized_vector_quantize.py:154-167: [B,T] -> [B, input_dim, T]"""
    z = F.embedding(sem, sd["quantizer.codebook.weight"]).transpose(1, 2)
    return F.conv1d(z, wn_weight(sd, "quantizer.out_project."), sd["quantizer.out_project.bias"])


def fsq_codes(levels, idx):
    """finite_scalar_quantization.py:139-162: index -> per-dimension level -> (level - L//2) / (L//2)"""
    lv = torch.tensor(levels, dtype=torch.int64)
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1]), dtype=torch.int64), 0)
    half = lv // 2
    return ((idx[..., None] // basis) % lv - half).float() / half.float()


def speaker_detokenize(sd, c, glob):
    """speaker_encoder.py:111-116 with residual_fsq.py:112-156: [B, nq, N] -> d_vector [B, out_dim]"""
    s = c["speaker"]
    lv = torch.tensor(s["fsq_levels"], dtype=torch.float32)
    idx = glob.transpose(1, 2)                                    # b n q
    codes = 0
    for qi in range(idx.shape[-1]):
        codes = codes + fsq_codes(s["fsq_levels"], idx[..., qi]) * (lv - 1) ** (-qi)
    zq = F.linear(codes, sd["speaker_encoder.quantizer.project_out.weight"], sd["speaker_encoder.quantizer.project_out.bias"])
    x = zq.transpose(1, 2).reshape(zq.shape[0], -1)               # [B, latent_dim * token_num], channel-major
    return F.linear(x, sd["speaker_encoder.project.weight"], sd["speaker_encoder.project.bias"])


def _norm(sd, p, x, cond, dim):
    """nn.LayerNorm(1e-6) or AdaLayerNorm (vocos.py:88-111); x [B,T,C]"""
    if cond is None:
        return F.layer_norm(x, (dim,), sd[p + "weight"], sd[p + "bias"], 1e-6)
    scale = F.linear(cond, sd[p + "scale.weight"], sd[p + "scale.bias"])
    shift = F.linear(cond, sd[p + "shift.weight"], sd[p + "shift.bias"])
    return F.layer_norm(x, (dim,), eps=1e-6) * scale[:, None] + shift[:, None]


def vocos_backbone(sd, p, x, layers, dim, cond=None, taps=None):
    """vocos.py:273-335: x [B,C,T] -> [B,T,dim]"""
    x = F.conv1d(x, sd[p + "embed.weight"], sd[p + "embed.bias"], padding=3)
    x = _norm(sd, p + "norm.", x.transpose(1, 2), cond, dim).transpose(1, 2)
    for i in range(layers):
        b = f"{p}convnext.{i}."
        y = F.conv1d(x, sd[b + "dwconv.weight"], sd[b + "dwconv.bias"], padding=3, groups=dim)
        y = _norm(sd, b + "norm.", y.transpose(1, 2), cond, dim)
        y = F.linear(F.gelu(F.linear(y, sd[b + "pwconv1.weight"], sd[b + "pwconv1.bias"])), sd[b + "pwconv2.weight"],
                     sd[b + "pwconv2.bias"])
        x = x + (sd[b + "gamma"] * y).transpose(1, 2)
    return F.layer_norm(x.transpose(1, 2), (dim,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-6)


def prenet_forward(sd, c, z_q, d_vector, taps=None):
    """feat_decoder.py:81-97: z_q [B,in,T], d_vector [B,cond] -> [B,out,T]"""
    p = c["prenet"]
    dim = p["vocos_dim"]
    x = F.linear(z_q.transpose(1, 2), sd["prenet.linear_pre.weight"], sd["prenet.linear_pre.bias"])      # [B,T,dim]
    for i, r in enumerate(p["sample_ratios"]):
        # SamplingBlock with up = down = 1: conv_res + skip1_res + skip2_res = 3 x   (samper.py:75-100)
        x = vocos_backbone(sd, f"prenet.downsample.{i}.1.", (x + x + x).transpose(1, 2), 2, dim)
    if taps is not None:
        taps["prenet.pre"] = x
    x = vocos_backbone(sd, "prenet.vocos_backbone.", x.transpose(1, 2), p["vocos_num_layers"], dim, d_vector)
    x = F.linear(x, sd["prenet.linear.weight"], sd["prenet.linear.bias"]).transpose(1, 2)
    return torch.tanh(x) if p["use_tanh_at_final"] else x


def wave_generator(sd, c, x, taps=None):
    """wave_generator.py:59-91: x [B, input_channel, T] -> [B, 1, T * prod(rates)]"""
    d = c["decoder"]
    x = F.conv1d(x, wn_weight(sd, "decoder.model.0."), sd["decoder.model.0.bias"], padding=3)
    for i, (k, r) in enumerate(zip(d["kernel_sizes"], d["rates"])):
        b = f"decoder.model.{i + 1}.block."
        x = snake(x, sd[b + "0.alpha"])
        x = F.conv_transpose1d(x, wn_weight(sd, b + "1."), sd[b + "1.bias"], stride=r, padding=(k - r) // 2)
        for j, dil in enumerate((1, 3, 9)):
            u = f"{b}{j + 2}.block."
            y = F.conv1d(snake(x, sd[u + "0.alpha"]), wn_weight(sd, u + "1."), sd[u + "1.bias"], dilation=dil, padding=3 * dil)
            y = F.conv1d(snake(y, sd[u + "2.alpha"]), wn_weight(sd, u + "3."), sd[u + "3.bias"])
            x = x + y
        if taps is not None:
            taps[f"dec.stage{i}"] = x
    n = len(d["rates"])
    x = snake(x, sd[f"decoder.model.{n + 1}.alpha"])
    x = F.conv1d(x, wn_weight(sd, f"decoder.model.{n + 2}."), sd[f"decoder.model.{n + 2}.bias"], padding=3)
    return torch.tanh(x)


@torch.no_grad()
def detokenize(sd, c, semantic_tokens, global_tokens, taps=None):
    """bicodec.py:182-199"""
    z_q = fvq_detokenize(sd, semantic_tokens)
    d_vector = speaker_detokenize(sd, c, global_tokens)
    x = prenet_forward(sd, c, z_q, d_vector, taps)
    x = x + d_vector.unsqueeze(-1)
    if taps is not None:
        taps["z_q"], taps["d_vector"], taps["prenet.out"] = z_q, d_vector, x
    return wave_generator(sd, c, x, taps)
