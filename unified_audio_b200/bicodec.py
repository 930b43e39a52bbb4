"""BiCodec `detokenize` (the decoder UniSE feeds its AR-LM tokens to) on libquark_b200 - SURVEY.md 8(f).1.

Mirrors QuarkAudio-UniSE/model/bicodec/bicodec.py:182-199:
    BiCodec(config).detokenize(semantic_tokens [B,T] int64, global_tokens [B,1,32] int64) -> wav [B,1,T*320] fp32
state_dict keys are the reference's (`quantizer.*`, `speaker_encoder.*`, `prenet.*`, `decoder.*`, old-style weight-norm
`weight_g / weight_v` pairs); keys of the tokenize side (`encoder.*`, `postnet.*`, ECAPA / perceiver, `mel_transformer.*`,
`quantizer.in_project.*`) are accepted at load and ignored.  The reference ships no `config.yaml` (it comes with the
Spark-TTS-0.5B checkpoint, U/README.md:57-74); BICODEC_CONFIG restates that published configuration.

How the path maps onto the library (every arithmetic op is a libquark_b200 kernel; channel-last activations):
  * FactorizedVectorQuantize.detokenize (modules/vq/factorized_vector_quantize.py:154-167) and the residual-FSQ de-quantiser
    (modules/fsq/residual_fsq.py:112-156) are index -> row gathers from tables prepared at load
    (codebook @ out_project, implicit FSQ codebook @ project_out);
  * prenet (modules/encoder_decoder/feat_decoder.py:81-97) = Vocos ConvNeXt stacks: the H-Codec ConvNeXt kernels, with
    AdaLayerNorm (modules/blocks/vocos.py:88-111) as a per-clip scale / shift row produced by ONE GEMM for all 13 norms;
  * WaveGenerator (modules/encoder_decoder/wave_generator.py:32-91): Snake -> fp16 planes (`qb_snake_planes`), dilated k=7
    convs as TMA-im2col GEMMs with a tap spacing (`qb_gemm_desc.dilation`), and every weight-normed ConvTranspose1d
    (k, stride s) as a ceil(k/s)-tap GEMM that produces all s output phases as s*Cout columns - its row-major output IS the
    up-sampled channel-last signal, shifted by the transposed conv's padding (no col2im, no zero insertion).
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn

from . import ops
from .codec import _Tree, _pad_to
from .ops import ACT_GELU, ACT_SNAKE, ACT_TANH, Planes, rowmap

BICODEC_CONFIG = dict(
    sample_rate=16000, hop=320,
    quantizer=dict(input_dim=1024, codebook_size=8192, codebook_dim=8),
    speaker=dict(out_dim=1024, latent_dim=128, token_num=32, fsq_levels=[4, 4, 4, 4, 4, 4], fsq_num_quantizers=1),
    prenet=dict(input_channels=1024, vocos_dim=384, vocos_intermediate_dim=2048, vocos_num_layers=12, out_channels=1024,
                condition_dim=1024, sample_ratios=[1, 1], use_tanh_at_final=False),
    decoder=dict(input_channel=1024, channels=1536, rates=[8, 5, 4, 2], kernel_sizes=[16, 11, 8, 4]),
)

# 3-term split (True) or single-pass fp16 (False) per GEMM group; "accurate" is the default until the error budget of the
# 26-conv generator is mapped
PRECISION = {
    "accurate": dict(prenet=True, gen=True),
    "mixed": dict(prenet=False, gen=True),
    "fast": dict(prenet=False, gen=False),
}


def bicodec_spec(c) -> Dict[str, tuple]:
    """Reference state-dict keys -> shapes of the detokenize path."""
    out: Dict[str, tuple] = {}
    q, s, p, d = c["quantizer"], c["speaker"], c["prenet"], c["decoder"]

    def wn(prefix, shape, n_out=None):
        out[prefix + "bias"] = (shape[0] if n_out is None else n_out,)
        out[prefix + "weight_g"] = (shape[0],) + (1,) * (len(shape) - 1)
        out[prefix + "weight_v"] = tuple(shape)

    out["quantizer.codebook.weight"] = (q["codebook_size"], q["codebook_dim"])
    wn("quantizer.out_project.", (q["input_dim"], q["codebook_dim"], 1))
    out["speaker_encoder.quantizer.project_out.weight"] = (s["latent_dim"], len(s["fsq_levels"]))
    out["speaker_encoder.quantizer.project_out.bias"] = (s["latent_dim"],)
    out["speaker_encoder.project.weight"] = (s["out_dim"], s["latent_dim"] * s["token_num"])
    out["speaker_encoder.project.bias"] = (s["out_dim"],)
    dim, inter = p["vocos_dim"], p["vocos_intermediate_dim"]

    def norm(pp, cond):
        if cond:
            out[pp + "scale.weight"] = (dim, cond); out[pp + "scale.bias"] = (dim,)
            out[pp + "shift.weight"] = (dim, cond); out[pp + "shift.bias"] = (dim,)
        else:
            out[pp + "weight"] = (dim,); out[pp + "bias"] = (dim,)

    def backbone(prefix, layers, cond):
        out[prefix + "embed.weight"] = (dim, dim, 7)
        out[prefix + "embed.bias"] = (dim,)
        norm(prefix + "norm.", cond)
        for i in range(layers):
            b = f"{prefix}convnext.{i}."
            out[b + "gamma"] = (dim,)
            out[b + "dwconv.weight"] = (dim, 1, 7); out[b + "dwconv.bias"] = (dim,)
            norm(b + "norm.", cond)
            out[b + "pwconv1.weight"] = (inter, dim); out[b + "pwconv1.bias"] = (inter,)
            out[b + "pwconv2.weight"] = (dim, inter); out[b + "pwconv2.bias"] = (dim,)
        out[prefix + "final_layer_norm.weight"] = (dim,); out[prefix + "final_layer_norm.bias"] = (dim,)

    out["prenet.linear_pre.weight"] = (dim, p["input_channels"]); out["prenet.linear_pre.bias"] = (dim,)
    for i, r in enumerate(p["sample_ratios"]):
        if r != 1:
            raise NotImplementedError("prenet SamplingBlock ratios other than 1 (the shipped configuration uses [1, 1])")
        backbone(f"prenet.downsample.{i}.1.", 2, None)
    backbone("prenet.vocos_backbone.", p["vocos_num_layers"], p["condition_dim"])
    out["prenet.linear.weight"] = (p["out_channels"], dim); out["prenet.linear.bias"] = (p["out_channels"],)
    ch = d["channels"]
    wn("decoder.model.0.", (ch, d["input_channel"], 7))
    for i, (k, r) in enumerate(zip(d["kernel_sizes"], d["rates"])):
        cin, cout = ch // 2 ** i, ch // 2 ** (i + 1)
        b = f"decoder.model.{i + 1}.block."
        out[b + "0.alpha"] = (1, cin, 1)
        wn(b + "1.", (cin, cout, k), cout)
        for j in range(3):
            u = f"{b}{j + 2}.block."
            out[u + "0.alpha"] = (1, cout, 1)
            wn(u + "1.", (cout, cout, 7))
            out[u + "2.alpha"] = (1, cout, 1)
            wn(u + "3.", (cout, cout, 1))
    n = len(d["rates"])
    out[f"decoder.model.{n + 1}.alpha"] = (1, ch // 2 ** n, 1)
    wn(f"decoder.model.{n + 2}.", (1, ch // 2 ** n, 7))
    return out


_IGNORED = ("encoder.", "postnet.", "mel_transformer.", "speaker_encoder.speaker_encoder.", "speaker_encoder.perceiver_sampler.",
            "speaker_encoder.quantizer.project_in.", "quantizer.in_project.", "quantizer.cluster_size")


class BiCodec(nn.Module):
    def __init__(self, config: dict = None, precision: str = "accurate"):
        super().__init__()
        self.cfg = dict(config or BICODEC_CONFIG)
        self.policy = PRECISION[precision]
        tree = _Tree.build(bicodec_spec(self.cfg))
        for name, child in tree.named_children():
            self.add_module(name, child)
        self._w, self._ws = None, {}
        self.eval()

    # ------------------------------------------------------------------ state
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = {k: v for k, v in state_dict.items() if not k.startswith(_IGNORED)}
        r = super().load_state_dict(sd, strict=strict, assign=assign)
        self._w = None
        return r

    def _apply(self, fn, *a, **k):
        self._w, self._ws = None, {}
        return super()._apply(fn, *a, **k)

    def _dev(self):
        return self.quantizer.codebook.weight.device

    def _buf(self, name, shape, dtype=torch.float32):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self._dev())
            self._ws[key] = t
        return t

    def _planes(self, name, shape, split):
        key = ("P", name, tuple(shape), bool(split))
        p = self._ws.get(key)
        if p is None:
            p = Planes.zeros(shape, split, self._dev())
            self._ws[key] = p
        return p

    # ------------------------------------------------------------------ load-time weight preparation
    def _prepare(self):
        if self._w is not None:
            return self._w
        dev = self._dev()
        if dev.type != "cuda":
            raise RuntimeError("unified_audio_b200.BiCodec runs on CUDA only (no CPU fallback): call .cuda() first")
        sd = {k: v.detach().float() for k, v in self.state_dict().items()}
        c = self.cfg
        q, s, p, d = c["quantizer"], c["speaker"], c["prenet"], c["decoder"]
        sp_pre, sp_gen = self.policy["prenet"], self.policy["gen"]

        def wnw(prefix):      # fold torch.nn.utils.weight_norm (dim 0): w = g * v / ||v||   (layers.py:24-29)
            v, g = sd[prefix + "weight_v"], sd[prefix + "weight_g"]
            return v * (g / v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape))

        def conv_pack(w, split):   # [Cout, Cin, k] -> planes [Cout, k * Cin_pad] (tap-major K, zero channel padding)
            co, ci, k = w.shape
            cp = _pad_to(ci, 64)
            out = torch.zeros(co, k, cp, device=dev)
            out[:, :, :ci] = w.permute(0, 2, 1)
            return Planes.from_f32(out.reshape(co, k * cp), split)

        def convt_pack(w, s_, split):
            """ConvTranspose1d weight [Cin, Cout, k], stride s -> [s * Cout, J * Cin_pad], J = ceil(k / s):
            row (r, co), GEMM tap t multiplies x[q + t - (J-1)] with w[:, co, r + (J-1-t) * s]."""
            ci, co, k = w.shape
            J = -(-k // s_)
            cp = _pad_to(ci, 64)
            wp = torch.zeros(ci, co, J * s_, device=dev)
            wp[:, :, :k] = w
            wp = wp.reshape(ci, co, J, s_)                          # [ci, co, j, r]
            out = torch.zeros(s_, co, J, cp, device=dev)
            out[:, :, :, :ci] = wp.permute(3, 1, 2, 0).flip(2)       # tap t <-> j = J-1-t
            return Planes.from_f32(out.reshape(s_ * co, J * cp), split), J

        W = {}
        # ---- gathers
        W["zq_table"] = (sd["quantizer.codebook.weight"] @ wnw("quantizer.out_project.")[:, :, 0].t()
                         + sd["quantizer.out_project.bias"])[None].contiguous()           # [1, K, input_dim]
        levels = torch.tensor(s["fsq_levels"], dtype=torch.int64, device=dev)
        basis = torch.cumprod(torch.tensor([1] + list(s["fsq_levels"][:-1]), dtype=torch.int64, device=dev), 0)
        half = (levels // 2).float()
        ids = torch.arange(int(torch.prod(levels)), device=dev)
        implicit = (((ids[:, None] // basis) % levels).float() - half) / half             # finite_scalar_quantization.py:139-162
        if s["fsq_num_quantizers"] != 1:
            raise NotImplementedError("residual FSQ with more than one quantizer (the shipped speaker encoder uses one)")
        W["fsq_table"] = (implicit @ sd["speaker_encoder.quantizer.project_out.weight"].t()
                          + sd["speaker_encoder.quantizer.project_out.bias"])[None].contiguous()   # [1, 4096, latent]
        L, N = s["latent_dim"], s["token_num"]
        pw = sd["speaker_encoder.project.weight"].reshape(s["out_dim"], L, N).permute(0, 2, 1).reshape(s["out_dim"], N * L)
        W["spk_w"] = Planes.from_f32(pw.contiguous(), True)       # flatten order (d, n) -> gather order (n, d)
        W["spk_b"] = sd["speaker_encoder.project.bias"].contiguous()

        # ---- prenet
        dim = p["vocos_dim"]

        def backbone(prefix, layers, cond, in_scale):
            blocks = []
            for i in range(layers):
                b = f"{prefix}convnext.{i}."
                blk = dict(dw_w=sd[b + "dwconv.weight"].reshape(dim, 7).contiguous(), dw_b=sd[b + "dwconv.bias"].contiguous(),
                           w1=Planes.from_f32(sd[b + "pwconv1.weight"], sp_pre), b1=sd[b + "pwconv1.bias"].contiguous(),
                           w2=Planes.from_f32(sd[b + "pwconv2.weight"], sp_pre), b2=sd[b + "pwconv2.bias"].contiguous(),
                           gamma=sd[b + "gamma"].contiguous())
                if not cond:
                    blk.update(ln_w=sd[b + "norm.weight"].contiguous(), ln_b=sd[b + "norm.bias"].contiguous())
                blocks.append(blk)
            out = dict(embed=conv_pack(sd[prefix + "embed.weight"] * in_scale, sp_pre), embed_b=sd[prefix + "embed.bias"].contiguous(),
                       blocks=blocks, fn_w=sd[prefix + "final_layer_norm.weight"].contiguous(),
                       fn_b=sd[prefix + "final_layer_norm.bias"].contiguous(), layers=layers)
            if not cond:
                out.update(n_w=sd[prefix + "norm.weight"].contiguous(), n_b=sd[prefix + "norm.bias"].contiguous())
            return out

        W["lin_pre"] = Planes.from_f32(sd["prenet.linear_pre.weight"], sp_pre)
        W["lin_pre_b"] = sd["prenet.linear_pre.bias"].contiguous()
        # SamplingBlock(up = down = 1) returns conv_res + skip1 + skip2 = 3 x (samper.py:75-100): folded into the embed conv
        W["down"] = [backbone(f"prenet.downsample.{i}.1.", 2, None, 3.0) for i in range(len(p["sample_ratios"]))]
        W["bb"] = backbone("prenet.vocos_backbone.", p["vocos_num_layers"], p["condition_dim"], 1.0)
        # all AdaLayerNorm scale / shift projections of the conditioned backbone as one [2 (L+1) dim, cond] matrix
        names = ["prenet.vocos_backbone.norm."] + [f"prenet.vocos_backbone.convnext.{i}.norm." for i in range(p["vocos_num_layers"])]
        W["cond_w"] = Planes.from_f32(torch.cat([torch.cat([sd[n + "scale.weight"], sd[n + "shift.weight"]], 0) for n in names], 0), True)
        W["cond_b"] = torch.cat([torch.cat([sd[n + "scale.bias"], sd[n + "shift.bias"]], 0) for n in names], 0).contiguous()
        W["lin"] = Planes.from_f32(sd["prenet.linear.weight"], sp_pre)
        W["lin_b"] = sd["prenet.linear.bias"].contiguous()

        # ---- WaveGenerator
        G = dict(conv0=conv_pack(wnw("decoder.model.0."), sp_gen), conv0_b=sd["decoder.model.0.bias"].contiguous(), stages=[])
        ch = d["channels"]
        for i, (k, r) in enumerate(zip(d["kernel_sizes"], d["rates"])):
            cin, cout = ch // 2 ** i, ch // 2 ** (i + 1)
            b = f"decoder.model.{i + 1}.block."
            wt, J = convt_pack(wnw(b + "1."), r, sp_gen)
            st = dict(alpha=sd[b + "0.alpha"].reshape(-1).contiguous(), wt=wt, J=J, k=k, s=r, cin=cin, cout=cout,
                      bt=sd[b + "1.bias"].repeat(r).contiguous(), units=[])
            for j, dil in enumerate((1, 3, 9)):
                u = f"{b}{j + 2}.block."
                st["units"].append(dict(dil=dil, a1=sd[u + "0.alpha"].reshape(-1).contiguous(), w1=conv_pack(wnw(u + "1."), sp_gen),
                                        b1=sd[u + "1.bias"].contiguous(), a2=sd[u + "2.alpha"].reshape(-1).contiguous(),
                                        w2=conv_pack(wnw(u + "3."), sp_gen), b2=sd[u + "3.bias"].contiguous()))
            G["stages"].append(st)
        n = len(d["rates"])
        G["alpha_f"] = sd[f"decoder.model.{n + 1}.alpha"].reshape(-1).contiguous()
        G["conv_f"] = conv_pack(wnw(f"decoder.model.{n + 2}."), sp_gen)
        G["conv_f_b"] = sd[f"decoder.model.{n + 2}.bias"].contiguous()
        W["gen"] = G
        self._w = W
        return W

    # ------------------------------------------------------------------ blocks
    def _conv(self, a: Planes, w: Planes, n, B, rows_in, ld, m, taps, **kw):
        ops.gemm(a, w, n, a_batch=B, a_rows_per_batch=rows_in, a_ld=ld, m_per_batch=m, taps=taps, **kw)

    def _backbone(self, bw, x, B, T, dim, inter, cond=None, cond_stride=0, tag=""):
        """VocosBackbone (vocos.py:273-335) on the fp32 trunk x [B*T, dim]; returns a new fp32 [B*T, dim]."""
        M, sp = B * T, self.policy["prenet"]
        cp = _pad_to(dim, 64)
        pad = self._planes("bb_pad", (B, T + 6, cp), sp)
        ops.rows_to_planes(x, B, T, dim, pad, cp, T + 6, 3)
        y = self._buf("bb_y", (M, dim))
        self._conv(pad, bw["embed"], dim, B, T + 6, cp, T, 7, bias=bw["embed_b"], out_f32=rowmap(y, dim, T, 0))
        h = self._buf("bb_h" + tag, (M, dim))
        if cond is None:
            ops.layernorm(y, bw["n_w"], bw["n_b"], B, T, dim, out_f32=h)
        else:
            ops.adalayernorm(y, cond[0], cond[0][dim:], cond_stride, B, T, dim, out_f32=h)
        t1 = self._planes("bb_t1", (M, dim), sp)
        hid = self._planes("bb_hid", (M, inter), sp)
        hm = rowmap(h, dim, M, 0)
        for i, blk in enumerate(bw["blocks"]):
            if cond is None:
                ops.dwconv7_ln(h, blk["dw_w"], blk["dw_b"], blk["ln_w"], blk["ln_b"], B, T, dim, t1)
            else:
                cs = cond[i + 1]
                ops.dwconv7_adaln(h, blk["dw_w"], blk["dw_b"], cs, cs[dim:], cond_stride, B, T, dim, t1)
            ops.gemm(t1, blk["w1"], inter, a_batch=1, a_rows_per_batch=M, a_ld=dim, m_per_batch=M, bias=blk["b1"], act=ACT_GELU,
                     out_planes=hid, out_planes_map=(inter, M, 0))
            ops.gemm(hid, blk["w2"], dim, a_batch=1, a_rows_per_batch=M, a_ld=inter, m_per_batch=M, bias=blk["b2"],
                     gamma=blk["gamma"], residual=hm, out_f32=hm)
        out = self._buf("bb_out" + tag, (M, dim))
        ops.layernorm(h, bw["fn_w"], bw["fn_b"], B, T, dim, out_f32=out)
        return out

    # ------------------------------------------------------------------ public surface
    @torch.no_grad()
    def detokenize(self, semantic_tokens: torch.Tensor, global_tokens: torch.Tensor, taps=None) -> torch.Tensor:
        """bicodec.py:182-199"""
        W = self._prepare()
        c = self.cfg
        q, s, p, d = c["quantizer"], c["speaker"], c["prenet"], c["decoder"]
        dev = self._dev()
        B, T = semantic_tokens.shape
        M = B * T
        D_in, dim, inter = q["input_dim"], p["vocos_dim"], p["vocos_intermediate_dim"]
        if dim % 64 or inter % 64 or D_in % 64 or p["out_channels"] % 64 or p["condition_dim"] % 64:
            raise ValueError("BiCodec widths must be multiples of 64")
        sp_pre, sp_gen = self.policy["prenet"], self.policy["gen"]
        # ---- z_q: codebook row @ out_project, gathered  (factorized_vector_quantize.py:154-167)
        zq = self._buf("zq", (M, D_in))
        ops.rvq_decode(semantic_tokens.reshape(M, 1).long().contiguous(), W["zq_table"], M, D_in, q["codebook_size"], 1, zq, D_in, 0)
        # ---- d_vector (speaker_encoder.py:111-116)
        N, L = s["token_num"], s["latent_dim"]
        if tuple(global_tokens.shape) != (B, s["fsq_num_quantizers"], N):
            raise ValueError(f"global_tokens must be [B, {s['fsq_num_quantizers']}, {N}]")
        codes = self._buf("spk_codes", (B * N, L))
        ops.rvq_decode(global_tokens.reshape(B * N, 1).long().contiguous(), W["fsq_table"], B * N, L, W["fsq_table"].shape[1], 1,
                       codes, L, 0)
        cpl = self._planes("spk_codes_p", (B, N * L), True)
        ops.split_f16(codes, cpl)
        dvec = self._buf("dvec", (B, s["out_dim"]))
        ops.gemm(cpl, W["spk_w"], s["out_dim"], a_batch=1, a_rows_per_batch=B, a_ld=N * L, m_per_batch=B, bias=W["spk_b"],
                 out_f32=rowmap(dvec, s["out_dim"], B, 0))
        # ---- AdaLayerNorm rows for all 13 norms: [B, (layers + 1) * 2 * dim]
        n_norm = p["vocos_num_layers"] + 1
        dpl = self._planes("dvec_p", (B, s["out_dim"]), True)
        ops.split_f16(dvec, dpl)
        cond = self._buf("cond", (B, n_norm * 2 * dim))
        ops.gemm(dpl, W["cond_w"], n_norm * 2 * dim, a_batch=1, a_rows_per_batch=B, a_ld=p["condition_dim"], m_per_batch=B,
                 bias=W["cond_b"], out_f32=rowmap(cond, n_norm * 2 * dim, B, 0))
        cond_rows = [cond.view(-1)[j * 2 * dim:] for j in range(n_norm)]        # scale at +0, shift at +dim, stride = row
        # ---- prenet (feat_decoder.py:81-97)
        zpl = self._planes("zq_p", (M, D_in), sp_pre)
        ops.split_f16(zq, zpl) if sp_pre else ops.rows_to_planes(zq, 1, M, D_in, zpl, D_in, M, 0)
        x = self._buf("pre_x", (M, dim))
        ops.gemm(zpl, W["lin_pre"], dim, a_batch=1, a_rows_per_batch=M, a_ld=D_in, m_per_batch=M, bias=W["lin_pre_b"],
                 out_f32=rowmap(x, dim, M, 0))
        for i, bw in enumerate(W["down"]):
            x = self._backbone(bw, x, B, T, dim, inter, tag=f"_d{i}")
        x = self._backbone(W["bb"], x, B, T, dim, inter, cond_rows, n_norm * 2 * dim, tag="_c")
        xpl = self._planes("pre_out_p", (M, dim), sp_pre)
        ops.split_f16(x, xpl) if sp_pre else ops.rows_to_planes(x, 1, M, dim, xpl, dim, M, 0)
        C0 = p["out_channels"]
        pre = self._buf("pre_out", (M, C0))
        ops.gemm(xpl, W["lin"], C0, a_batch=1, a_rows_per_batch=M, a_ld=dim, m_per_batch=M, bias=W["lin_b"],
                 out_f32=rowmap(pre, C0, M, 0))
        if p["use_tanh_at_final"]:
            raise NotImplementedError("prenet use_tanh_at_final (False in the shipped configuration)")
        # ---- WaveGenerator (wave_generator.py:59-91)
        # Every Snake but one per stage rides in a GEMM epilogue: a conv's fp32 output is the residual trunk, its fp16-plane
        # output is Snake_alpha(trunk) written straight into the interior of the NEXT convolution's zero-padded buffer
        # (qb_gemm_desc.act2 / act2_param); the dilated conv's own Snake is `act`.  Only the transposed conv's output - all
        # s phases of a frame in one row, shifted by the padding - needs the stand-alone Snake kernel.
        G = W["gen"]
        a0 = self._planes("gen_in", (B, T + 6, C0), sp_gen)
        ops.addvec_planes(pre, dvec, B, T, C0, a0, C0, T + 6, 3)                    # x + d_vector[:, :, None]  (bicodec.py:197)
        ch = d["channels"]
        stages = G["stages"]

        def up_in(si, Tc):      # padded input buffer of stage si's transposed conv
            st = stages[si]
            cpi = _pad_to(st["cin"], 64)
            return self._planes(f"gen_up_in{si}", (B, Tc + 2 * (st["J"] - 1), cpi), sp_gen), cpi, Tc + 2 * (st["J"] - 1), st["J"] - 1

        Tc = T
        nxt_pl, nxt_ld, nxt_rpb, nxt_off = up_in(0, Tc)
        self._conv(a0, G["conv0"], ch, B, T + 6, C0, T, 7, bias=G["conv0_b"], act2=ACT_SNAKE, act2_param=stages[0]["alpha"],
                   out_planes=nxt_pl, out_planes_map=(nxt_ld, nxt_rpb, nxt_off))
        if taps is not None:
            taps["z_q"], taps["d_vector"], taps["prenet.out"] = zq.clone(), dvec.clone(), pre.clone()
        cl = ch // 2 ** len(d["rates"])
        for si, st in enumerate(stages):
            cin, cout, J, sdn, k = st["cin"], st["cout"], st["J"], st["s"], st["k"]
            cpo = _pad_to(cout, 64)
            a, cpi, rows_in, _ = up_in(si, Tc)
            # transposed conv as a J-tap GEMM producing all `s` phases of every input frame
            up = self._buf(f"gen_up{si}", (B, Tc + J - 1, sdn * cout))
            self._conv(a, st["wt"], sdn * cout, B, rows_in, cpi, Tc + J - 1, J, bias=st["bt"],
                       out_f32=rowmap(up, sdn * cout, Tc + J - 1, 0))
            pad_t = (k - sdn) // 2
            Tn = Tc * sdn
            # the up-sampled clip b is rows [pad_t, pad_t + Tn) of up[b] viewed as [(Tc + J - 1) * s, cout]
            res = rowmap(up, cout, (Tc + J - 1) * sdn, pad_t)
            units = st["units"]
            bufs = [self._planes(f"gen_u{si}_{u['dil']}", (B, Tn + 6 * u["dil"], cpo), sp_gen) for u in units]
            ops.snake_planes(up.view(-1)[pad_t * cout:], (Tc + J - 1) * sdn * cout, units[0]["a1"], B, Tn, cout, bufs[0], cpo,
                             Tn + 6 * units[0]["dil"], 3 * units[0]["dil"])
            a2 = self._planes(f"gen_v{si}", (B * Tn, cpo), sp_gen)
            dense = [self._buf(f"gen_x{si}a", (B * Tn, cout)), self._buf(f"gen_x{si}b", (B * Tn, cout))]
            for ui, un in enumerate(units):
                dil = un["dil"]
                # Snake -> dilated k7 conv -> Snake (epilogue) -> planes
                self._conv(bufs[ui], un["w1"], cout, B, Tn + 6 * dil, cpo, Tn, 7, dilation=dil, bias=un["b1"], act=ACT_SNAKE,
                           act_param=un["a2"], out_planes=a2, out_planes_map=(cpo, Tn, 0))
                # 1x1 conv + residual -> fp32 trunk, and Snake(trunk) planes for whoever consumes it next
                last = ui == len(units) - 1
                if not last:
                    nd = units[ui + 1]["dil"]
                    n_pl, n_ld, n_rpb, n_off, n_alpha = bufs[ui + 1], cpo, Tn + 6 * nd, 3 * nd, units[ui + 1]["a1"]
                elif si + 1 < len(stages):
                    n_pl, n_ld, n_rpb, n_off = up_in(si + 1, Tn)
                    n_alpha = stages[si + 1]["alpha"]
                else:
                    cpf = _pad_to(cl, 64)
                    n_pl, n_ld, n_rpb, n_off, n_alpha = self._planes("gen_f", (B, Tn + 6, cpf), sp_gen), cpf, Tn + 6, 3, G["alpha_f"]
                nxt = dense[ui & 1]
                need_f32 = (not last) or taps is not None
                ops.gemm(a2, un["w2"], cout, a_batch=B, a_rows_per_batch=Tn, a_ld=cpo, m_per_batch=Tn, bias=un["b2"], residual=res,
                         out_f32=rowmap(nxt, cout, Tn, 0) if need_f32 else None, act2=ACT_SNAKE, act2_param=n_alpha,
                         out_planes=n_pl, out_planes_map=(n_ld, n_rpb, n_off))
                res = rowmap(nxt, cout, Tn, 0)
            Tc = Tn
            if taps is not None:
                taps[f"dec.stage{si}"] = dense[(len(units) - 1) & 1].clone().reshape(B, Tc, cout)
        cpf = _pad_to(cl, 64)
        af = self._planes("gen_f", (B, Tc + 6, cpf), sp_gen)
        wav = torch.empty(B, 1, Tc, device=dev)
        self._conv(af, G["conv_f"], 1, B, Tc + 6, cpf, Tc, 7, bias=G["conv_f_b"], act=ACT_TANH, out_f32=rowmap(wav, 1, Tc, 0))
        return wav

    def forward(self, *a, **k):
        raise RuntimeError("unified_audio_b200.BiCodec implements detokenize only (the decoder UniSE uses, model.py:193)")
