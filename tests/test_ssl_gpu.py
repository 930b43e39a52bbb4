"""SSL front end on the GPU and the tokenizer / segmenting glue (8f.3) against the oracle (oracle/hubert.py, pinned
against transformers.HubertModel and torchaudio's Resample by tests/golden/hubert_small.npz)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def gpu_small():
    """reduced widths the kernels support (head_dim 64; the committed transformers fixture uses head_dim 16)"""
    return dict(conv_dim=[64] * 7, conv_kernel=[10, 3, 3, 3, 3, 2, 2], conv_stride=[5, 2, 2, 2, 2, 2, 2], hidden=128, layers=2, heads=2,
                ffn=256, pos_k=16, pos_groups=4, eps=1e-5)


def _build(c, seed, in_rate, compress):
    from oracle import hubert as oh
    from unified_audio_b200.ssl import SSLFrontEnd
    sd = oh.make_state_dict(c, seed)
    m = SSLFrontEnd(dict(c, kind="hubert"), in_rate=in_rate, compress=compress)
    m.load_state_dict(sd, strict=True)
    return m.cuda(), sd


def _decompress(y):
    return torch.sign(y) * y.abs() ** (1 / 0.3)


@pytest.mark.parametrize("cfg_name,B,seconds", [("small", 3, 0.5), ("base", 2, 1.0)])
def test_hubert_front_end_vs_oracle(lib, cfg_name, B, seconds):
    from oracle import hubert as oh
    c = gpu_small() if cfg_name == "small" else oh.HUBERT_BASE
    m, sd = _build(c, 5, 48000, True)
    g = torch.Generator().manual_seed(77)
    T48 = int(48000 * seconds)
    wav48 = 0.1 * torch.randn(B, T48, generator=g)
    # resampler == torchaudio's polyphase sinc filter
    rs = m.resample(wav48.cuda())
    torch.cuda.synchronize()
    e_rs = rel(rs, oh.resample(wav48))
    assert rs.shape[1] == math.ceil(T48 / 3) and e_rs < 1e-5, e_rs
    taps = {}
    feats = m(wav48.cuda(), taps=taps)
    torch.cuda.synchronize()
    w16 = torch.nn.functional.pad(oh.resample(wav48), (160, 160))
    ref_feats = oh.feature_encoder(sd, c, w16).transpose(1, 2)
    hs = oh.hubert_hidden_states(sd, c, w16)
    ref_mean = torch.stack(hs, 1).mean(1)
    e_f, e_0, e_l, e_m = rel(taps["features"], ref_feats), rel(taps["hs0"], hs[0]), rel(taps[f"hs{c['layers']}"], hs[-1]), rel(taps["mean"], ref_mean)
    # sign(x)|x|^0.3 has an infinite slope at 0: compare after undoing the compression (the reference's own fp32 noise moves
    # a 1e-6 entry by 1.6e-2 after compression), and directly away from zero
    ref_out = oh.extract_ssl_features(sd, c, wav48)
    e_c = rel(_decompress(feats), _decompress(ref_out))
    big = ref_mean.abs() > 1e-2 * ref_mean.abs().max()
    e_big = float(((feats.cpu() - ref_out).abs()[big]).max() / ref_out.abs().max())
    print(f"[hubert {cfg_name} B={B} {seconds}s] resample {e_rs:.1e} conv features {e_f:.2e} hs0 {e_0:.2e} last {e_l:.2e} mean {e_m:.2e}; "
          f"compressed: decompressed rel {e_c:.2e}, direct (|x| > 1% of max) {e_big:.2e}; frames {feats.shape[1]}")
    assert feats.shape == ref_out.shape
    assert max(e_f, e_0, e_l, e_m, e_c, e_big) < TOL
    cf = m(wav48.cuda(), channel_first=True)
    torch.cuda.synchronize()
    assert torch.equal(cf.transpose(1, 2), feats)


def test_tokenizer_glue_and_end_to_end(lib):
    """pad_wav / wrap_segments on the device == the reference's host glue; HCodecTokenizer.tokenize (wav -> codes) end to end
    against the oracle chain (reduced widths)."""
    from oracle import hcodec2, hubert as oh, weights
    from oracle.parity import audit_codes
    from unified_audio_b200.codec import Codec
    from unified_audio_b200.ssl import HCodecTokenizer, pad_wav, wrap_segments
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 10001, generator=g)
    p = pad_wav(x.cuda(), 3840)
    assert p.shape == (2, 11520) and torch.equal(p[:, :10001].cpu(), x) and float(p[:, 10001:].abs().max()) == 0
    src = torch.randn(1, 23000, generator=g)
    seg = wrap_segments(src.cuda(), 8000)
    want = torch.from_numpy(np.pad(src.numpy(), [(0, 0), (0, 1000)], "wrap")).reshape(-1, 8000)
    assert torch.equal(seg.cpu(), want)
    # end to end: a codec whose semantic encoder takes the SSL width
    cfg = weights.h2_small()
    c = gpu_small()
    cfg["semantic_encoder_config"]["input_channels"] = c["hidden"]
    sd = weights.make_h2_state_dict(cfg, 11)
    codec = Codec(cfg["encoder_config"], cfg["decoder_config"], cfg["quantizer_config"], cfg["semantic_encoder_config"],
                  cfg["semantic_decoder_config"])
    codec.load_state_dict(sd)
    fe, fsd = _build(c, 5, 48000, True)
    tok = HCodecTokenizer(codec.cuda(), fe, 48000, 12.5)
    wav = 0.1 * torch.randn(2, 3 * 3840 - 700, generator=g)
    ac, sc = tok.tokenize(wav.cuda())
    torch.cuda.synchronize()
    wp = torch.nn.functional.pad(wav, (0, 700))
    feats = oh.extract_ssl_features(fsd, c, wp).transpose(1, 2)
    otaps = {}
    oa, os_ = hcodec2.codec_encode(sd, cfg, wp, feats, taps=otaps)
    assert ac.shape == oa.shape == (2, 4, 3)
    rec = tok.detokenize(ac, sc)
    torch.cuda.synchronize()
    print(f"[tokenize] acoustic match {float((ac.cpu() == oa).float().mean()):.3f} semantic match {float((sc.cpu() == os_).float().mean()):.3f}; "
          f"rec {tuple(rec.shape)}")
    assert rec.shape == (2, 3 * 3840) and float((ac.cpu() == oa).float().mean()) > 0.9 and float((sc.cpu() == os_).float().mean()) > 0.9


def test_wavlm_front_end_vs_oracle(lib):
    """WavLM-base-plus (gated relative position bias attention) as UniSE uses it (U/model/model.py:38-51): 16 kHz in, no
    compression; reduced widths and the full configuration against the oracle (pinned against transformers.WavLMModel)."""
    from oracle import hubert as oh
    from unified_audio_b200.ssl import SSLFrontEnd
    for name, c, B, T in (("small", dict(gpu_small(), num_buckets=32, max_distance=80), 3, 8000), ("base-plus", oh.WAVLM_BASE_PLUS, 2, 16000)):
        sd = oh.wavlm_make_state_dict(c, 8)
        m = SSLFrontEnd(dict(c, kind="wavlm"), in_rate=16000, compress=False)
        m.load_state_dict(sd, strict=True)
        m = m.cuda()
        wav = 0.1 * torch.randn(B, T, generator=torch.Generator().manual_seed(5))
        taps = {}
        got = m(wav.cuda(), taps=taps)
        torch.cuda.synchronize()
        hs = oh.wavlm_hidden_states(sd, c, torch.nn.functional.pad(wav, (160, 160)))
        ref = oh.extract_semantic_features(sd, c, wav)
        e1, el, e = rel(taps["hs1"], hs[1]), rel(taps[f"hs{c['layers']}"], hs[-1]), rel(got, ref)
        print(f"[wavlm {name} B={B}] first layer {e1:.2e} last {el:.2e} mean of hidden states {e:.2e}; frames {got.shape[1]}")
        assert got.shape == ref.shape and max(e1, el, e) < TOL
