"""UniSE inference surface (unified_audio_b200.unise.Model.test_step, U/model/model.py:170-286) on the GPU against the chain of
oracles the reference's own path is restated by: wrap-pad + 5 s segmenting (NumPy, as the reference does it) -> WavLM mean hidden
state (oracle/hubert.py) -> LLM_SFT.generate (oracle/llama.py) -> BiCodec.detokenize (oracle/bicodec.py).  Reduced widths, the
reference's fixed 5 s segment length."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3
SEG = 5 * 16000


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def build():
    from oracle import bicodec as ob
    from oracle import hubert as oh
    from oracle import llama
    from unified_audio_b200.bicodec import BiCodec
    from unified_audio_b200.llm import LLM_SFT
    from unified_audio_b200.ssl import SSLFrontEnd
    from unified_audio_b200.unise import BiCodecTokenizer, Model
    c = dict(conv_dim=[64] * 7, conv_kernel=[10, 3, 3, 3, 3, 2, 2], conv_stride=[5, 2, 2, 2, 2, 2, 2], hidden=128, layers=2, heads=2,
             ffn=256, pos_k=16, pos_groups=4, eps=1e-5, num_buckets=32, max_distance=80)
    wsd = oh.wavlm_make_state_dict(c, 8)
    wavlm = SSLFrontEnd(dict(c, kind="wavlm"), in_rate=16000, compress=False)
    wavlm.load_state_dict(wsd, strict=True)
    bc = ob.bicodec_small()
    bc["speaker"]["token_num"] = 32                 # the shipped count: all 32 generated global tokens feed the decoder
    bsd = ob.make_state_dict(bc, 4)
    codec = BiCodec(bc, precision="accurate")
    codec.load_state_dict(bsd, strict=True)
    lcfg = llama.lm_small(hidden=128, layers=2, heads=2, gsize=4096, ssize=bc["quantizer"]["codebook_size"], feats=c["hidden"])
    lsd = llama.make_lm_state_dict(lcfg, 3, 2.0)
    lm = LLM_SFT(num_tasks=lcfg["num_tasks"], task_map=lcfg["task_map"], feats_dim=lcfg["feats_dim"], llm_base_config=lcfg["llm_base_config"])
    lm.load_state_dict(lsd, strict=True)
    model = Model(dict(stft_config=dict(hop_length=320, win_length=640, n_fft=640, n_mels=80)), tokenizer=BiCodecTokenizer(codec.cuda()),
                  dnn=lm.cuda(), semantic_model=wavlm.cuda())
    return model, dict(c=c, wsd=wsd, bc=bc, bsd=bsd, lcfg=lcfg, lsd=lsd)


def ref_segments(src, normalise):
    pad_len = math.ceil(src.size(-1) / SEG) * SEG - src.size(-1)
    seg = torch.from_numpy(np.pad(src.numpy(), [(0, 0), (0, pad_len)], "wrap")).reshape(-1, SEG)      # model.py:176-180
    return seg / src.abs().max(dim=-1, keepdim=True)[0] if normalise else seg


def check_tokens(tag, got, want, margins, thr=1e-4):
    """first differing step per sequence must be a numerically unsafe decision of the oracle (tiny top-2 logit margin)"""
    nbad = 0
    for b in range(want.shape[0]):
        d = (got[b] != want[b]).nonzero()
        if len(d):
            nbad += 1
            assert float(margins[b, int(d[0])]) < thr, f"{tag}: token differs although the oracle's decision margin is safe"
    print(f"[{tag}] sequences with a differing token: {nbad}/{want.shape[0]}")


@pytest.mark.parametrize("mode", ["se", "tse"])
def test_unise_test_step_vs_oracle_chain(lib, mode):
    from oracle import bicodec as ob
    from oracle import hubert as oh
    from oracle import llama
    model, o = build()
    g = torch.Generator().manual_seed(31)
    T = 2 * SEG - 12345                                       # two segments, the second one wrap-padded
    src = 0.1 * torch.randn(1, T, generator=g)
    enroll = 0.1 * torch.randn(1, SEG // 2, generator=g) if mode == "tse" else None
    est, gids, sids = model.enhance(mode, None if enroll is None else enroll.cuda(), src.cuda(), return_ids=True)
    torch.cuda.synchronize()
    assert est.shape == (T,) and gids.shape == (2, 32) and sids.shape == (2, SEG // 320)
    # the reference's path on the oracles
    seg = ref_segments(src, normalise=mode == "se")
    feats = oh.extract_semantic_features(o["wsd"], o["c"], seg)
    e_feat = rel(model.extract_semantic_features(seg.cuda()), feats)
    efeats = None
    if enroll is not None:
        efeats = torch.cat([oh.extract_semantic_features(o["wsd"], o["c"], enroll)] * seg.size(0), 0)
    og, os_, margins = llama.sft_generate(o["lsd"], o["lcfg"], mode, efeats, feats, SEG // 320, return_margins=True)
    got = torch.cat([gids.cpu(), sids.cpu()], 1)
    want = torch.cat([og, os_], 1)
    check_tokens(mode, got, want, torch.cat([margins[:, :32], margins[:, 33:]], 1))
    # waveform from the tokens the GPU produced, on the BiCodec oracle
    wav = ob.detokenize(o["bsd"], o["bc"], sids.cpu(), gids.cpu()[:, None, :]).squeeze(1).reshape(-1)[:T]
    e_wav = rel(est, wav)
    print(f"[unise {mode}] WavLM feats rel {e_feat:.2e}  tokens identical {bool((got == want).all())}  waveform rel {e_wav:.2e}")
    assert e_feat < TOL and e_wav < TOL


def test_unise_surface(lib):
    """test_step's batch tuple and return value, the 'ss' chain (se -> tse -> rtse), the shape-only mel against stft_logmel, the
    checkpoint surface (state_dict holds the LM only, under `dnn.`) and the refusal to run off the GPU."""
    model, o = build()
    g = torch.Generator().manual_seed(32)
    src = 0.1 * torch.randn(1, SEG + 4000, generator=g)
    out = model.test_step(("se", None, src.cuda(), None, [16000], None, ["utt"]), 0)
    assert isinstance(out, np.ndarray) and out.shape == (SEG + 4000,) and np.isfinite(out).all()
    est = model.enhance("se", None, src.cuda())
    assert np.array_equal(est.cpu().numpy(), out)                    # deterministic (greedy)
    s1, s2 = model.enhance("ss", None, src.cuda())
    torch.cuda.synchronize()
    assert s1.shape == s2.shape == (SEG + 4000,) and bool(torch.isfinite(s1).all()) and bool(torch.isfinite(s2).all())
    short = src[:, :30000].cuda()                                    # shorter than one segment: wrapped up to 5 s first
    a, b = model.enhance("ss", None, short)
    assert a.shape == b.shape == (30000,)
    x = src[:, :48000].cuda()
    assert model.mel_like(x).shape == model.stft_logmel(x).shape == (1, 150, 80)
    keys = list(model.state_dict().keys())
    assert keys and all(k.startswith("dnn.") for k in keys)
    model.load_state_dict({k: v for k, v in model.state_dict().items()})
    with pytest.raises(RuntimeError):
        model.enhance("se", None, src)                               # CPU tensor: no fallback
    with pytest.raises(ValueError):
        model.enhance("bogus", None, src.cuda())
    with pytest.raises(NotImplementedError):
        model.tokenizer.tokenize(src.cuda())
