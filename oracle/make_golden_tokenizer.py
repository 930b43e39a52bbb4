"""Pin the oracle's wav -> codes chain against the REFERENCE'S OWN `HCodecTokenizer` (QuarkAudio-HCodec/HCodec-2.0/audio_tokenizer.py:21-79).

TEST INFRASTRUCTURE.  Run in the build container only:  python -m oracle.make_golden_tokenizer

The reference's `audio_tokenizer.py` is imported with a stub for `librosa` (absent, unused on this path) and the stand-in for the
un-vendored `vector_quantize_pytorch` oracle/make_golden.py already uses; an `HCodecTokenizer` is built WITHOUT running its `__init__`
(which reads a checkpoint and downloads bosonai/hubert_base) and given
    model             = the reference's own `vq.Codec` (reduced widths, seeded weights),
    feature_extractor = `transformers.HubertModel` (reduced widths, seeded weights - what `AutoModel.from_pretrained` returns),
    resample          = `torchaudio.transforms.Resample(48000, 16000)`,
    hop_length        = 3840;
then its unmodified `pad_wav`, `extract_ssl_features`, `tokenize` and `detokenize` run on a clip whose length is not a multiple of the
hop.  The oracle chain (zero-pad -> oracle/hubert.py `extract_ssl_features` -> oracle/hcodec2.py `codec_encode` / `codec_decode`) is
checked against it and the reference's outputs are written to tests/golden/tokenizer_small.npz for tests/test_host.py.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/QuarkAudio-HCodec/HCodec-2.0"


def import_reference_tokenizer():
    from oracle.make_golden import import_reference_codec
    import_reference_codec()                                   # registers the vector_quantize_pytorch stand-in, puts REF on sys.path
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    spec = importlib.util.spec_from_file_location("ref_audio_tokenizer", os.path.join(REF, "audio_tokenizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.HCodecTokenizer


def main():
    import torchaudio
    from oracle import hcodec2, weights
    from oracle import hubert as oh
    from oracle.make_golden import build_reference
    from oracle.make_golden_hubert import hf_model
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
    c = oh.hubert_small()
    cfg = weights.h2_small()
    cfg["semantic_encoder_config"]["input_channels"] = c["hidden"]
    seed_codec, seed_ssl = 11, 5
    sd = weights.make_h2_state_dict(cfg, seed_codec)
    fsd = oh.make_state_dict(c, seed_ssl)
    HCodecTokenizer = import_reference_tokenizer()
    tok = HCodecTokenizer.__new__(HCodecTokenizer)
    torch.nn.Module.__init__(tok)
    tok.model = build_reference(cfg, sd)
    tok.feature_extractor = hf_model(c, fsd)
    tok.resample = torchaudio.transforms.Resample(cfg["sampling_rate"], 16000)
    tok.hop_length = int(cfg["sampling_rate"] / cfg["encoder_config"]["target_frame_rate"])
    assert tok.hop_length == 3840
    g = torch.Generator().manual_seed(41)
    wav = 0.1 * torch.randn(2, 4 * 3840 - 913, generator=g)
    with torch.no_grad():
        padded_r = tok.pad_wav(wav)
        feats_r = tok.extract_ssl_features(padded_r)
        ac_r, sc_r = tok.tokenize(wav)
        rec_r = tok.detokenize(ac_r, sc_r)
    # the oracle chain
    padded_o = torch.nn.functional.pad(wav, (0, padded_r.shape[-1] - wav.shape[-1]))
    feats_o = oh.extract_ssl_features(fsd, c, padded_o)
    ac_o, sc_o = hcodec2.codec_encode(sd, cfg, padded_o, feats_o.transpose(1, 2))
    rec_o = hcodec2.codec_decode(sd, cfg, ac_r, sc_r)
    big = feats_r.abs() > 0.2 * feats_r.abs().max()            # sign(x)|x|^0.3 has an infinite slope at 0: compare away from it
    report = dict(padded_equal=bool(torch.equal(padded_r, padded_o)), frames=int(feats_r.shape[1]),
                  feats_rel_away_from_zero=float(((feats_r - feats_o).abs()[big]).max() / feats_r.abs().max()),
                  acoustic_identical=bool(torch.equal(ac_r, ac_o)), semantic_identical=bool(torch.equal(sc_r, sc_o)),
                  rec_rel=rel(rec_o, rec_r), codes_shape=list(ac_r.shape), rec_shape=list(rec_r.shape))
    print(report)
    assert report["padded_equal"] and report["feats_rel_away_from_zero"] < 1e-4 and report["acoustic_identical"] and report["semantic_identical"]
    assert report["rec_rel"] < 1e-5 and rec_r.shape[-1] == padded_r.shape[-1]
    meta = dict(hubert=c, codec_cfg=cfg, seed_codec=seed_codec, seed_ssl=seed_ssl, report=report,
                reference="QuarkAudio-HCodec/HCodec-2.0/audio_tokenizer.py:47-79 (unmodified pad_wav / extract_ssl_features / tokenize / detokenize)")
    np.savez_compressed(os.path.join(GOLD, "tokenizer_small.npz"), wav=wav.numpy(), feats=feats_r.numpy(), acoustic=ac_r.numpy(),
                        semantic=sc_r.numpy(), rec=rec_r.numpy(), meta=np.array(json.dumps(meta)))
    print("wrote tokenizer_small.npz", os.path.getsize(os.path.join(GOLD, "tokenizer_small.npz")), "bytes")


if __name__ == "__main__":
    main()
