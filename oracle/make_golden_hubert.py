"""Pin oracle/hubert.py against transformers.HubertModel and torchaudio.transforms.Resample; write the fixture.

Run in the build container only:  python -m oracle.make_golden_hubert
HuBERT weights are not available offline: both sides load the SAME seeded random weights (the architecture is what is
pinned).  Outputs: tests/golden/hubert_small.npz, tests/golden/hubert_pinning_report.json.
"""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hf_model(c, sd):
    from transformers import HubertConfig, HubertModel
    cfg = HubertConfig(hidden_size=c["hidden"], num_hidden_layers=c["layers"], num_attention_heads=c["heads"],
                       intermediate_size=c["ffn"], conv_dim=tuple(c["conv_dim"]), conv_kernel=tuple(c["conv_kernel"]),
                       conv_stride=tuple(c["conv_stride"]), num_conv_pos_embeddings=c["pos_k"],
                       num_conv_pos_embedding_groups=c["pos_groups"], feat_extract_norm="group", conv_bias=False,
                       do_stable_layer_norm=False, hidden_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0,
                       layerdrop=0.0, mask_time_prob=0.0)
    m = HubertModel(cfg).eval()
    own = m.state_dict()
    extra = set(own) - set(sd) - {"masked_spec_embed"}
    assert not extra, extra
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {"masked_spec_embed"}, (missing, unexpected)
    return m


def main():
    import torchaudio
    from oracle import hubert as oh
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    report = {}
    # ---- resampler
    g = torch.Generator().manual_seed(3)
    w48 = 0.1 * torch.randn(2, 48000 + 123, generator=g)
    ref = torchaudio.transforms.Resample(48000, 16000)(w48)
    got = oh.resample(w48)
    report["resample_48k_16k"] = dict(rel_err=rel(got, ref), shape=list(ref.shape))
    assert got.shape == ref.shape and report["resample_48k_16k"]["rel_err"] < 1e-6
    # ---- HuBERT
    for name, c, B, T, seed in (("small", oh.hubert_small(), 2, 16000, 5), ("base", oh.HUBERT_BASE, 1, 8000, 6)):
        sd = oh.make_state_dict(c, seed)
        m = hf_model(c, sd)
        gw = torch.Generator().manual_seed(seed + 50)
        wav = 0.1 * torch.randn(B, T, generator=gw)
        with torch.no_grad():
            out = m(wav, output_hidden_states=True)
        hs = oh.hubert_hidden_states(sd, c, wav)
        errs = [rel(a, b) for a, b in zip(hs, out.hidden_states)]
        assert len(hs) == len(out.hidden_states) == c["layers"] + 1
        mix_ref = torch.stack(out.hidden_states, 1).mean(1)
        report[name] = dict(max_hidden_state_rel_err=max(errs), frames=int(hs[0].shape[1]), hidden_states=len(hs))
        print(name, report[name])
        assert max(errs) < 2e-5
        if name == "small":
            w48s = 0.1 * torch.randn(2, 48000, generator=gw)
            feats = oh.extract_ssl_features(sd, c, w48s)
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hubert_small.npz"), meta=json.dumps(dict(seed=seed)),
                                wav=wav.numpy(), mix=mix_ref.numpy(), last=out.last_hidden_state.numpy(),
                                wav48=w48s.numpy(), feats=feats.numpy(), resampled=torchaudio.transforms.Resample(48000, 16000)(w48s).numpy())
    # ---- WavLM (UniSE's semantic model)
    from transformers import WavLMConfig, WavLMModel
    for name, c, B, T, seed in (("wavlm_small", oh.wavlm_small(), 2, 12000, 8), ("wavlm_base_plus", oh.WAVLM_BASE_PLUS, 1, 8000, 9)):
        sd = oh.wavlm_make_state_dict(c, seed)
        cfg = WavLMConfig(hidden_size=c["hidden"], num_hidden_layers=c["layers"], num_attention_heads=c["heads"],
                          intermediate_size=c["ffn"], conv_dim=tuple(c["conv_dim"]), conv_kernel=tuple(c["conv_kernel"]),
                          conv_stride=tuple(c["conv_stride"]), num_conv_pos_embeddings=c["pos_k"],
                          num_conv_pos_embedding_groups=c["pos_groups"], num_buckets=c["num_buckets"],
                          max_bucket_distance=c["max_distance"], feat_extract_norm="group", conv_bias=False,
                          do_stable_layer_norm=False, hidden_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0,
                          layerdrop=0.0, mask_time_prob=0.0)
        m = WavLMModel(cfg).eval()
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and set(missing) <= {"masked_spec_embed"}, (missing, unexpected)
        assert set(m.state_dict()) - {"masked_spec_embed"} == set(sd)
        gw = torch.Generator().manual_seed(seed + 50)
        wav = 0.1 * torch.randn(B, T, generator=gw)
        with torch.no_grad():
            out = m(wav, output_hidden_states=True)
        hs = oh.wavlm_hidden_states(sd, c, wav)
        errs = [rel(a, b) for a, b in zip(hs, out.hidden_states)]
        report[name] = dict(max_hidden_state_rel_err=max(errs), frames=int(hs[0].shape[1]), hidden_states=len(hs))
        print(name, report[name])
        assert len(hs) == len(out.hidden_states) and max(errs) < 2e-5
        if name == "wavlm_small":
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wavlm_small.npz"), meta=json.dumps(dict(seed=seed)),
                                wav=wav.numpy(), mix=torch.stack(out.hidden_states, 1).mean(1).numpy())
    json.dump(report, open(os.path.join(ROOT, "tests", "golden", "hubert_pinning_report.json"), "w"), indent=1)
    print(report)


if __name__ == "__main__":
    main()
