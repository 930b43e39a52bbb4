"""Pin oracle/hcodec15.py against the reference's own H-Codec-1.5 modules and write tests/golden/h15_*.npz.

TEST INFRASTRUCTURE - runs ONLY in the build container (needs /root/reference).   python -m oracle.make_golden_h15

 1. registers oracle.rvq.ResidualVQ under `vector_quantize_pytorch` (the un-vendored dependency, as oracle/make_golden.py does)
    and stubs funasr / dac / easydict / audiotools (module-level imports of adaptive/modeling_flexicodec_new.py that the codec
    path never touches), sets NO_TORCH_COMPILE=1 (module/compile.py:41) and imports HCodec-1.5's `vq.codec_adaptive.Codec`;
 2. builds it from conf/config_adaptive_v3.yaml (optionally with fewer layers), checks state_dict keys / shapes against
    oracle.hcodec15.param_specs and loads the seeded weights;
 3. runs Codec.encode / Codec.decode (vq/codec_adaptive.py:150-207) and the two aggregators / bottleneck on seeded inputs whose
    semantic features come in runs (so that groups of 1..8 frames and the 8-frame cap all occur);
 4. asserts the restatement reproduces the reference and stores the REFERENCE outputs.
Also re-derives tests/golden/h1_full_1s.npz (H-Codec-1.0, round 1) from HCodec-1.0's own `vq.Codec` and asserts the committed
fixture is what the reference produces (the round-1 generator was not committed).
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch
from torch import nn

os.environ["NO_TORCH_COMPILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF15 = "/root/reference/QuarkAudio-HCodec/HCodec-1.5"
REF10 = "/root/reference/QuarkAudio-HCodec/HCodec-1.0"
GOLD = os.path.join(ROOT, "tests", "golden")


def _stubs():
    from oracle import rvq

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return m
    stub("vector_quantize_pytorch", ResidualVQ=rvq.ResidualVQ, ResidualSimVQ=object, ResidualFSQ=object)
    stub("funasr", AutoModel=object)
    stub("easydict", EasyDict=dict)
    stub("dac"); stub("dac.nn"); stub("dac.model")
    stub("dac.nn.layers", WNConv1d=lambda *a, **k: nn.utils.weight_norm(nn.Conv1d(*a, **k)), Snake1d=nn.Identity)
    stub("dac.model.base", CodecMixin=object)
    stub("audiotools", AudioSignal=object)
    stub("audiotools.ml", BaseModel=nn.Module)


def _purge(prefixes):
    for k in [k for k in sys.modules if k.split(".")[0] in prefixes]:
        del sys.modules[k]


def reference_kwargs(c):
    """conf/config_adaptive_v3.yaml with this config's depths"""
    import yaml
    y = yaml.safe_load(open(os.path.join(REF15, "conf", "config_adaptive_v3.yaml")))
    y["decoder_config"]["decoder"].update(dim=c["dec_dim"], intermediate_dim=c["dec_inter"], convnext_layers=c["dec_layers"])
    a = y["adaptive_config"]
    for k in ("semantic_aggregator", "acoustic_aggregator"):
        a["aggregators"][k]["num_layers"] = c["agg"]["layers"]
    a["transformer_kwargs"]["num_layers"] = c["bottleneck"]["layers"]
    assert a["manual_threshold"] == c["threshold"] and a["max_tokens_per_group"] == c["max_group"]
    assert y["encoder_config"]["encoder"]["ratios"] == list(reversed(c["ratios"]))
    return y


def build_reference(c, sd):
    _stubs()
    _purge({"vq", "adaptive"})
    sys.path.insert(0, REF15)
    try:
        from vq.codec_adaptive import Codec
        y = reference_kwargs(c)
        ref = Codec(y["encoder_config"], y["decoder_config"], y["quantizer_config"], y["adaptive_config"]).eval()
    finally:
        sys.path.remove(REF15)
    theirs = {k: tuple(v.shape) for k, v in ref.state_dict().items() if not k.startswith("semantic_decoder.")}
    ours = {k: tuple(v.shape) for k, v in sd.items()}
    assert ours == theirs, (sorted(set(ours) ^ set(theirs))[:20], [k for k in ours if k in theirs and ours[k] != theirs[k]][:20])
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("semantic_decoder.") for k in missing)
    return ref


def synth(c, batch, n_frames25, seed):
    """wav [B,1,T] @16 kHz and 50 Hz features in runs of 1..14 frames + noise, compressed like H2/audio_tokenizer.py:58-60"""
    g = torch.Generator().manual_seed(seed)
    T50 = 2 * n_frames25
    wav = 0.1 * torch.randn(batch, 1, T50 * 320, generator=g)
    feats = []
    for _ in range(batch):
        cols = []
        while sum(x.shape[1] for x in cols) < T50:
            run = int(torch.randint(1, 15, (1,), generator=g))
            cols.append(torch.randn(c["sem_in"], 1, generator=g).expand(-1, run))
        f = torch.cat(cols, 1)[:, :T50] + 0.25 * torch.randn(c["sem_in"], T50, generator=g)
        feats.append(f)
    f = torch.stack(feats)
    return wav, torch.sign(f) * f.abs() ** 0.3


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def run(name, c, batch, n_frames25, seed_w, seed_x):
    from oracle import adaptive as ad
    from oracle import hcodec15 as o15
    sd = o15.make_state_dict(c, seed_w)
    ref = build_reference(c, sd)
    wav, feat = synth(c, batch, n_frames25, seed_x)
    with torch.no_grad():
        emb_ref = ref.encoder(wav)
        sem_ref = ref.semantic_encoder(feat)
        from adaptive.modeling_flexicodec_new import FlexiCodec
        al_ref, _, n_ref = FlexiCodec._perform_similarity_alignment_vectorized(
            sem_ref.transpose(1, 2), x_lens=torch.full((batch,), sem_ref.shape[-1]), current_threshold=c["threshold"],
            max_tokens_per_group=c["max_group"])
        sem_tok_ref = ref.semantic_aggregator(sem_ref, al_ref, n_ref)
        ac_tok_ref = ref.acoustic_aggregator(emb_ref, al_ref, n_ref)
        out = ref.encode(wav, feat)
        ac_ref, sc_ref = out["acoustic_codes"], out["semantic_codes"]
        rec_ref = ref.decode(ac_ref, sc_ref)
        plain, lens = ref._extract_length_from_codes_index(ac_ref)
        z_ref = torch.cat([ref.quantizer.get_output_from_indices(
            FlexiCodec._deaggregate_features_from_token_lengths(plain, lens).transpose(1, 2)).transpose(1, 2),
            ref.semantic_quantizer.get_output_from_indices(FlexiCodec._deaggregate_features_from_token_lengths(
                ref._extract_length_from_codes_index(sc_ref)[0], lens).transpose(1, 2)).transpose(1, 2)], 1)
        bn_ref = ref.bottleneck_transformer(z_ref)
    taps = {}
    ac, sc = o15.codec_encode(sd, c, wav, feat, taps)
    dtaps = {}
    rec = o15.codec_decode(sd, c, ac_ref, sc_ref, dtaps)
    lens_hist = torch.bincount(ad.token_lengths(al_ref).flatten(), minlength=9).tolist()
    report = dict(
        name=name, batch=batch, frames=n_frames25, seed_w=seed_w, seed_x=seed_x,
        groups_per_item=n_ref.tolist(), token_length_histogram=lens_hist,
        align_equal=bool(torch.equal(taps["align"], al_ref)),
        emb_rel=rel(taps["enc.out"], emb_ref), sem_rel=rel(taps["sem.out"], sem_ref),
        sem_tok_rel=rel(taps["sem_agg.out"], sem_tok_ref), ac_tok_rel=rel(taps["ac_agg.out"], ac_tok_ref),
        bottleneck_rel=rel(dtaps["bottleneck.out"], bn_ref), rec_rel=rel(rec, rec_ref),
        acoustic_codes_equal=bool(torch.equal(ac, ac_ref)), semantic_codes_equal=bool(torch.equal(sc, sc_ref)),
        acoustic_index_match=float((ac == ac_ref).float().mean()), semantic_index_match=float((sc == sc_ref).float().mean()),
        negative_codes_in_padded_groups=int((ac_ref < 0).sum()),
    )
    print(json.dumps(report))
    assert report["align_equal"]
    for k in ("emb_rel", "sem_rel", "sem_tok_rel", "ac_tok_rel", "bottleneck_rel", "rec_rel"):
        assert report[k] < 2e-5, (k, report[k])
    assert report["acoustic_index_match"] > 0.995 and report["semantic_index_match"] > 0.995
    np.savez_compressed(
        os.path.join(GOLD, f"h15_{name}.npz"),
        meta=json.dumps(dict(name=name, batch=batch, frames=n_frames25, seed_w=seed_w, seed_x=seed_x)),
        emb=emb_ref.numpy(), sem=sem_ref.numpy(), align=al_ref.numpy().astype(np.uint8), n_groups=n_ref.numpy(),
        sem_tok=sem_tok_ref.numpy(), ac_tok=ac_tok_ref.numpy(), acoustic_codes=ac_ref.numpy(), semantic_codes=sc_ref.numpy(),
        z=z_ref.numpy(), bottleneck=bn_ref.numpy(), wav_rec=rec_ref.numpy())
    return report


def check_h1_fixture():
    """tests/golden/h1_full_1s.npz must be what HCodec-1.0's own vq.Codec produces from the seeds in its meta"""
    from oracle import hcodec1
    _stubs()
    _purge({"vq", "adaptive"})
    sys.path.insert(0, REF10)
    try:
        from vq.codec import Codec
        ref = Codec({}, {}, {}).eval()
    finally:
        sys.path.remove(REF10)
    z = np.load(os.path.join(GOLD, "h1_full_1s.npz"))
    meta = json.loads(str(z["meta"]))
    sd = hcodec1.make_state_dict(hcodec1.H1, meta["seed_w"])
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("semantic_decoder.") for k in missing), (missing[:5], unexpected[:5])
    g = torch.Generator().manual_seed(meta["seed_x"])
    x = 0.1 * torch.randn(1, 1, 16000, generator=g)
    g2 = torch.Generator().manual_seed(meta["seed_x"] + 1)
    f = torch.randn(1, 768, 50, generator=g2)
    feat = torch.sign(f) * f.abs() ** 0.3
    with torch.no_grad():
        out = ref.encode(x, feat)
        ac, sc = (out["acoustic_codes"], out["semantic_codes"]) if isinstance(out, dict) else out
        rec = ref.decode(ac, sc)
    ok = dict(acoustic=bool(np.array_equal(ac.numpy(), z["acoustic_codes"])), semantic=bool(np.array_equal(sc.numpy(), z["semantic_codes"])),
              wav_rel=rel(rec.reshape(-1), torch.from_numpy(z["wav_rec"]).reshape(-1)))
    print("h1_full_1s fixture vs HCodec-1.0 reference:", ok)
    assert ok["acoustic"] and ok["semantic"] and ok["wav_rel"] < 1e-5
    return ok


def main():
    from oracle import hcodec15 as o15
    torch.set_num_threads(16)
    reports = [run("shallow", o15.h15_shallow(), 3, 48, 11, 12), run("full", o15.H15, 2, 40, 21, 22)]
    h1 = check_h1_fixture()
    json.dump(dict(reports=reports, h1_full_1s=h1), open(os.path.join(GOLD, "h15_pinning_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
