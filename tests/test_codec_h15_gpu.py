"""H-Codec-1.5 adaptive frame-rate codec on the GPU against the golden outputs of the reference's own modules
(tests/golden/h15_*.npz, written by oracle/make_golden_h15.py) and against the oracle's intermediate taps."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


# the hyper-parameters of HCodec-1.5/conf/config_adaptive_v3.yaml that the reference's constructors read
_AGG = dict(dim=512, in_out_dim=512, num_heads=8, num_layers=32, dim_feedforward=2048, causal=False, use_mean_pooling_init=True,
            context_frames=16)
SHIPPED_CONFIG = dict(
    encoder_config=dict(encoder=dict(n_filters=32, dimension=512, ratios=[2, 4, 5, 8]),
                        semantic_encoder=dict(input_channels=1024, encode_channels=1024, out_channels=512, strides=[2, 1])),
    decoder_config=dict(decoder=dict(input_channels=1024, dim=1024, intermediate_dim=2304)),
    quantizer_config=dict(quantizer=dict(dim=512, codebook_size=1024, num_quantizers=4)),
    adaptive_config=dict(use_similarity_alignment=True, similarity_threshold=0.7, max_tokens_per_group=8, manual_threshold=0.6,
                         use_query_token_aggregator=True, use_bottleneck_transformer=True,
                         aggregators=dict(semantic_aggregator=dict(_AGG), acoustic_aggregator=dict(_AGG)),
                         transformer_kwargs=dict(d_model=1024, num_heads=8, num_layers=32, causal=False, layer_scale=0.01, context=16,
                                                 conv_layout=True, gating="none", norm="layer_norm", positional_embedding="rope",
                                                 dim_feedforward=2048, input_dimension=1024, output_dimensions=[1024])))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def _setup(name, precision):
    from oracle import hcodec15 as o15
    from oracle.make_golden_h15 import synth
    from unified_audio_b200.codec_h15 import CodecH15
    z = np.load(os.path.join(GOLD, f"h15_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    c = o15.h15_shallow() if name == "shallow" else o15.H15
    sd = o15.make_state_dict(c, meta["seed_w"])
    cfg = {k: v for k, v in c.items() if k != "layer_scale"}
    m = CodecH15(precision=precision, _cfg=cfg)
    m.load_state_dict(sd, strict=True)
    wav, feat = synth(c, meta["batch"], meta["frames"], meta["seed_x"])
    return z, c, sd, m.cuda(), wav, feat


def _codebooks(sd, name, nq):
    return torch.stack([sd[f"{name}.layers.{i}._codebook.embed"][0] for i in range(nq)], 0)


@pytest.mark.parametrize("name,precision", [("shallow", "mixed"), ("shallow", "accurate"), ("full", "mixed")])
def test_h15_encode_decode_against_reference_golden(lib, name, precision):
    from oracle import adaptive as oad
    from oracle.parity import audit_codes
    z, c, sd, m, wav, feat = _setup(name, precision)
    B, K, nq = wav.shape[0], c["codebook_size"], c["nq"]
    taps = {}
    out = m.encode(wav.cuda(), feat.cuda(), taps=taps)
    torch.cuda.synchronize()
    e_emb, e_sem = rel(taps["enc.out"], torch.from_numpy(z["emb"])), rel(taps["sem.out"], torch.from_numpy(z["sem"]))
    print(f"[h15 {name}/{precision}] emb rel {e_emb:.2e} sem rel {e_sem:.2e}")
    assert e_emb < TOL and e_sem < TOL
    # grouping: the frame -> token map must reproduce the reference's alignment matrix (a flip needs a cosine similarity within
    # float error of the threshold; report the closest one)
    align = torch.from_numpy(z["align"]).float()
    seg_ref = align.argmax(1)
    margin = float((taps["sim"].cpu() - c["threshold"]).abs().min())
    print(f"[h15 {name}] groups per item {taps['n_groups'].tolist()} (reference {z['n_groups'].tolist()}); closest similarity to the "
          f"threshold: {margin:.2e}")
    assert torch.equal(taps["seg"].cpu().long(), seg_ref), "grouping differs from the reference"
    assert torch.equal(taps["n_groups"].cpu(), torch.from_numpy(z["n_groups"]))
    assert torch.equal(taps["token_lengths"].cpu(), oad.token_lengths(align))
    e_st, e_at = rel(taps["sem_agg.out"], torch.from_numpy(z["sem_tok"])), rel(taps["ac_agg.out"], torch.from_numpy(z["ac_tok"]))
    print(f"[h15 {name}/{precision}] aggregator tokens: semantic rel {e_st:.2e} acoustic rel {e_at:.2e}")
    assert e_st < TOL and e_at < TOL
    G = align.shape[1]
    for tag, key, qname, tok_key, tok_ref in (("acoustic", "acoustic_codes", "quantizer", "ac_agg.out", "ac_tok"),
                                              ("semantic", "semantic_codes", "semantic_quantizer", "sem_agg.out", "sem_tok")):
        got, want = out[key].cpu(), torch.from_numpy(z[key])
        assert got.shape == want.shape == (B, nq, G) and got.dtype == torch.int64
        gp, gl = oad.extract_lengths(got, K)
        wp, wl = oad.extract_lengths(want, K)
        assert torch.equal(gl, wl), "token lengths packed into the indices differ"
        rows = lambda t: t.double().cpu().transpose(1, 2).reshape(B * G, -1)
        a = audit_codes(gp, wp, rows(taps[tok_key]), rows(torch.from_numpy(z[tok_ref])), _codebooks(sd, qname, nq))
        print(f"[h15 {name}/{precision}] {tag}: {a}")
        assert a["explained"], f"{tag} index differs at a numerically safe decision"
        assert a["index_match_rate"] > 0.97
    # decode the REFERENCE's codes
    dtaps = {}
    rec = m.decode(torch.from_numpy(z["acoustic_codes"]).cuda(), torch.from_numpy(z["semantic_codes"]).cuda(), taps=dtaps)
    torch.cuda.synchronize()
    assert rel(dtaps["dec.z"], torch.from_numpy(z["z"])) < 1e-6
    e_bn, e_wav = rel(dtaps["bottleneck.out"], torch.from_numpy(z["bottleneck"])), rel(rec, torch.from_numpy(z["wav_rec"]))
    print(f"[h15 {name}/{precision}] bottleneck rel {e_bn:.2e} wav rel {e_wav:.2e}")
    assert rec.shape == tuple(z["wav_rec"].shape)
    assert e_bn < TOL and e_wav < TOL


def test_h15_taps_vs_oracle_and_reference_surface(lib):
    """layer-level taps of the aggregator / bottleneck stacks against the oracle; constructor from the reference's config blocks"""
    from oracle import hcodec15 as o15
    from unified_audio_b200.codec_h15 import CodecH15, H15, config_from_kwargs
    z, c, sd, m, wav, feat = _setup("shallow", "mixed")
    otaps, gtaps = {}, {}
    o15.codec_encode(sd, c, wav, feat, otaps)
    m.encode(wav.cuda(), feat.cuda(), taps=gtaps)
    for k in ("sem_agg.interleaved", "sem_agg.layer0", f"sem_agg.layer{c['agg']['layers'] - 1}", "ac_agg.interleaved", "ac_agg.layer0"):
        e = rel(gtaps[k], otaps[k])
        print(f"  tap {k}: {e:.2e}")
        assert e < TOL
    # threshold argument (codec_adaptive.py:153-161): a higher threshold merges less
    out_hi = m.encode(wav.cuda(), feat.cuda(), threshold=0.95)
    oa, _ = o15.codec_encode(sd, c, wav, feat, threshold=0.95)
    assert out_hi["acoustic_codes"].shape == oa.shape
    with pytest.raises(ValueError):
        m.encode(wav.cuda(), feat.cuda(), threshold=1.5)
    y = SHIPPED_CONFIG
    cfg = config_from_kwargs(y["encoder_config"], y["decoder_config"], y["quantizer_config"], y["adaptive_config"])
    assert cfg == {k: v for k, v in H15.items()}
    full = CodecH15(y["encoder_config"], y["decoder_config"], y["quantizer_config"], y["adaptive_config"])
    assert set(full.state_dict()) == set(o15.param_specs(o15.H15))


def test_h15_bench_shape_vs_oracle(lib):
    """the shipped config at the bench leg's clip length (10 s -> 250 frames, T + G ~ 350 rows per aggregator sequence), 2 clips,
    bench-style inputs, against the oracle on the same weights: every float tap < 1e-3, grouping identical, indices audited"""
    from oracle import adaptive as oad
    from oracle import hcodec15 as o15
    from oracle.make_golden_h15 import synth
    from oracle.parity import audit_codes
    from unified_audio_b200.codec_h15 import CodecH15
    c = o15.H15
    sd = o15.make_state_dict(c, 31)
    m = CodecH15(precision="mixed", _cfg={k: v for k, v in c.items() if k != "layer_scale"})
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    B, N = 2, 250
    wav, feat = synth(c, B, N, 32)
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    otaps, gtaps = {}, {}
    oa, osem = o15.codec_encode(sd, c, wav, feat, otaps)
    out = m.encode(wav.cuda(), feat.cuda(), taps=gtaps)
    torch.cuda.synchronize()
    align = otaps["align"]
    assert torch.equal(gtaps["seg"].cpu().long(), align.argmax(1)), "grouping differs from the oracle"
    G = align.shape[1]
    print(f"[h15 bench shape] groups per clip {otaps['n_groups'].tolist()} of {N} frames; sequences of {N + G} rows")
    for k in ("enc.out", "sem.out", "sem_agg.layer0", "sem_agg.layer31", "ac_agg.layer31", "sem_agg.out", "ac_agg.out"):
        e = rel(gtaps[k], otaps[k])
        print(f"  [h15 bench shape] tap {k}: {e:.2e}")
        assert e < TOL
    K, nq = c["codebook_size"], c["nq"]
    rows = lambda t: t.double().cpu().transpose(1, 2).reshape(B * G, -1)
    for tag, got, want, qname, key in (("acoustic", out["acoustic_codes"], oa, "quantizer", "ac_agg.out"),
                                       ("semantic", out["semantic_codes"], osem, "semantic_quantizer", "sem_agg.out")):
        gp, gl = oad.extract_lengths(got.cpu(), K)
        wp, wl = oad.extract_lengths(want, K)
        assert torch.equal(gl, wl)
        a = audit_codes(gp, wp, rows(gtaps[key]), rows(otaps[key]), _codebooks(sd, qname, nq))
        print(f"[h15 bench shape] {tag}: {a}")
        assert a["explained"] and a["index_match_rate"] > 0.97
    dt_o, dt_g = {}, {}
    ref = o15.codec_decode(sd, c, oa, osem, dt_o)
    rec = m.decode(oa.cuda(), osem.cuda(), taps=dt_g)
    torch.cuda.synchronize()
    for k in ("bottleneck.layer0", "bottleneck.out", "dec.tf", "dec.post"):
        e = rel(dt_g[k], dt_o[k])
        print(f"  [h15 bench shape] tap {k}: {e:.2e}")
        assert e < TOL
    e_wav = rel(rec, ref)
    print(f"[h15 bench shape] wav rel {e_wav:.2e}")
    assert rec.shape == ref.shape == (B, N * 640) and e_wav < TOL
