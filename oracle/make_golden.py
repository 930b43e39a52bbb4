"""Pin the oracle against the reference's own modules and write tests/golden/.

TEST INFRASTRUCTURE - runs ONLY in the build container (needs /root/reference,
which does not exist on the GPU box).  What it does:

 1. registers oracle.rvq.ResidualVQ under the module name
    `vector_quantize_pytorch` (the one un-vendored dependency), then imports the
    reference's `vq.Codec` from /root/reference/QuarkAudio-HCodec/HCodec-2.0;
 2. checks the reference's state_dict keys/shapes == oracle/weights.py specs;
 3. loads the seeded state-dict into the reference modules and runs
    Codec.encode / Codec.decode (vq/codec.py:75-99) on seeded inputs;
 4. asserts the oracle restatement reproduces the reference (encoder/semantic
    encoder/decoder float outputs and the codes), and
 5. stores the REFERENCE outputs as golden fixtures (npz) together with the
    seeds, so tests can replay them anywhere.

Usage:  python -m oracle.make_golden
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/QuarkAudio-HCodec/HCodec-2.0"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference_codec():
    from oracle import rvq
    m = types.ModuleType("vector_quantize_pytorch")
    m.ResidualVQ = rvq.ResidualVQ
    sys.modules["vector_quantize_pytorch"] = m
    sys.path.insert(0, REF)
    from vq import Codec  # noqa: E402  (reference module)
    return Codec


def build_reference(cfg, sd):
    Codec = import_reference_codec()
    ref = Codec(cfg["encoder_config"], cfg["decoder_config"], cfg["quantizer_config"],
                cfg["semantic_encoder_config"], cfg["semantic_decoder_config"]).eval()
    ref_sd = ref.state_dict()
    ours = {k: tuple(v.shape) for k, v in sd.items()}
    theirs = {k: tuple(v.shape) for k, v in ref_sd.items() if not k.startswith("semantic_decoder.")}
    assert ours == theirs, (set(ours) ^ set(theirs), [k for k in ours if k in theirs and ours[k] != theirs[k]])
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("semantic_decoder.") for k in missing)
    return ref


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def run(name, cfg, batch, n_tokens, seed_w, seed_x):
    from oracle import hcodec2, weights
    sd = weights.make_h2_state_dict(cfg, seed_w)
    ref = build_reference(cfg, sd)
    wav, feat = weights.synth_inputs(cfg, batch, n_tokens, seed_x)
    with torch.no_grad():
        emb_ref = ref.encoder(wav)
        sem_ref = ref.semantic_encoder(feat)
        ac_ref, sc_ref = ref.encode(wav, feat)
        rec_ref = ref.decode(ac_ref, sc_ref)
    taps = {}
    emb = hcodec2.encoder_forward(sd, cfg["encoder_config"], wav, taps)
    sem = hcodec2.semantic_encoder_forward(sd, cfg["semantic_encoder_config"], feat)
    ac, sc = hcodec2.codec_encode(sd, cfg, wav, feat)
    rec = hcodec2.codec_decode(sd, cfg, ac_ref, sc_ref)
    # explicit-recurrence LSTM variant (the restated arithmetic) vs ATen's fused LSTM
    emb_loop = hcodec2.encoder_forward(sd, cfg["encoder_config"], wav, aten_lstm=False)
    report = dict(
        name=name, batch=batch, n_tokens=n_tokens, seed_w=seed_w, seed_x=seed_x,
        enc_rel=rel(emb, emb_ref), sem_rel=rel(sem, sem_ref), rec_rel=rel(rec, rec_ref),
        enc_loop_lstm_rel=rel(emb_loop, emb_ref),
        codes_equal=bool((ac == ac_ref).all() and (sc == sc_ref).all()),
        wav_len=int(rec_ref.shape[-1]),
    )
    print(json.dumps(report))
    assert report["enc_rel"] < 1e-5 and report["sem_rel"] < 1e-5 and report["rec_rel"] < 1e-5, report
    assert report["codes_equal"], report
    assert rec_ref.shape[-1] == n_tokens * 3840
    np.savez_compressed(
        os.path.join(GOLD, f"h2_{name}.npz"),
        emb=emb_ref.numpy(), sem=sem_ref.numpy(), acoustic_codes=ac_ref.numpy(), semantic_codes=sc_ref.numpy(),
        wav_rec=rec_ref.numpy(), meta=np.array(json.dumps(dict(report, cfg=cfg))),
    )
    with open(os.path.join(GOLD, f"h2_keys_{name}.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in ref.state_dict().items()}, f, indent=0)
    return report


def main():
    from oracle import weights
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    reports = [
        run("small", weights.h2_small(), batch=2, n_tokens=6, seed_w=11, seed_x=2000),
        run("mid", weights.h2_small(dim=512, inter=1536, enc_layers=3, dec_layers=4, tf_layers=2, sem_ch=512,
                                    nq=16, cb=1024, qdim=512), batch=1, n_tokens=4, seed_w=12, seed_x=2100),
    ]
    if "--full" in sys.argv:
        reports.append(run("full", weights.H2_FULL, batch=1, n_tokens=4, seed_w=0, seed_x=2000))
    with open(os.path.join(GOLD, "h2_pinning_report.json"), "w") as f:
        json.dump(reports, f, indent=1)


if __name__ == "__main__":
    main()
