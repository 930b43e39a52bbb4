"""GPU parity of unified_audio_b200.CodecH1 (H-Codec-1.0, BASELINE configs[0]: 1 s 16 kHz mono clip round trip)
against the golden outputs of the reference's own modules (tests/golden/h1_full_1s.npz) and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("precision", ["mixed", "accurate"])
def test_h1_config1_roundtrip(lib, precision):
    from oracle import hcodec1, rvq
    from unified_audio_b200.codec_h1 import CodecH1
    z = np.load(os.path.join(GOLD, "h1_full_1s.npz"))
    meta = json.loads(str(z["meta"]))
    c = hcodec1.H1
    sd = hcodec1.make_state_dict(c, meta["seed_w"])
    m = CodecH1({}, {}, {}, precision=precision)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    g = torch.Generator().manual_seed(meta["seed_x"])
    x = 0.1 * torch.randn(1, 1, 16000, generator=g)
    g2 = torch.Generator().manual_seed(meta["seed_x"] + 1)
    f = torch.randn(1, 768, 50, generator=g2)
    feat = torch.sign(f) * f.abs() ** 0.3
    otaps, gtaps = {}, {}
    hcodec1.seanet_encoder(sd, c, x, otaps)
    ac, sc = m.encode(x.cuda(), feat.cuda(), taps=gtaps)
    torch.cuda.synchronize()
    for k in otaps:
        if k in gtaps:
            print(f"  tap {k}: max-rel {rel(gtaps[k], otaps[k]):.2e}")
    e_emb, e_sem = rel(gtaps["enc.out"], torch.from_numpy(z["emb"])), rel(gtaps["sem.out"], torch.from_numpy(z["sem"]))
    print(f"[h1/{precision}] emb rel {e_emb:.2e} sem rel {e_sem:.2e}")
    assert e_emb < 1e-3 and e_sem < 1e-3
    want_a, want_s = torch.from_numpy(z["acoustic_codes"]), torch.from_numpy(z["semantic_codes"])
    assert ac.shape == want_a.shape == (1, 4, 25)
    emb_ref = torch.from_numpy(z["emb"])
    rows = emb_ref.transpose(1, 2).reshape(25, 512)
    cb = torch.stack([sd[f"quantizer.layers.{i}._codebook.embed"][0] for i in range(4)], 0)
    _, margin = rvq.rvq_margin_audit(rows, cb, want_a.transpose(1, 2).reshape(25, 4))
    bad = (ac.cpu() != want_a).transpose(1, 2).reshape(25, 4)
    print(f"[h1/{precision}] differing acoustic indices {int(bad.sum())}/100, semantic {int((sc.cpu() != want_s).sum())}/100;"
          f" min margin {float(margin.min()):.2e}")
    for t in range(25):
        if bad[t].any():
            q = int(bad[t].nonzero()[0])
            assert float(margin[t, q]) < 1e-3, "index differs at a numerically safe decision"
    rec = m.decode(want_a.cuda(), want_s.cuda())
    torch.cuda.synchronize()
    e_wav = rel(rec, torch.from_numpy(z["wav_rec"]))
    print(f"[h1/{precision}] wav rel {e_wav:.2e}")
    assert rec.shape == (1, 16000) and e_wav < 1e-3
