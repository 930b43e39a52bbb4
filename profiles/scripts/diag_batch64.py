"""Diagnostic: per-clip, per-tap error of a 64-clip encode against the oracle (which stage / which clips go wrong)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import hcodec2, weights
from unified_audio_b200.codec import Codec

cfg = weights.H2_FULL
clips = int(os.environ.get("CLIPS", "64"))
sd = weights.make_h2_state_dict(cfg, 0)
m = Codec(cfg["encoder_config"], cfg["decoder_config"], cfg["quantizer_config"], cfg["semantic_encoder_config"],
          cfg["semantic_decoder_config"], precision=os.environ.get("PREC", "mixed"))
m.load_state_dict(sd); m = m.cuda()
wav, feat = weights.synth_inputs(cfg, clips, 125, 2000)
torch.set_num_threads(32)
ot = {}
for i in range(0, clips, 8):
    t = {}
    hcodec2.codec_encode(sd, cfg, wav[i:i + 8], feat[i:i + 8], taps=t)
    for k, v in t.items():
        ot.setdefault(k, []).append(v)
ot = {k: torch.cat(v) for k, v in ot.items()}
gt = {}
m.encode(wav.cuda(), feat.cuda(), taps=gt)
torch.cuda.synchronize()
for k in ot:
    if k not in gt or k == "enc.feat":
        continue
    a, b = gt[k].float().cpu().double(), ot[k].double()
    per = (a - b).abs().flatten(1).amax(1) / b.abs().max()
    bad = (per > 1e-3).nonzero().flatten().tolist()
    print(f"{k:16s} max-rel {float(per.max()):.2e}  clips > 1e-3: {bad[:40]}")
# same clips alone (batch of 1) for the worst clip
k = "enc.out"
per = ((gt[k].float().cpu().double() - ot[k].double()).abs().flatten(1).amax(1) / ot[k].double().abs().max())
w = int(per.argmax())
g1 = {}
m.encode(wav[w:w + 1].cuda(), feat[w:w + 1].cuda(), taps=g1)
torch.cuda.synchronize()
for k in ("enc.embed_norm", "enc.prior", "enc.post", "enc.out"):
    a, b = g1[k].float().cpu().double(), ot[k][w:w + 1].double()
    print(f"clip {w} alone: {k:16s} max-rel {float((a - b).abs().max() / ot[k].double().abs().max()):.2e}")
# where in time is the error for the worst clip at enc.post?
a, b = gt["enc.post"].float().cpu().double()[w], ot["enc.post"].double()[w]
et = (a - b).abs().amax(0) / ot["enc.post"].double().abs().max()
print("enc.post err over time (every 25 frames):", [f"{float(x):.1e}" for x in et[::25]])
a, b = gt["enc.prior"].float().cpu().double()[w], ot["enc.prior"].double()[w]
et = (a - b).abs().amax(0) / ot["enc.prior"].double().abs().max()
print("enc.prior err over time (every 25 frames):", [f"{float(x):.1e}" for x in et[::25]])
