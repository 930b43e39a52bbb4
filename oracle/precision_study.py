"""Error-budget study behind DESIGN.md's precision policy (TEST INFRASTRUCTURE).

Re-runs the oracle with every dense contraction's operands rounded to a
reduced mantissa (emulating tensor-core input formats, fp32 accumulation) and
reports the end-to-end deviation from the fp32 oracle and from fp64 truth.
Usage: python -m oracle.precision_study [--full] [--tokens N]
"""
import sys, json, time
import torch, torch.nn.functional as F
from oracle import hcodec2, weights

def round_mant(x, bits):
    """round-to-nearest to `bits` explicit mantissa bits (tf32: 10, bf16: 7)."""
    if bits >= 23: return x
    xi = x.contiguous().view(torch.int32)
    drop = 23 - bits
    xi = (xi + (1 << (drop - 1))) & ~((1 << drop) - 1)
    return xi.view(torch.float32)

class Emu:
    def __init__(self, bits, skip_dft=True):
        self.bits = bits
    def __enter__(self):
        self.lin, self.conv, self.mm = F.linear, F.conv1d, torch.matmul
        b = self.bits
        lin, conv, mm = self.lin, self.conv, self.mm
        def linear(x, w, bias=None): return lin(round_mant(x, b), round_mant(w, b), bias)
        def conv1d(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
            if groups != 1: return conv(x, w, bias, stride, padding, dilation, groups)
            return conv(round_mant(x, b), round_mant(w, b), bias, stride, padding, dilation, groups)
        def matmul(a, c): return mm(round_mant(a, b), round_mant(c, b))
        F.linear, F.conv1d, torch.matmul = linear, conv1d, matmul
        return self
    def __exit__(self, *a):
        F.linear, F.conv1d, torch.matmul = self.lin, self.conv, self.mm

def metrics(a, b):
    a, b = a.double(), b.double()
    return dict(max_rel=float((a - b).abs().max() / b.abs().max()), l2_rel=float((a - b).norm() / b.norm()))

def main():
    full = "--full" in sys.argv
    ntok = int(sys.argv[sys.argv.index("--tokens") + 1]) if "--tokens" in sys.argv else 8
    cfg = weights.H2_FULL if full else weights.h2_small(dim=512, inter=1536, enc_layers=6, dec_layers=8, tf_layers=2,
                                                        sem_ch=512, nq=16, cb=1024, qdim=512)
    sd = weights.make_h2_state_dict(cfg, 0)
    wav, feat = weights.synth_inputs(cfg, 1, ntok, 2000)
    t0 = time.time()
    emb32 = hcodec2.encoder_forward(sd, cfg["encoder_config"], wav, aten_lstm=False)
    sem32 = hcodec2.semantic_encoder_forward(sd, cfg["semantic_encoder_config"], feat)
    ac, sc = hcodec2.codec_encode(sd, cfg, wav, feat)
    rec32 = hcodec2.codec_decode(sd, cfg, ac, sc, aten_lstm=False)
    print("fp32 oracle pass %.1fs" % (time.time() - t0), flush=True)
    sd64 = hcodec2.to_dtype(sd, torch.float64)
    emb64 = hcodec2.encoder_forward(sd64, cfg["encoder_config"], wav.double(), aten_lstm=False)
    sem64 = hcodec2.semantic_encoder_forward(sd64, cfg["semantic_encoder_config"], feat.double())
    rec64 = hcodec2.codec_decode(sd64, cfg, ac, sc, aten_lstm=False)
    out = {"fp32_vs_fp64": dict(emb=metrics(emb32, emb64), sem=metrics(sem32, sem64), wav=metrics(rec32, rec64))}
    print(json.dumps(out), flush=True)
    for name, bits in (("tf32", 10), ("bf16", 7), ("m16", 16)):
        with Emu(bits):
            emb = hcodec2.encoder_forward(sd, cfg["encoder_config"], wav, aten_lstm=False)
            sem = hcodec2.semantic_encoder_forward(sd, cfg["semantic_encoder_config"], feat)
            rec = hcodec2.codec_decode(sd, cfg, ac, sc, aten_lstm=False)
        r = {name: dict(emb=metrics(emb, emb32), sem=metrics(sem, sem32), wav=metrics(rec, rec32),
                        emb_vs64=metrics(emb, emb64), wav_vs64=metrics(rec, rec64))}
        print(json.dumps(r), flush=True)
        out.update(r)
    # how many RVQ decisions would flip with the tf32 embedding?
    return out

if __name__ == "__main__":
    main()
