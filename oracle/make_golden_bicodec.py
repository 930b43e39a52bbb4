"""Pin oracle/bicodec.py against the reference's own BiCodec classes and write the golden fixture.

Run in the build container only (reads /root/reference):  python -m oracle.make_golden_bicodec
Builds QuarkAudio-UniSE/model/bicodec's FactorizedVectorQuantize, SpeakerEncoder, Decoder (prenet) and WaveGenerator
with oracle.bicodec's hyper-parameters, loads the seeded weights into them and runs the reference's
`BiCodec.detokenize` (bicodec.py:182-199).  Absent third-party packages are supplied minimally:
  omegaconf  - only `DictConfig` is imported by bicodec.py (type annotation)
  einx       - residual_fsq.py uses one call, get_at("q [c] d, b n q -> q b n d", codebooks, indices): a gather
Outputs: tests/golden/bicodec_small.npz (tokens, wav, intermediate taps), tests/golden/bicodec_keys.json
(reference state-dict keys + shapes of the detokenize path), tests/golden/bicodec_pinning_report.json.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/QuarkAudio-UniSE"


def _stub_modules():
    om = types.ModuleType("omegaconf")
    om.DictConfig = dict
    om.OmegaConf = type("OmegaConf", (), {})
    sys.modules.setdefault("omegaconf", om)
    ex = types.ModuleType("einx")

    def get_at(pattern, codebooks, indices):
        assert pattern == "q [c] d, b n q -> q b n d"
        q = codebooks.shape[0]
        return torch.stack([codebooks[i][indices[..., i]] for i in range(q)], 0)
    ex.get_at = get_at
    sys.modules.setdefault("einx", ex)
    for name, path in (("model", REF + "/model"), ("model.bicodec", REF + "/model/bicodec")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m


def build_reference(cfg, sd):
    _stub_modules()
    from model.bicodec.bicodec import BiCodec
    from model.bicodec.modules.encoder_decoder.feat_decoder import Decoder
    from model.bicodec.modules.encoder_decoder.wave_generator import WaveGenerator
    from model.bicodec.modules.speaker.speaker_encoder import SpeakerEncoder
    from model.bicodec.modules.vq.factorized_vector_quantize import FactorizedVectorQuantize
    q, s = cfg["quantizer"], cfg["speaker"]
    quantizer = FactorizedVectorQuantize(q["input_dim"], q["codebook_size"], q["codebook_dim"], commitment=0.25)
    speaker = SpeakerEncoder(input_dim=128, out_dim=s["out_dim"], latent_dim=s["latent_dim"], token_num=s["token_num"],
                             fsq_levels=s["fsq_levels"], fsq_num_quantizers=s["fsq_num_quantizers"])
    prenet = Decoder(**cfg["prenet"])
    decoder = WaveGenerator(**cfg["decoder"])
    model = BiCodec.__new__(BiCodec)          # BiCodec.__init__ also builds a torchaudio mel front end (tokenize side)
    torch.nn.Module.__init__(model)
    model.quantizer, model.speaker_encoder, model.prenet, model.decoder = quantizer, speaker, prenet, decoder
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # everything the detokenize path touches must have come from the oracle's weights
    untouched = ("quantizer.in_project", "quantizer.cluster_size", "speaker_encoder.speaker_encoder.", "speaker_encoder.perceiver_sampler.",
                 "speaker_encoder.quantizer.project_in")
    bad = [k for k in missing if not k.startswith(untouched)]
    assert not bad, bad
    return model.eval()


def main():
    from oracle import bicodec as ob
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    report = {}
    keys = None
    for name, cfg, B, T, seed in (("small", ob.bicodec_small(), 3, 12, 11), ("full", ob.BICODEC_FULL, 1, 6, 12)):
        sd = ob.make_state_dict(cfg, seed)
        ref = build_reference(cfg, sd)
        sem, glob = ob.synth_tokens(cfg, B, T, seed + 100)
        with torch.no_grad():
            want = ref.detokenize(sem, glob)
        taps = {}
        got = ob.detokenize(sd, cfg, sem, glob, taps)
        err = float((got - want).abs().max() / want.abs().max())
        report[name] = dict(rel_err=err, wav_absmax=float(want.abs().max()), wav_rms=float(want.pow(2).mean().sqrt()),
                            shape=list(want.shape))
        print(name, report[name])
        assert err < 1e-5 and want.shape == (B, 1, T * cfg["hop"])
        path_keys = {k: list(v.shape) for k, v in ref.state_dict().items() if k in sd}
        assert set(path_keys) == set(sd), set(sd) ^ set(path_keys)
        if name == "full":
            keys = path_keys
        if name == "small":
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bicodec_small.npz"),
                                meta=json.dumps(dict(seed=seed, token_seed=seed + 100, B=B, T=T)),
                                semantic=sem.numpy(), global_tokens=glob.numpy(), wav=want.numpy(),
                                z_q=taps["z_q"].numpy(), d_vector=taps["d_vector"].numpy(),
                                prenet_out=taps["prenet.out"].numpy(), stage0=taps["dec.stage0"].numpy())
    json.dump(keys, open(os.path.join(ROOT, "tests", "golden", "bicodec_keys.json"), "w"), indent=0)
    json.dump(report, open(os.path.join(ROOT, "tests", "golden", "bicodec_pinning_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
