"""Pin the control flow of UniSE's `Model.test_step` (QuarkAudio-UniSE/model/model.py:170-286) against the REFERENCE'S OWN CODE.

TEST INFRASTRUCTURE.  Run in the build container only:  python -m oracle.make_golden_unise

`model/model.py` cannot be imported as shipped here (pytorch_lightning, soundfile, x_transformers, the Spark-TTS checkpoint and the
WavLM download are all absent), but its `test_step` is plain tensor glue around four components.  This script imports the reference's
module with stub packages for its missing imports, builds a `Model` WITHOUT running its `__init__`, plugs deterministic stand-ins for the
four components (`oracle.unise_stubs`) and runs the reference's unmodified `test_step` for the modes 'se', 'tse' and 'ss'; the
waveforms it hands to `sf.write` are captured and their digests written to tests/golden/unise_glue.npz (the inputs are re-made from the seed).
tests/test_host.py drives `unified_audio_b200.unise.Model._enhance` with the same stand-ins and must reproduce them exactly:
wrap-padding, segmenting, the 'se' normalisation, the enrollment repetition, the se -> tse -> rtse chain of 'ss', the trimming.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/QuarkAudio-UniSE/model"


def import_reference_model():
    written = []
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass
    pl.LightningModule = LightningModule
    sf = types.ModuleType("soundfile")
    sf.write = lambda path, data, samplerate: written.append((str(path), np.array(data, copy=True), int(samplerate)))
    pkg = types.ModuleType("refmodel")
    pkg.__path__ = [REF]
    bic = types.ModuleType("refmodel.bicodec")
    bic.BiCodecTokenizer = object          # only named in __init__, which is not run
    llm = types.ModuleType("refmodel.llm")
    llm.LLM_SFT = object
    for name, mod in (("pytorch_lightning", pl), ("soundfile", sf), ("refmodel", pkg), ("refmodel.bicodec", bic), ("refmodel.llm", llm)):
        sys.modules[name] = mod
    spec = importlib.util.spec_from_file_location("refmodel.model", os.path.join(REF, "model.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refmodel.model"] = mod
    spec.loader.exec_module(mod)
    return mod.Model, written


def make_cases():
    """the inputs are regenerated from the seed by the test (they are not stored)"""
    SEG = 5 * 16000
    g = torch.Generator().manual_seed(20)
    return {
        "se": (None, 0.1 * torch.randn(1, 2 * SEG + 4321, generator=g)),
        "tse": (0.1 * torch.randn(1, 30000, generator=g), 0.1 * torch.randn(1, SEG + 777, generator=g)),
        "ss_long": (None, 0.1 * torch.randn(1, SEG + 12345, generator=g)),
        "ss_short": (None, 0.1 * torch.randn(1, 30011, generator=g)),
    }


def digest(x: np.ndarray) -> np.ndarray:
    """what the fixture keeps of a waveform: length, fp64 sum / sum of squares, every 53rd sample and the last 64 (exact float32 values)"""
    x = np.asarray(x, dtype=np.float32)
    head = np.array([x.size, x.astype(np.float64).sum(), (x.astype(np.float64) ** 2).sum()], dtype=np.float64)
    return np.concatenate([head, x[::53].astype(np.float64), x[-64:].astype(np.float64)])


def main():
    from oracle import unise_stubs as st
    Model, written = import_reference_model()
    m = Model.__new__(Model)
    torch.nn.Module.__init__(m)
    m.config = {"save_enhanced": "/nonexistent"}            # sf.write is the capture stub: nothing touches the disk
    m.stft_conf = dict(hop_length=320, win_length=640, n_fft=640, n_mels=80)
    m.tokenizer = st.Tokenizer()
    m.dnn = st.Dnn()
    m.semantic_model = st.HFSemanticModel()
    cases = make_cases()
    out = {}
    for name, (enroll, src) in cases.items():
        mode = name.split("_")[0]
        del written[:]
        m.dnn.calls = []
        with torch.no_grad():
            m.test_step((mode, enroll, src, None, [16000], None, ["utt"]), 0)
        assert len(written) == (2 if mode == "ss" else 1), written
        for i, (path, data, sr) in enumerate(written):
            assert sr == 16000 and data.shape == (src.size(-1),) and data.dtype == np.float32
            out[f"{name}.est{i}"] = digest(data)
        out[f"{name}.calls"] = np.array(json.dumps(m.dnn.calls))
        print(name, "reference test_step ->", [(os.path.basename(p), d.shape) for p, d, _ in written], "generate calls:", m.dnn.calls)
    # ---- stft_logmel (model.py:53-79), unmodified: dead compute for the LM (only its length is read) but part of the surface
    xm = 0.1 * torch.randn(2, 16000 - 77, generator=torch.Generator().manual_seed(21))
    mel = m.stft_logmel(xm)
    out["mel.y"] = mel.numpy()                       # the input is re-made from its seed (21) by the test
    print("stft_logmel", tuple(mel.shape))
    meta = dict(reference="QuarkAudio-UniSE/model/model.py:170-286 (unmodified test_step, stub components oracle/unise_stubs.py)",
                cases=list(cases))
    path = os.path.join(ROOT, "tests", "golden", "unise_glue.npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
