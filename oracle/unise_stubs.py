"""Deterministic stand-ins for the four components `Model.test_step` glues together (TEST INFRASTRUCTURE).

Used twice with identical arithmetic: by oracle/make_golden_unise.py under the REFERENCE's unmodified `test_step`
(QuarkAudio-UniSE/model/model.py:170-286) and by tests/test_host.py under `unified_audio_b200.unise.Model._enhance`.  Every output
depends on every input the real component reads (features of every segment, the enrollment, the task, the semantic length), so a
difference in padding, segmenting, normalisation, enrollment repetition, call order or trimming changes the waveform.
"""
import torch

TASKS = {"se": 1, "tse": 2, "rtse": 3}


def _features(padded: torch.Tensor) -> torch.Tensor:
    """padded [B, T + 320] -> [B, F, 8]: 25 ms frames every 20 ms through a fixed projection, as three 'hidden states' averaged the way
    model.py:44-45 averages WavLM's (torch.stack(...).mean(1))."""
    P = torch.randn(400, 8, generator=torch.Generator().manual_seed(1234)) / 20.0
    h = padded.unfold(-1, 400, 320) @ P
    return torch.stack((0.5 * h, h, 1.5 * h), dim=1).mean(1)


class _HFOut:
    def __init__(self, hidden_states):
        self.hidden_states = hidden_states


class HFSemanticModel:
    """what the reference calls: self.semantic_model(F.pad(wavs, (160, 160)), output_hidden_states=True).hidden_states"""

    def __call__(self, wavs, output_hidden_states=True):
        h = wavs.unfold(-1, 400, 320) @ (torch.randn(400, 8, generator=torch.Generator().manual_seed(1234)) / 20.0)
        return _HFOut((0.5 * h, h, 1.5 * h))


class SemanticModel:
    """what unise.Model calls (the SSLFrontEnd contract): unpadded wavs in, mean hidden state out"""

    def __call__(self, wavs):
        return _features(torch.nn.functional.pad(wavs, (160, 160))).detach()


class Dnn:
    def __init__(self):
        self.calls = []

    def generate(self, task_name, enroll_mel, enroll_feats, mix_mel, mix_feats, do_sample=True, **kw):
        assert (enroll_mel is None) == (enroll_feats is None) and do_sample is False
        B, T, F = mix_feats.size(0), mix_mel.size(1), mix_feats.size(1)
        assert mix_mel.size(0) == B and (enroll_feats is None or enroll_feats.size(0) == B)
        task = TASKS[task_name]
        key = mix_feats.double().sum((1, 2))
        if enroll_feats is not None:
            key = key + 3.0 * enroll_feats.double().sum((1, 2))
        j = torch.arange(32, dtype=torch.float64)
        gids = (torch.floor(key.abs()[:, None] * 1e3) + 17 * j[None] + 1000 * task).long() % 4096
        t = torch.arange(T)
        per_frame = mix_feats.double().sum(-1)[:, t.clamp(max=F - 1)]
        sids = (torch.floor(per_frame.abs() * 1e4) + 31 * t[None].double() + task + torch.floor(key.abs()[:, None] * 10)).long() % 8192
        self.calls.append([task_name, enroll_feats is not None, int(B), int(T)])
        return gids, sids


class Tokenizer:
    def detokenize(self, global_tokens, semantic_tokens):
        """[B, 1, 32], [B, T] -> [B, 1, T * 320]"""
        B, T = semantic_tokens.shape
        assert global_tokens.shape == (B, 1, 32)
        k = torch.arange(320, dtype=torch.float32)
        phase = 0.37 * semantic_tokens.float()[:, :, None] + 0.01 * k[None, None]
        gain = 0.5 + (global_tokens.sum((1, 2)) % 97).float() / 200.0
        return (torch.sin(phase).reshape(B, 1, T * 320) * gain[:, None, None])
