"""UniSE inference surface (`Model.test_step`) on the device: the caller of the AR-LM hot path.

Mirrors QuarkAudio-UniSE/model/model.py:20-286 (a LightningModule in the reference; a plain nn.Module here - the Lightning task,
training / validation steps and checkpoint callbacks are out of scope):
    Model(config, tokenizer=BiCodecTokenizer(BiCodec), dnn=LLM_SFT, semantic_model=SSLFrontEnd(WAVLM_BASE_PLUS))
    .extract_semantic_features(wavs [B, T] @ 16 kHz) -> [B, T/320, 768]        model.py:37-51
    .stft_logmel(x [B, T]) -> [B, ceil(T/320), 80]                               model.py:53-79
    .test_step((mode, enroll, src, tgt, fs, lengths, names), batch_idx)          model.py:170-286, modes 'se' / 'tse' / 'ss'
and audio_tokenizer.py:30-125 for `BiCodecTokenizer.detokenize(global_tokens, semantic_tokens)`.

Everything between the waveform in and the waveform out stays on the GPU: wrap-pad + 5 s segmenting (`qb_pad_wav`, no NumPy round
trip), WavLM features (csrc/ssl.cu + the conv-GEMM / attention kernels), `LLM_SFT.generate` (csrc/llm.cu), `BiCodec.detokenize`.
`stft_logmel` is dead compute on this path - `generate` reads only `mix_mel.size(1)` (llm_sft.py:166) - so `test_step` hands the LM a
shape-only tensor (`mel_like`); `stft_logmel` itself is provided for callers that want the values (torch.stft: plumbing, not a kernel
of this library).  No CPU fallback: the three sub-modules refuse to run off the GPU.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
from torch import nn

from .ssl import wrap_segments

STFT_CONFIG = dict(hop_length=320, win_length=640, n_fft=640, n_mels=80)      # U/conf/config.yaml:124-128
SEG_LEN = 5 * 16000                                                            # model.py:175


class BiCodecTokenizer(nn.Module):
    """audio_tokenizer.py:30-125, detokenize side.  `tokenize` (wav2vec2-large-xlsr-53 features + the BiCodec encoder) is only
    called by the training / validation steps (model.py:96-99,139-142): out of scope, raises."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def tokenize(self, wav):
        raise NotImplementedError("BiCodecTokenizer.tokenize is used by the training / validation steps only (model.py:96-99); "
                                  "the inference path needs detokenize")

    @torch.no_grad()
    def detokenize(self, global_tokens: torch.Tensor, semantic_tokens: torch.Tensor) -> torch.Tensor:
        """global_tokens [B, 1, 32], semantic_tokens [B, T] -> wav [B, 1, T * 320]   (audio_tokenizer.py:108-125)"""
        return self.model.detokenize(semantic_tokens, global_tokens)


class Model(nn.Module):
    def __init__(self, config: Optional[dict] = None, *, tokenizer: BiCodecTokenizer, dnn, semantic_model):
        super().__init__()
        self.config = dict(config or {})
        self.stft_conf = dict(self.config.get("stft_config", STFT_CONFIG))
        self.tokenizer, self.dnn, self.semantic_model = tokenizer, dnn, semantic_model

    # ------------------------------------------------------------------ state (model.py:81-91): tokenizer / semantic_model excluded
    def state_dict(self, *args, **kwargs):
        state = super().state_dict(*args, **kwargs)
        for key in list(state.keys()):
            if key.startswith(("tokenizer.", "semantic_model.")):
                del state[key]
        return state

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """A Lightning checkpoint's `state_dict` holds the LM under `dnn.` (model.py:28)."""
        sd = {k[4:]: v for k, v in state_dict.items() if k.startswith("dnn.")}
        return self.dnn.load_state_dict(sd, strict=False)

    # ------------------------------------------------------------------ features
    @torch.no_grad()
    def extract_semantic_features(self, wavs: torch.Tensor) -> torch.Tensor:
        """model.py:37-51: pad 160 / 160, WavLM-base-plus, mean of the 13 hidden states (no compression)."""
        return self.semantic_model(wavs)

    def mel_frames(self, n_samples: int) -> int:
        return math.ceil(n_samples / self.stft_conf["hop_length"])

    def mel_like(self, x: torch.Tensor) -> torch.Tensor:
        """A tensor with stft_logmel's shape and no arithmetic behind it: the LM reads `mix_mel.size(1)` only (llm_sft.py:166)."""
        return torch.zeros(1, device=x.device).expand(x.shape[0], self.mel_frames(x.shape[-1]), self.stft_conf["n_mels"])

    @torch.no_grad()
    def stft_logmel(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:53-79 verbatim in meaning (torch.stft + HTK mel filter bank + log); not on the generate path (see module doc)."""
        from torchaudio.functional import melscale_fbanks
        if x.ndim != 2:
            raise AssertionError("x: (B, T)")
        hop, win, n_fft, n_mels = (self.stft_conf[k] for k in ("hop_length", "win_length", "n_fft", "n_mels"))
        pad_length = math.ceil(x.size(-1) / hop) * hop - x.size(-1)
        x = torch.nn.functional.pad(x, ((win - hop) // 2, pad_length + (win - hop) // 2))
        spec = torch.stft(x, n_fft, hop, win_length=win, window=torch.hann_window(win).to(x.device), onesided=True, center=False,
                          return_complex=True).transpose(1, 2)
        if not hasattr(self, "fb"):
            self.fb = melscale_fbanks(n_freqs=n_fft // 2 + 1, f_min=0.0, f_max=8000.0, n_mels=n_mels, sample_rate=16000).to(x.device)
        return torch.log(spec.abs() @ self.fb + 1e-10)

    def forward(self, batch):
        """model.py:93-94: the reference's forward is empty; inference goes through test_step."""
        return None

    # ------------------------------------------------------------------ inference (model.py:170-286)
    def _segments(self, src: torch.Tensor) -> torch.Tensor:
        return wrap_segments(src.float().contiguous(), SEG_LEN)          # np.pad(..., 'wrap') + reshape(-1, seg_len) on the device

    def _generate(self, task, enroll_feats, seg_src, do_sample, **gen_kw):
        mix_mel = self.mel_like(seg_src)
        mix_feats = self.extract_semantic_features(seg_src)
        enroll_mel = None
        if enroll_feats is not None:                                    # torch.cat([enroll] * n_segments) (model.py:207-208)
            n = seg_src.size(0)
            enroll_feats = torch.cat([enroll_feats for _ in range(n)], 0)
            enroll_mel = mix_mel            # placeholder: only `is None` is tested for the enrollment mel (llm_sft.py:110-121)
        gids, sids = self.dnn.generate(task_name=task, enroll_mel=enroll_mel, enroll_feats=enroll_feats, mix_mel=mix_mel,
                                       mix_feats=mix_feats, do_sample=do_sample, **gen_kw)
        return gids, sids

    def _detok(self, gids, sids, n_samples):
        est = self.tokenizer.detokenize(gids.unsqueeze(1), sids).squeeze(1)          # (B, t)
        return est.reshape(-1)[:n_samples]

    @torch.no_grad()
    def enhance(self, mode: str, enroll: Optional[torch.Tensor], src: torch.Tensor, do_sample: bool = False, return_ids: bool = False,
                **gen_kw):
        """The body of test_step with tensors in and a device tensor out.  src [1, T] (the reference's test loader yields one
        utterance per batch; like the reference, a batch of several is folded into the segment axis), enroll [1, Te] for 'tse'.
        'ss' returns (s1, s2)."""
        if src.device.type != "cuda":
            raise RuntimeError("unified_audio_b200.unise.Model runs on CUDA only (no CPU fallback)")
        return self._enhance(mode, enroll, src, do_sample, return_ids, **gen_kw)

    def _enhance(self, mode, enroll, src, do_sample=False, return_ids=False, **gen_kw):
        """test_step's control flow (model.py:174-286) over the four components; pinned against the reference's own `test_step`
        driven with stub components (oracle/make_golden_unise.py -> tests/golden/unise_glue.npz, tests/test_host.py)."""
        n_samples = src.size(-1)
        if mode == "se":                                                 # model.py:174-193
            seg = self._segments(src)
            seg = seg / src.abs().max(dim=-1, keepdim=True)[0]
            gids, sids = self._generate("se", None, seg, do_sample, **gen_kw)
            est = self._detok(gids, sids, n_samples)
            return (est, gids, sids) if return_ids else est
        if mode == "tse":                                                # model.py:197-224
            seg = self._segments(src)
            enroll_feats = self.extract_semantic_features(enroll)
            gids, sids = self._generate("tse", enroll_feats, seg, do_sample, **gen_kw)
            est = self._detok(gids, sids, n_samples)
            return (est, gids, sids) if return_ids else est
        if mode == "ss":                                                 # model.py:225-286: se on the first 5 s, then tse, then rtse
            first = src[:, :SEG_LEN] if n_samples > SEG_LEN else self._segments(src)[:src.size(0)]
            gids, sids = self._generate("se", None, first, do_sample, **gen_kw)
            enr = self.tokenizer.detokenize(gids.unsqueeze(1), sids).squeeze(1)[:, :SEG_LEN]
            enr = enr / (torch.max(torch.abs(enr)) + 1e-5) * 0.99
            enroll_feats = self.extract_semantic_features(enr)
            seg = self._segments(src)
            g1, s1 = self._generate("tse", enroll_feats, seg, do_sample, **gen_kw)
            est1 = self._detok(g1, s1, n_samples)
            g2, s2 = self._generate("rtse", enroll_feats, seg, do_sample, **gen_kw)
            est2 = self._detok(g2, s2, n_samples)
            return est1, est2
        raise ValueError(f"unknown mode {mode!r} (the reference's test_step handles 'se', 'tse', 'ss')")

    def test_step(self, batch, batch_idx=0):
        """model.py:170-286: batch = (mode, enroll, src, tgt, fs, lengths, names); greedy decoding (do_sample = False, model.py:173).
        Returns the enhanced waveform(s) as NumPy (the reference's last step before its optional sf.write) and writes
        `<save_enhanced>/<name>.wav` (`_s1` / `_s2` for 'ss') when the config asks for it."""
        mode, enroll, src, tgt, fs, lengths, names = batch
        out = self.enhance(mode, enroll, src, do_sample=False)
        outs = out if isinstance(out, tuple) else (out,)
        arrays = [o.cpu().numpy() for o in outs]
        save_dir = self.config.get("save_enhanced")
        if save_dir is not None:
            suffixes = ("_s1", "_s2") if mode == "ss" else ("",)
            for a, sfx in zip(arrays, suffixes):
                _write_wav(os.path.join(str(save_dir), f"{names[0]}{sfx}.wav"), a, int(fs[0]))
        return arrays if mode == "ss" else arrays[0]


def _write_wav(path: str, data, rate: int) -> None:
    try:
        import soundfile as sf           # the reference's writer (model.py:196); not in every image
        sf.write(path, data, samplerate=rate)
    except ImportError:
        from scipy.io import wavfile
        wavfile.write(path, rate, data)
