"""CPU oracle for the QuarkAudio audio-token hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a CPU restatement (PyTorch CPU
ops, fp32 with an fp64 "truth" mode) of the reference's algorithm for the
path  H-Codec encoder -> ResidualVQ -> AR-LM forward -> H-Codec decoder.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import it, and only as the checker or the timed CPU
baseline - never as part of the shipped product path (`unified_audio_b200`
does not import it and fails loudly when its CUDA library is missing).

Pinning status (see DESIGN.md "Oracle"):
  * H-Codec-2.0 encoder / semantic encoder / decoder: pinned against the
    reference's own modules imported from /root/reference in the build
    container (oracle/make_golden.py; fixtures under tests/golden/).
  * ResidualVQ: third-party `vector-quantize-pytorch==1.22.15` is not
    vendored in the reference and not installable offline -> restated from
    its published algorithm and the reference's in-repo arithmetic template
    (HCodec-2.0/vq/core_vq.py:223-238, 394-412).  PARITY UNPINNED for this
    one dependency: no golden vector of the real library exists here.
  * AR-LM: the reference's CustomLlamaModel cannot be constructed under the
    installed transformers (llm.py:79); restated over HF LlamaModel layers
    semantics and pinned against `transformers.LlamaModel` run here.
"""
