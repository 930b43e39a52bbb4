"""Parity AT THE BENCHMARKED SHAPES (VERDICT r01 "what's weak" 1-5): the shipped H-Codec-2.0 configuration on 10 s clips
(T = 500 STFT frames, 125 tokens / stream) against the CPU oracle, both precision policies and both weight
initialisations (oracle.weights and bench.py::random_init_); end-to-end RVQ index identity asserted with the
Lipschitz audit of oracle/parity.py; the tcgen05 LSTM against torch's own fp64 / fp32 nn.LSTM at H = 1536, T = 500,
B = 64; the UniSE LM at the SR (prefix 252, B = 32) and TSE (prefix 503, B = 16) shapes over all 283 cached steps.

Sizes: QB_PARITY_CLIPS (default 64 = one full bench batch = 8000 tokens / stream) sets the audit batch."""
import math
import os

import pytest
import torch

from oracle.parity import audit_codes, feat_tap_error, l2, phase_wrap_clips, rel

pytestmark = pytest.mark.gpu
TOL = 1e-3          # north_star: floats within 1e-3 relative


def _build(cfg, precision, init):
    from oracle import weights
    from unified_audio_b200.codec import Codec
    m = Codec(cfg["encoder_config"], cfg["decoder_config"], cfg["quantizer_config"], cfg["semantic_encoder_config"],
              cfg["semantic_decoder_config"], precision=precision)
    if init == "oracle":
        sd = weights.make_h2_state_dict(cfg, 0)
        m.load_state_dict(sd, strict=True)
        m = m.cuda()
    else:                       # the weights bench.py times (generated on the device)
        import bench
        m = m.cuda()
        bench.random_init_(m, 1234)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    return m, sd


@pytest.mark.parametrize("precision,init", [("mixed", "oracle"), ("accurate", "oracle"), ("mixed", "bench")])
def test_h2_bench_shape_vs_oracle(lib, precision, init):
    """2 clips x 125 tokens (10 s @ 48 kHz), every tap of encoder / semantic encoder / decoder < 1e-3 vs the oracle."""
    from oracle import hcodec2, weights
    cfg = weights.H2_FULL
    model, sd = _build(cfg, precision, init)
    wav, feat = weights.synth_inputs(cfg, 2, 125, 2000)
    otaps, gtaps = {}, {}
    oa, os_ = hcodec2.codec_encode(sd, cfg, wav, feat, taps=otaps)
    ac, sc = model.encode(wav.cuda(), feat.cuda(), taps=gtaps)
    torch.cuda.synchronize()
    worst = 0.0
    for k in otaps:
        if k not in gtaps:
            continue
        a, b = gtaps[k].float().cpu(), otaps[k]
        if k == "enc.feat":
            e = feat_tap_error(a, b)
        else:
            e = rel(a, b)
        print(f"  [{precision}/{init}] tap {k}: max-rel {e:.2e}  l2-rel {l2(a, b):.2e}")
        assert e < TOL, f"tap {k} off by {e:.2e}"
        worst = max(worst, e)
    B, D, N = otaps["enc.out"].shape
    assert N == 125 and tuple(ac.shape) == (2, 16, 125)
    rows = lambda t: t.float().cpu().transpose(1, 2).reshape(B * N, D)
    for tag, got, want, key, q in (("acoustic", ac, oa, "enc.out", "quantizer"), ("semantic", sc, os_, "sem.out", "semantic_quantizer")):
        a = audit_codes(got, want, rows(gtaps[key]), rows(otaps[key]), hcodec2._codebooks(sd, q))
        print(f"  [{precision}/{init}] {tag} codes: {a}")
        assert a["explained"], f"{tag}: an index differs beyond the reach of the embedding tolerance: {a}"
    dtaps, odtaps = {}, {}
    rec = model.decode(oa.cuda(), os_.cuda(), taps=dtaps)
    torch.cuda.synchronize()
    ref = hcodec2.codec_decode(sd, cfg, oa, os_, taps=odtaps)
    for k in odtaps:
        if k in dtaps:
            e = rel(dtaps[k].float(), odtaps[k])
            print(f"  [{precision}/{init}] tap {k}: max-rel {e:.2e}")
            assert e < TOL, f"tap {k} off by {e:.2e}"
    e = rel(rec, ref)
    print(f"[{precision}/{init}] 10 s clips: worst encoder tap {worst:.2e}, wav rel {e:.2e} l2 {l2(rec, ref):.2e}")
    assert rec.shape == ref.shape == (2, 480000) and e < TOL


def test_h2_index_identity_full_batch(lib):
    """The north-star's hard gate at scale: one full bench batch (64 clips x 125 tokens x 2 streams x 16 layers = 256 k
    decisions, `mixed` policy).  Every differing index must be explained by the embedding tolerance (oracle/parity.py);
    the match rate is printed and recorded."""
    from oracle import hcodec2, weights
    cfg = weights.H2_FULL
    clips = int(os.environ.get("QB_PARITY_CLIPS", "64"))
    model, sd = _build(cfg, "mixed", "oracle")
    wav, feat = weights.synth_inputs(cfg, clips, 125, 2000)
    try:
        torch.set_num_threads(min(32, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    gtaps = {}
    oa, os_ = [], []
    chunk = 8
    oemb, osem, ofeat = [], [], []
    for i in range(0, clips, chunk):              # the oracle in chunks (bounded host memory)
        t = {}
        a, s = hcodec2.codec_encode(sd, cfg, wav[i:i + chunk], feat[i:i + chunk], taps=t)
        oa.append(a); os_.append(s); oemb.append(t["enc.out"]); osem.append(t["sem.out"]); ofeat.append(t["enc.feat"])
    oa, os_, oemb, osem, ofeat = torch.cat(oa), torch.cat(os_), torch.cat(oemb), torch.cat(osem), torch.cat(ofeat)
    ac, sc = model.encode(wav.cuda(), feat.cuda(), taps=gtaps)
    torch.cuda.synchronize()
    B, D, N = oemb.shape
    # clips with a phase bin on the angle() branch cut (oracle/parity.py::phase_wrap_clips): checked on the circle, reported, and
    # excluded from the float-tolerance assertions below
    cut, n_wrapped, worst_circ = phase_wrap_clips(gtaps["enc.feat"].float().cpu(), ofeat)
    ok = ~cut
    print(f"[index identity, {clips} clips] {int(cut.sum())} clip(s) have a phase bin on the branch cut ({n_wrapped} of "
          f"{ofeat.shape[0] * (ofeat.shape[1] // 2) * ofeat.shape[2]} bins; on-circle difference there {worst_circ:.1e}): {cut.nonzero().flatten().tolist()}")
    assert worst_circ < 1e-3 and int(cut.sum()) <= max(1, clips // 3)
    assert feat_tap_error(gtaps["enc.feat"].float().cpu()[ok], ofeat[ok]) < TOL
    rows = lambda t: t.float().cpu()[ok].transpose(1, 2).reshape(-1, D)
    e_emb, e_sem = rel(gtaps["enc.out"].float().cpu()[ok], oemb[ok]), rel(gtaps["sem.out"], osem)
    e_cut = rel(gtaps["enc.out"].float().cpu()[cut], oemb[cut]) if bool(cut.any()) else 0.0
    print(f"[index identity, {clips} clips] emb rel {e_emb:.2e} (branch-cut clips: {e_cut:.2e}) sem rel {e_sem:.2e}")
    assert e_emb < TOL and e_sem < TOL
    for tag, got, want, g, o, q in (("acoustic", ac, oa, gtaps["enc.out"], oemb, "quantizer"),
                                    ("semantic", sc, os_, gtaps["sem.out"], osem, "semantic_quantizer")):
        cb = hcodec2._codebooks(sd, q)
        a = audit_codes(got[ok.to(got.device)], want[ok], rows(g), rows(o), cb)
        print(f"[index identity] {tag}: {a}")
        assert a["tokens"] == int(ok.sum()) * 125 and a["explained"], f"{tag}: unexplained index difference {a}"
        # RVQ kernel alone on the ORACLE's embedding: bit-exact on every one of the clips*125*16 decisions (all clips)
        allrows = o.float().transpose(1, 2).reshape(B * N, D)
        qz = model.engine().rvq(0 if q == "quantizer" else 1) if model._use_engine() else (model.quantizer if q == "quantizer" else model.semantic_quantizer)
        idx, _ = qz.encode_rows(allrows.cuda())
        same = torch.equal(idx.cpu().reshape(B, N, -1).transpose(1, 2), want)
        if not same:
            from oracle import rvq
            wr = want.transpose(1, 2).reshape(B * N, -1)
            _, margin = rvq.rvq_margin_audit(allrows, cb, wr)
            diff = (idx.cpu() != wr)
            # only a token's FIRST differing layer is a decision on identical inputs (later layers see another residual)
            first = torch.where(diff.any(1), diff.float().argmax(1), torch.full((B * N,), -1))
            toks = (first >= 0).nonzero().flatten()
            fm = margin[toks, first[toks]]
            print(f"   rvq-on-oracle-embedding: {len(toks)} of {B * N} tokens differ; fp64 relative margins of the oracle's (fp32) decision "
                  f"at the first differing layer: {fm.tolist()} - the kernel returns the exact-arithmetic arg-min there")
            assert float(fm.max()) < 1e-6, "RVQ kernel differs from the oracle on identical inputs at a safe margin"
        else:
            print(f"   RVQ kernel on the oracle's embedding: all {B * N * idx.shape[1]} indices identical")
    # decode of the oracle's codes at the full batch: compare 4 clips' waveforms with the oracle
    rec = model.decode(oa.cuda(), os_.cuda())
    torch.cuda.synchronize()
    ref = hcodec2.codec_decode(sd, cfg, oa[:4], os_[:4])
    e = rel(rec[:4], ref)
    print(f"[index identity] decode of the oracle's codes (B={clips}): wav rel {e:.2e}")
    assert e < TOL


def test_lstm_vs_torch_lstm_bench_shape(lib):
    """tcgen05 LSTM (fp16 W_hh / h operands, fp32 accumulate and cell state) against torch.nn.LSTM in fp64 (truth) and in
    fp32 (what the reference runs, encoder_modules/transformer.py:115,133) at the benchmarked shape H=1536, T=500, B=64."""
    from unified_audio_b200 import ops
    B, T, H = 64, 500, 1536
    torch.manual_seed(5)
    ref64 = torch.nn.LSTM(H, H, 1, batch_first=True).double().cuda()
    x = torch.randn(B, T, H, generator=torch.Generator().manual_seed(6)).cuda()
    x = torch.nn.functional.rms_norm(x, (H,))                       # the layer's input is an RMSNorm output
    with torch.no_grad():
        want = ref64(x.double())[0]
        ref32 = torch.nn.LSTM(H, H, 1, batch_first=True).cuda()
        ref32.load_state_dict({k: v.float() for k, v in ref64.state_dict().items()})
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        got32 = ref32(x)[0]
        torch.backends.cudnn.allow_tf32 = old
        xp = (x.double() @ ref64.weight_ih_l0.t() + ref64.bias_ih_l0 + ref64.bias_hh_l0).float().contiguous()
    U = ops.lstm_tc_units(H)
    out = ops.Planes.zeros((B, T, H), True, "cuda")
    ws = torch.zeros(ops.lstm_tc_workspace_bytes(B, H), dtype=torch.uint8, device="cuda")
    ops.lstm_tc(xp, ops.lstm_tc_permute(ref64.weight_hh_l0.detach().float(), U), U, B, T, H, out, ws)
    torch.cuda.synchronize()
    got = out.hi.double() + out.lo.double()
    e_k, e_32 = rel(got, want), rel(got32, want)
    per_t = (got - want).abs().amax((0, 2)) / want.abs().max()
    print(f"lstm_tc vs nn.LSTM fp64 @ B{B} T{T} H{H}: max-rel {e_k:.2e} l2 {l2(got, want):.2e} "
          f"(t<50: {float(per_t[:50].max()):.2e}, t>=450: {float(per_t[450:].max()):.2e});  nn.LSTM fp32 vs fp64: {e_32:.2e}")
    # the recurrence is contractive: the error does not grow with T
    assert float(per_t[450:].max()) < 2.0 * float(per_t[:100].max()) + 1e-4
    assert e_k < TOL


def _lm(seed=7, gain=2.0):
    from oracle import llama
    from unified_audio_b200.llm import LLM_SFT
    cfg = llama.LM_FULL
    sd = llama.make_lm_state_dict(cfg, seed, gain)
    m = LLM_SFT(num_tasks=cfg["num_tasks"], task_map=cfg["task_map"], feats_dim=cfg["feats_dim"],
                llm_base_config=cfg["llm_base_config"])
    m.load_state_dict(sd, strict=True)
    return m.cuda(), sd, cfg


@pytest.mark.parametrize("task,B", [("se", 32), ("tse", 16)])
def test_lm_bench_shape_vs_oracle(lib, task, B):
    """UniSE LM at the benchmarked shapes: SR (prefix 252, B=32) and TSE (prefix 503, B=16, KV <= 786), 33 + 250 cached
    steps.  (1) teacher-forced along the ORACLE's token path every step's hidden state is within 1e-3 (no divergence
    ambiguity); (2) free-running greedy generation: the first differing token of any sequence must sit at an oracle
    top-2 margin below 1e-4 (relative to the largest logit)."""
    from oracle import llama
    m, sd, cfg = _lm()
    T = 250
    g = torch.Generator().manual_seed(3000 if task == "se" else 4001)
    mix = torch.randn(B, T, 768, generator=g)
    enr = torch.randn(B, T, 768, generator=g) if task == "tse" else None
    try:
        torch.set_num_threads(min(32, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    og, os_, margins = llama.sft_generate(sd, cfg, task, enr, mix, T, return_margins=True)
    # ---- (1) teacher-forced cached decode along the oracle's path, hidden states vs the oracle's full forward
    goff, soff = 3, 3 + 4096
    col = lambda v: torch.full((B, 1), v, dtype=torch.long)
    ids = torch.cat([col(0), og + goff, col(1), os_[:, :-1] + soff], 1)            # inputs of the 283 steps
    prefix = llama._prefix(sd, cfg, task, enr, mix)
    P = prefix.shape[1]
    assert P == (503 if task == "tse" else 252)
    full = torch.cat([prefix, sd["codec_embedding.weight"][ids]], 1)
    ref, _ = llama.llm_forward(sd, cfg, full)
    enr_d = None if enr is None else enr.cuda()
    pre_d = m._prefix(task, enr_d, mix.cuda())
    e_prefix = rel(pre_d, prefix)
    out = m.llm_forward(pre_d, use_cache=True, max_new_tokens=ids.shape[1])
    cache = out.past_key_values
    e_pre = rel(out.last_hidden_state, ref[:, :P])
    emb = sd["codec_embedding.weight"].cuda()
    worst, worst_t = 0.0, -1
    for t in range(ids.shape[1]):
        h = m.llm_forward(emb[ids[:, t].cuda()][:, None], past_key_values=cache, use_cache=True).last_hidden_state
        e = rel(h[:, 0], ref[:, P + t])
        if e > worst:
            worst, worst_t = e, t
    torch.cuda.synchronize()
    print(f"[lm {task} B={B} P={P}] prefix rel {e_prefix:.2e} prefill rel {e_pre:.2e}; cached decode over {ids.shape[1]} steps: "
          f"worst hidden-state rel {worst:.2e} at step {worst_t} (KV length {P + worst_t + 1})")
    assert e_prefix < TOL and e_pre < TOL and worst < TOL
    # ---- (2) free-running greedy generation
    gg, ss = m.generate(task, enr_d, enr_d, mix.cuda(), mix.cuda(), do_sample=False)
    torch.cuda.synchronize()
    margins = margins.clone()
    margins[:, 32] = 1.0                                  # the 33rd global step's token is discarded (llm_sft.py:139,164)
    zero = torch.zeros(B, 1, dtype=torch.long)
    got, want = torch.cat([gg.cpu(), zero, ss.cpu()], 1), torch.cat([og, zero, os_], 1)
    nbad, n_match = 0, 0
    for b in range(B):
        d = (got[b] != want[b]).nonzero()
        if len(d):
            t = int(d[0])
            nbad += 1
            n_match += t
            print(f"   seq {b}: first divergence at step {t}, oracle top-2 margin {float(margins[b, t]):.2e}")
            assert float(margins[b, t]) < 1e-4, "token differs although the oracle's decision margin is safe"
        else:
            n_match += got.shape[1]
    print(f"[lm {task} B={B}] greedy generate: {B - nbad}/{B} sequences identical over all 283 steps; "
          f"{n_match}/{B * got.shape[1]} tokens identical up to the first (unsafe-margin) divergence; min oracle margin {float(margins.min()):.2e}")
    assert gg.shape == (B, 32) and ss.shape == (B, T)
