"""Torch-tensor front ends of the C-ABI ops.  PyTorch here is device memory + streams only: every
function hands raw device pointers to libquark_b200 on the current CUDA stream."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_ELU, ACT_GELU, ACT_NONE, ACT_SNAKE, ACT_SWIGLU, ACT_TANH, GemmDesc, RowMap  # noqa: F401


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libquark_b200 needs contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


@dataclass
class Planes:
    """fp16 hi (+ optional lo) planes of a channel-last activation / weight."""
    hi: torch.Tensor
    lo: Optional[torch.Tensor] = None

    @staticmethod
    def zeros(shape, split: bool, device):
        hi = torch.zeros(shape, dtype=torch.float16, device=device)
        return Planes(hi, torch.zeros(shape, dtype=torch.float16, device=device) if split else None)

    @staticmethod
    def from_f32(x: torch.Tensor, split: bool = True):
        """Load-time helper (weights): hi = rn_fp16(x) saturated, lo = rn_fp16(x - hi)."""
        x = x.float().clamp(-65504.0, 65504.0)
        hi = x.half()
        lo = (x - hi.float()).half() if split else None
        return Planes(hi.contiguous(), lo.contiguous() if lo is not None else None)

    def float(self):
        return self.hi.float() + (self.lo.float() if self.lo is not None else 0.0)


def rowmap(t: Optional[torch.Tensor], ld=0, rows_per_batch=0, row_off=0) -> RowMap:
    return RowMap(_p(t), ld, rows_per_batch, row_off)


def gemm(a: Planes, w: Planes, n: int, *, a_batch: int, a_rows_per_batch: int, a_ld: int, m_per_batch: int,
         taps: int = 1, stride: int = 1, bias=None, gamma=None, residual: Optional[RowMap] = None, act=ACT_NONE,
         act2=ACT_NONE, out_f32: Optional[RowMap] = None, out_planes: Optional[Planes] = None,
         out_planes_map=(0, 0, 0), simt: bool = False, dilation: int = 1, act_param=None, act2_param=None, a_cols: int = 0,
         a_col_off: int = 0):
    """One dense contraction (see qb_gemm_desc).  Split mode iff both a.lo and w.lo are given."""
    split = a.lo is not None and w.lo is not None
    d = GemmDesc()
    d.a_hi, d.a_lo = _p(a.hi), (_p(a.lo) if split else None)
    if a_col_off:       # grouped convolution: contract channels [a_col_off, a_col_off + a_cols) of every row
        d.a_hi = C.c_void_p(a.hi.data_ptr() + 2 * a_col_off)
        d.a_lo = C.c_void_p(a.lo.data_ptr() + 2 * a_col_off) if split else None
    d.a_cols = a_cols
    d.a_batch, d.a_rows_per_batch, d.a_ld = a_batch, a_rows_per_batch, a_ld
    d.taps, d.stride, d.m_per_batch, d.dilation = taps, stride, m_per_batch, dilation
    d.w_hi, d.w_lo, d.n = _p(w.hi), (_p(w.lo) if split else None), n
    d.bias, d.gamma = _p(bias), _p(gamma)
    d.residual = residual if residual is not None else RowMap(None, 0, 0, 0)
    d.act, d.act2 = act, act2
    d.act_param, d.act2_param = _p(act_param), _p(act2_param)
    d.out_f32 = out_f32 if out_f32 is not None else RowMap(None, 0, 0, 0)
    if out_planes is not None:
        ld, rpb, off = out_planes_map
        d.out_hi = RowMap(_p(out_planes.hi), ld, rpb, off)
        d.out_lo = RowMap(_p(out_planes.lo), ld, rpb, off) if out_planes.lo is not None else RowMap(None, 0, 0, 0)
    else:
        d.out_hi = RowMap(None, 0, 0, 0)
        d.out_lo = RowMap(None, 0, 0, 0)
    lib = _lib.load()
    _lib.check((lib.qb_gemm_simt if simt else lib.qb_gemm)(C.byref(d), _stream()))


def gemm_kernel_name(m_per_batch: int, n: int, split: bool) -> str:
    return _lib.load().qb_gemm_kernel_name(m_per_batch, n, int(split)).decode()


def split_f16(x: torch.Tensor, out: Planes):
    _lib.check(_lib.load().qb_split_f16(_p(x), _p(out.hi), _p(out.lo), x.numel(), _stream()))


def rows_to_planes(x, B, rows, Cc, out: Planes, ld, rows_per_batch, row_off, repeat=1, act=ACT_NONE):
    _lib.check(_lib.load().qb_rows_to_planes(_p(x), B, rows, Cc, repeat, act, _p(out.hi), _p(out.lo), ld,
                                             rows_per_batch, row_off, _stream()))


def bct_to_planes(x, out: Planes, ld, rows_per_batch, row_off):
    B, Cc, T = x.shape
    _lib.check(_lib.load().qb_bct_to_planes(_p(x), B, Cc, T, _p(out.hi), _p(out.lo), ld, rows_per_batch, row_off,
                                            _stream()))


def layernorm(x, w, b, B, rows, Cc, eps=1e-6, out_f32=None, out: Optional[Planes] = None, ld=0, rows_per_batch=0,
              row_off=0):
    hi = out.hi if out is not None else None
    lo = out.lo if out is not None else None
    if out is not None and ld == 0:
        ld, rows_per_batch, row_off = Cc, rows, 0
    _lib.check(_lib.load().qb_layernorm(_p(x), _p(w), _p(b), eps, B, rows, Cc, _p(out_f32), _p(hi), _p(lo), ld,
                                        rows_per_batch, row_off, _stream()))


def rmsnorm(x, w, rows, Cc, out: Optional[Planes] = None, eps=1e-6, out_f32=None):
    hi = out.hi if out is not None else None
    lo = out.lo if out is not None else None
    _lib.check(_lib.load().qb_rmsnorm(_p(x), _p(w), eps, rows, Cc, _p(out_f32), _p(hi), _p(lo), _stream()))


def dwconv7_ln(x, dw_w, dw_b, ln_w, ln_b, B, T, Cc, out: Planes):
    _lib.check(_lib.load().qb_dwconv7_ln(_p(x), _p(dw_w), _p(dw_b), _p(ln_w), _p(ln_b), B, T, Cc, _p(out.hi), _p(out.lo),
                                         _stream()))


def dwconv7_adaln(x, dw_w, dw_b, scale, shift, cond_stride, B, T, Cc, out: Planes):
    _lib.check(_lib.load().qb_dwconv7_adaln(_p(x), _p(dw_w), _p(dw_b), _p(scale), _p(shift), cond_stride, B, T, Cc, _p(out.hi),
                                            _p(out.lo), _stream()))


def adalayernorm(x, scale, shift, cond_stride, B, rows, Cc, eps=1e-6, out_f32=None, out: Optional[Planes] = None, ld=0,
                 rows_per_batch=0, row_off=0):
    hi = out.hi if out is not None else None
    lo = out.lo if out is not None else None
    if out is not None and ld == 0:
        ld, rows_per_batch, row_off = Cc, rows, 0
    _lib.check(_lib.load().qb_adalayernorm(_p(x), _p(scale), _p(shift), cond_stride, eps, B, rows, Cc, _p(out_f32), _p(hi),
                                           _p(lo), ld, rows_per_batch, row_off, _stream()))


def snake_planes(x, x_batch_stride, alpha, B, T, Cc, out: Planes, ld, rows_per_batch, row_off):
    _lib.check(_lib.load().qb_snake_planes(_p(x), x_batch_stride, _p(alpha), B, T, Cc, _p(out.hi), _p(out.lo), ld,
                                           rows_per_batch, row_off, _stream()))


def addvec_planes(x, vec, B, T, Cc, out: Planes, ld, rows_per_batch, row_off):
    _lib.check(_lib.load().qb_addvec_planes(_p(x), _p(vec), B, T, Cc, _p(out.hi), _p(out.lo), ld, rows_per_batch, row_off,
                                            _stream()))


def groupnorm_stats(x, B, T, Cc, stats, groups=32, eps=1e-6):
    _lib.check(_lib.load().qb_groupnorm_stats(_p(x), B, T, Cc, groups, eps, _p(stats), _stream()))


def groupnorm_apply(x, stats, w, b, B, T, Cc, swish, out_f32=None, out: Optional[Planes] = None, ld=0,
                    rows_per_batch=0, row_off=0, groups=32):
    hi = out.hi if out is not None else None
    lo = out.lo if out is not None else None
    _lib.check(_lib.load().qb_groupnorm_apply(_p(x), _p(stats), _p(w), _p(b), B, T, Cc, groups, int(swish), _p(out_f32),
                                              _p(hi), _p(lo), ld, rows_per_batch, row_off, _stream()))


def wav_to_hopblocks(wav, hop, out: Planes):
    B, T = wav.shape
    _lib.check(_lib.load().qb_wav_to_hopblocks(_p(wav), B, T, hop, _p(out.hi), _p(out.lo), _stream()))


def stft_post(spec, ld_spec, B, frames, nf, out: Planes, ld, rows_per_batch, row_off):
    _lib.check(_lib.load().qb_stft_post(_p(spec), ld_spec, B, frames, nf, _p(out.hi), _p(out.lo), ld, rows_per_batch,
                                        row_off, _stream()))


def stft_gather(wav, hop, n_fft, P, Q, window, out: Planes):
    B, T = wav.shape
    _lib.check(_lib.load().qb_stft_gather(_p(wav), B, T, hop, n_fft, P, Q, _p(window), _p(out.hi), _p(out.lo), _stream()))


def stft_twiddle(Y, ldY, frames_total, P, Q, twiddle, out: Planes):
    _lib.check(_lib.load().qb_stft_twiddle(_p(Y), ldY, frames_total, P, Q, _p(twiddle), _p(out.hi), _p(out.lo), _stream()))


def stft_post2(X, ldX, B, frames, nf, P, out: Planes, ld, rows_per_batch, row_off):
    _lib.check(_lib.load().qb_stft_post2(_p(X), ldX, B, frames, nf, P, _p(out.hi), _p(out.lo), ld, rows_per_batch, row_off, _stream()))


def istft_pre(head, ld_in, M, nf, out: Planes, ld):
    _lib.check(_lib.load().qb_istft_pre(_p(head), ld_in, M, nf, _p(out.hi), _p(out.lo), ld, _stream()))


def istft_ola(frames, window, B, F, n_fft, wav, hop=None):
    _lib.check(_lib.load().qb_istft_ola(_p(frames), _p(window), B, F, n_fft, hop if hop is not None else n_fft // 2, _p(wav),
                                        _stream()))


def reflect_pad_rows(buf: Planes, B, rows_per_batch, ld, T, row_off, pad_l, pad_r):
    _lib.check(_lib.load().qb_reflect_pad_rows(_p(buf.hi), _p(buf.lo), B, rows_per_batch, ld, T, row_off, pad_l, pad_r,
                                               _stream()))


def dwconv(x, w, bias, B, T, Cc, k, out):
    _lib.check(_lib.load().qb_dwconv(_p(x), _p(w), _p(bias), B, T, Cc, k, _p(out), _stream()))


def attention_hd(qkv, B, T, heads, head_dim, rope_cos, rope_sin, out: Planes):
    _lib.check(_lib.load().qb_attention_hd(_p(qkv), B, T, heads, head_dim, _p(rope_cos), _p(rope_sin), _p(out.hi), _p(out.lo),
                                           _stream()))


def attention(qkv, B, T, heads, rope_cos, rope_sin, out: Planes):
    _lib.check(_lib.load().qb_attention(_p(qkv), B, T, heads, _p(rope_cos), _p(rope_sin), _p(out.hi), _p(out.lo),
                                        _stream()))


def attention_tc_workspace_bytes(B, T, heads):
    return int(_lib.load().qb_attention_tc_workspace_bytes(B, T, heads))


def attention_tc(qkv, B, T, heads, rope_cos, rope_sin, out: Planes, workspace):
    _lib.check(_lib.load().qb_attention_tc(_p(qkv), B, T, heads, _p(rope_cos), _p(rope_sin), _p(out.hi), _p(out.lo),
                                           _p(workspace), _stream()))


def attention_umma_workspace_bytes(B, L, heads, head_dim, split):
    return int(_lib.load().qb_attention_umma_workspace_bytes(B, L, heads, head_dim, int(bool(split))))


def attention_umma(qkv, B, L, heads, head_dim, rope_cos, rope_sin, out: Planes, workspace, split=None, causal=False):
    """tcgen05 attention (head_dim 64 / 128); split defaults to whether `out` carries a lo plane"""
    split = (out.lo is not None) if split is None else split
    _lib.check(_lib.load().qb_attention_umma(_p(qkv), B, L, heads, head_dim, _p(rope_cos), _p(rope_sin), _p(out.hi), _p(out.lo),
                                             int(bool(split)), int(bool(causal)), _p(workspace), _stream()))


def lstm_workspace_bytes(B, H):
    return int(_lib.load().qb_lstm_workspace_bytes(B, H))


def lstm(xp, whh: Planes, B, T, H, out: Planes, workspace):
    _lib.check(_lib.load().qb_lstm(_p(xp), _p(whh.hi), None, B, T, H, _p(out.hi), _p(out.lo), _p(workspace), _stream()))


def lstm_tc_units(H):
    return int(_lib.load().qb_lstm_tc_units(H))


def lstm_tc_workspace_bytes(B, H):
    return int(_lib.load().qb_lstm_tc_workspace_bytes(B, H))


def lstm_tc_permute(whh: torch.Tensor, U: int) -> torch.Tensor:
    """[4H,H] (gate-major i|f|g|o) -> fp16 [H/U][4U][H], row 4j+g of slice c = gate g of unit c*U+j."""
    H = whh.shape[1]
    rows = (torch.arange(4, device=whh.device)[None, None, :] * H +
            torch.arange(H, device=whh.device).reshape(H // U, U)[:, :, None]).reshape(-1)
    return whh.float().clamp(-65504.0, 65504.0)[rows].half().contiguous()


def lstm_tc(xp, whh_perm, U, B, T, H, out: Planes, workspace):
    _lib.check(_lib.load().qb_lstm_tc(_p(xp), _p(whh_perm), U, B, T, H, _p(out.hi), _p(out.lo), _p(workspace), _stream()))


def rvq_workspace_bytes(M, D, K):
    return int(_lib.load().qb_rvq_workspace_bytes(M, D, K))


def rvq_encode(x, codebooks, cb: Planes, neg_half_e2, e2max, M, D, K, nq, idx, quantized, workspace):
    _lib.check(_lib.load().qb_rvq_encode(_p(x), _p(codebooks), _p(cb.hi), _p(cb.lo), _p(neg_half_e2), float(e2max), M, D,
                                         K, nq, _p(idx), _p(quantized), _p(workspace), _stream()))


def rvq_decode(idx, codebooks, M, D, K, nq, out, out_ld, col_off):
    _lib.check(_lib.load().qb_rvq_decode(_p(idx), _p(codebooks), M, D, K, nq, _p(out), out_ld, col_off, _stream()))


def lm_qkv_prep(qkv, B, L, heads, pos0, cos, sin, q16, kc, vc, Lmax):
    _lib.check(_lib.load().qb_lm_qkv_prep(_p(qkv), B, L, heads, pos0, _p(cos), _p(sin), _p(q16), _p(kc), _p(vc), Lmax,
                                          _stream()))


def lm_flash_attn(q16, kc, vc, B, L, heads, pos0, Lmax, out: Planes):
    _lib.check(_lib.load().qb_lm_flash_attn(_p(q16), _p(kc), _p(vc), B, L, heads, pos0, Lmax, _p(out.hi), _p(out.lo),
                                            _stream()))


def lm_decode_layer(x, B, hidden, heads, inter, L, kc, vc, Lmax, pos, cos, sin, q_buf, attn_buf, mlp_buf):
    _lib.check(_lib.load().qb_lm_decode_layer(_p(x), B, hidden, heads, inter, _p(L["in_w"]), _p(L["wqkv32"]), _p(L["wo32"]),
                                              _p(L["post_w"]), _p(L["wg32"]), _p(L["wu32"]), _p(L["wd32"]), _p(kc), _p(vc),
                                              Lmax, _p(pos), _p(cos), _p(sin), _p(q_buf), _p(attn_buf), _p(mlp_buf),
                                              _stream()))


def lm_head_argmax(x, B, hidden, final_norm, w_head, rng, max_cols, emb, x_next, out_ids, out_stride, pos, slot, pv, pi):
    _lib.check(_lib.load().qb_lm_head_argmax(_p(x), B, hidden, _p(final_norm), _p(w_head), _p(rng), max_cols, _p(emb),
                                             _p(x_next), _p(out_ids), out_stride, _p(pos), _p(slot), _p(pv), _p(pi),
                                             _stream()))


def lm_pack_weight(w: torch.Tensor) -> torch.Tensor:
    """fp32 [n,k] -> fp16 [n,2k] groups of {hi[4], lo[4]} (decode B-fragment layout, include/quark_b200.h)."""
    w = w.float().contiguous()
    n, k = w.shape
    out = torch.empty(n, 2 * k, dtype=torch.float16, device=w.device)
    _lib.check(_lib.load().qb_lm_pack_weight(_p(w), n, k, _p(out), _stream()))
    return out


def lm_set_att_unroll(keys_per_lane: int):
    _lib.check(_lib.load().qb_lm_set_att_unroll(int(keys_per_lane)))


def lm_decode_layer_tc(x, B, hidden, heads, inter, L, kc, vc, Lmax, pos, cos, sin, q_buf, attn_buf, mlp_buf):
    _lib.check(_lib.load().qb_lm_decode_layer_tc(_p(x), B, hidden, heads, inter, _p(L["wqkv_p"]), _p(L["wo_p"]), _p(L["wg_p"]),
                                                 _p(L["wu_p"]), _p(L["wd_p"]), _p(kc), _p(vc), Lmax, _p(pos), _p(cos), _p(sin),
                                                 _p(q_buf), _p(attn_buf), _p(mlp_buf), _stream()))


def lm_head_argmax_tc(x, B, hidden, w_head_p, rng, max_cols, emb, x_next, out_ids, out_stride, pos, slot, pv, pi):
    _lib.check(_lib.load().qb_lm_head_argmax_tc(_p(x), B, hidden, _p(w_head_p), _p(rng), max_cols, _p(emb), _p(x_next),
                                                _p(out_ids), out_stride, _p(pos), _p(slot), _p(pv), _p(pi), _stream()))


def lm_head_sample_tc(x, B, hidden, w_head_p, rng, max_cols, emb, x_next, out_ids, out_stride, pos, slot, pv, pi, logits,
                      temperature, top_k, top_p, seed, debug=None):
    _lib.check(_lib.load().qb_lm_head_sample_tc(_p(x), B, hidden, _p(w_head_p), _p(rng), max_cols, _p(emb), _p(x_next),
                                                _p(out_ids), out_stride, _p(pos), _p(slot), _p(pv), _p(pi), _p(logits),
                                                float(temperature), int(top_k), float(top_p), _p(seed), _p(debug), _stream()))


def ssl_conv0_gn_gelu(x, w, gn_w, gn_b, eps, k, stride, out: Planes, ld, rows_per_batch, row_off, y_scratch, workspace):
    B, T_in = x.shape
    _lib.check(_lib.load().qb_ssl_conv0_gn_gelu(_p(x), B, T_in, _p(w), w.shape[0], k, stride, _p(gn_w), _p(gn_b), float(eps), _p(y_scratch),
                                                _p(workspace), _p(out.hi), _p(out.lo), ld, rows_per_batch, row_off, _stream()))


def ssl_conv0_workspace_bytes(B, T0, Cc):
    return int(_lib.load().qb_ssl_conv0_workspace_bytes(B, T0, Cc))


def wavlm_gate(x, B, T, heads, head_dim, w, bias, cst, gate):
    _lib.check(_lib.load().qb_wavlm_gate(_p(x), B, T, heads, head_dim, _p(w), _p(bias), _p(cst), _p(gate), _stream()))


def attention_relbias(qkv, B, T, heads, head_dim, rel_table, gate, out: Planes):
    _lib.check(_lib.load().qb_attention_relbias(_p(qkv), B, T, heads, head_dim, _p(rel_table), _p(gate), _p(out.hi), _p(out.lo), _stream()))


def axpy(x, scale, out, accumulate=True):
    _lib.check(_lib.load().qb_axpy(_p(x), float(scale), x.numel(), int(accumulate), _p(out), _stream()))


def ssl_compress(x, B, T, Cc, power, channel_first, out):
    _lib.check(_lib.load().qb_ssl_compress(_p(x), B, T, Cc, float(power), int(channel_first), _p(out), _stream()))


def pad_wav(x, left, T_out, wrap=False):
    """[B, T] fp32 -> [B, T_out]: out[b, i] = x[b, i - left], zero (or wrapped) outside"""
    x = x.float().contiguous()
    B, T = x.shape
    out = torch.empty(B, T_out, device=x.device)
    _lib.check(_lib.load().qb_pad_wav(_p(x), B, T, left, T_out, int(wrap), _p(out), _stream()))
    return out


def ptr_array(tensors):
    """host array of device pointers (kept alive by the caller)"""
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def lm_decode_steps(x, B, hidden, heads, inter, layer_ptrs, Lmax, w_head_p, rng, max_cols, emb, cos, sin, q_buf, attn_buf, mlp_buf, pv, pi,
                    out_ids, out_stride, pos, slot, n_steps, barrier):
    """layer_ptrs: dict of ctypes pointer arrays wqkv / wo / wg / wu / wd / k / v (ptr_array)"""
    L = layer_ptrs
    _lib.check(_lib.load().qb_lm_decode_steps(_p(x), B, hidden, heads, inter, L["n"], L["wqkv"], L["wo"], L["wg"], L["wu"], L["wd"], L["k"], L["v"],
                                              Lmax, _p(w_head_p), _p(rng), max_cols, _p(emb), _p(cos), _p(sin), _p(q_buf), _p(attn_buf),
                                              _p(mlp_buf), _p(pv), _p(pi), _p(out_ids), out_stride, _p(pos), _p(slot), n_steps, _p(barrier),
                                              _stream()))


def lm_loss(logits, ld, M, V, targets, label_smoothing):
    """-> float32 [2] = {label-smoothed KL (batchmean), arg-max accuracy}"""
    ws = torch.empty(2 * M, device=logits.device)
    out = torch.empty(2, device=logits.device)
    _lib.check(_lib.load().qb_lm_loss(_p(logits), ld, M, V, _p(targets), float(label_smoothing), _p(ws), _p(out), _stream()))
    return out


def launch_count() -> int:
    return int(_lib.load().qb_launch_count())


def launch_count_reset():
    _lib.load().qb_launch_count_reset()
