"""UniSE AR-LM (`LLM_SFT`) with the reference's surface, running on libquark_b200.

Mirrors QuarkAudio-UniSE/model/llm/llm_sft.py:13-195 and llm.py:13-228:
    LLM_SFT(num_tasks, task_map, feats_dim, llm_base_config)
    .llm_forward(inputs_embeds, past_key_values=None, use_cache=False) -> .last_hidden_state / .past_key_values
    .forward(task_name, enroll_mel, enroll_feats, mix_mel, mix_feats, global_ids, semantic_ids) -> (loss, acc)
    .generate(task_name, enroll_mel, enroll_feats, mix_mel, mix_feats, global_length=32, ..., do_sample) -> (global, semantic)
state_dict keys are the reference's (HF Llama layer names under `layers.*`, `norm.weight`, `codec_embedding`,
`output_head`, `adapter`, `task_embedding`, `enroll_sos_embedding`, `mix_sos_embedding`); the conformer
condition encoder (`cond_*`, built but never executed: llm_sft.py:62-65,112-115) is accepted at load and ignored.

Prefill / teacher-forced forward: tcgen05 GEMMs (3-term split) + causal split-precision flash attention over a static fp32 KV cache.
Decode: fused skinny kernels (3-term fp16-split mma.sync over pre-packed weights, programmatic dependent launch; the fp32
SIMT versions are the cross-check, QB_LM_DECODE=simt), one CUDA graph per step replayed 33 + T times; greedy (do_sample=False, the shipped
setting U/model/model.py:173).  No PyTorch / CPU fallback for the transformer stack.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
import os
from typing import Optional

import torch
from torch import nn

from . import ops
from .codec import _Tree
from .ops import ACT_SWIGLU, Planes, rowmap


def lm_spec(cfg_base: dict, num_tasks: int, feats_dim: int):
    H, L = cfg_base["hidden_size"], cfg_base["num_layers"]
    V = 3 + cfg_base["global_size"] + cfg_base["semantic_size"]
    out = {"mix_sos_embedding.weight": (1, H), "codec_embedding.weight": (V, H)}
    for i in range(L):
        p = f"layers.{i}."
        for n in "qkvo":
            out[p + f"self_attn.{n}_proj.weight"] = (H, H)
        out[p + "mlp.gate_proj.weight"] = (4 * H, H)
        out[p + "mlp.up_proj.weight"] = (4 * H, H)
        out[p + "mlp.down_proj.weight"] = (H, 4 * H)
        out[p + "input_layernorm.weight"] = (H,)
        out[p + "post_attention_layernorm.weight"] = (H,)
    out["norm.weight"] = (H,)
    out["output_head.weight"] = (V, H)
    out["task_embedding.weight"] = (num_tasks, H)
    out["enroll_sos_embedding.weight"] = (1, H)
    out["adapter.weight"] = (H, feats_dim)
    out["adapter.bias"] = (H,)
    return out


class StaticKVCache:
    """Pre-allocated fp32 cache [layers][B, heads, Lmax, 64]; replaces HF DynamicCache's torch.cat growth."""

    def __init__(self, layers, B, heads, Lmax, device):
        self.k = [torch.zeros(B, heads, Lmax, 64, dtype=torch.float32, device=device) for _ in range(layers)]
        self.v = [torch.zeros(B, heads, Lmax, 64, dtype=torch.float32, device=device) for _ in range(layers)]
        self.B, self.heads, self.layers, self.Lmax, self.length = B, heads, layers, Lmax, 0
        self.pos = torch.zeros(1, dtype=torch.int32, device=device)     # device copy used by the decode kernels

    def get_seq_length(self):
        return self.length

    def reserve(self, n_positions: int):
        """Make room for `n_positions` in total (HF DynamicCache grows without bound, llm.py:189-193): reallocate to the
        next multiple of 256 and copy the filled prefix.  Captured graphs never call this (their capacity is fixed)."""
        if n_positions <= self.Lmax:
            return
        new = -(-n_positions // 256) * 256
        for buf in (self.k, self.v):
            for i, t in enumerate(buf):
                g = torch.zeros(self.B, self.heads, new, 64, dtype=t.dtype, device=t.device)
                g[:, :, :self.length] = t[:, :, :self.length]
                buf[i] = g
        self.Lmax = new


@dataclass
class LMOutput:
    last_hidden_state: torch.Tensor
    past_key_values: Optional[StaticKVCache] = None


class LLM_SFT(nn.Module):
    def __init__(self, num_tasks: int = 1, task_map: dict = None, feats_dim: int = 768, llm_base_config: dict = None):
        super().__init__()
        b = dict(llm_base_config or {})
        self.cfg = b
        self.task_map = dict(task_map or {"se": 0})
        self.hidden, self.n_layers, self.heads = b["hidden_size"], b["num_layers"], b["num_attention_heads"]
        if self.hidden != self.heads * 64 or self.hidden % 128:
            raise ValueError("kernels assume head_dim 64 and hidden % 128 == 0 (shipped config: 512 = 8 x 64)")
        self.global_size, self.semantic_size = b["global_size"], b["semantic_size"]
        self.vocab_size = 3 + self.global_size + self.semantic_size
        self.global_offset, self.semantic_offset = 3, 3 + self.global_size
        self.global_sos_token_id, self.semantic_sos_token_id, self.semantic_eos_token_id = 0, 1, 2
        self.label_smoothing = b.get("label_smoothing", 0.1)
        self.max_pos = b.get("max_position_embeddings", 4096)
        tree = _Tree.build(lm_spec(b, num_tasks, feats_dim))
        for name, child in tree.named_children():
            self.add_module(name, child)
        self._w, self._ws = None, {}
        # "persistent": the whole decoding loop in one cooperative kernel (csrc/llm_step.cu; greedy, shipped dimensions);
        # "tc": one kernel per stage, packed fp16-split weights + mma.sync + programmatic dependent launch, 8 steps per CUDA graph
        # (also the sampled-decoding path);  "simt": fp32 cross-check
        self.decode_kernel = os.environ.get("QB_LM_DECODE", "tc")
        self.graph_steps = int(os.environ.get("QB_LM_GRAPH_STEPS", "8"))     # decode steps per replayed CUDA graph
        self._gen_state = {}
        # generate() walks a batch in chunks of <= `chunk` sequences; `lanes` > 1 runs that many chunks CONCURRENTLY, each on its own
        # CUDA stream with its own KV cache / workspace / captured graphs (the decode step is a chain of ~62 short dependent kernels:
        # one chain leaves most of the GPU idle, independent chains fill it).  Tokens do not depend on either setting.
        # Measured on B200 (profiles/r02_lm_lanes_ab.md), 256 sequences: 851.7 ms serial, 584.9 / 513.5 / 511.0 ms with 2 / 4 / 8 lanes;
        # cutting a batch of <= 32 into smaller chunks is slower (a chain costs the same for 8, 16 or 32 rows), hence chunk = 32.
        self.lanes = max(1, int(os.environ.get("QB_LM_LANES", "4")))
        # K / V rows a lane of the decode attention keeps in flight: 8 for one chain (latency-bound), 4 on concurrent lanes
        # (throughput-bound); QB_LM_ATT_U pins it.  Fixed per decode state (it is baked into the captured graphs).
        self.att_unroll = int(os.environ.get("QB_LM_ATT_U", "8"))
        self.lane_att_unroll = int(os.environ.get("QB_LM_ATT_U", "4"))
        self.chunk = min(32, max(1, int(os.environ.get("QB_LM_CHUNK", "32"))))
        self._lane_views = None
        self.eval()

    # ------------------------------------------------------------------ state
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = {k: v for k, v in state_dict.items() if not k.startswith(("cond_", "rotary_emb."))}
        r = super().load_state_dict(sd, strict=strict, assign=assign)
        self._w, self._gen_state, self._lane_views = None, {}, None          # captured graphs point at the old prepared weights
        return r

    def _apply(self, fn, *a, **k):
        self._w, self._ws, self._gen_state, self._lane_views = None, {}, {}, None
        return super()._apply(fn, *a, **k)

    def _dev(self):
        return self.norm.weight.device

    def _prepare(self):
        if self._w is not None:
            return self._w
        dev = self._dev()
        if dev.type != "cuda":
            raise RuntimeError("unified_audio_b200.LLM_SFT runs on CUDA only (no CPU fallback): call .cuda() first")
        sd = {k: v.detach().float() for k, v in self.state_dict().items()}
        layers = []
        for i in range(self.n_layers):
            p = f"layers.{i}."
            wqkv = torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0).contiguous()
            wg, wu = sd[p + "mlp.gate_proj.weight"].contiguous(), sd[p + "mlp.up_proj.weight"].contiguous()
            layers.append(dict(
                in_w=sd[p + "input_layernorm.weight"].contiguous(), post_w=sd[p + "post_attention_layernorm.weight"].contiguous(),
                wqkv=Planes.from_f32(wqkv, True), wo=Planes.from_f32(sd[p + "self_attn.o_proj.weight"], True),
                wgu=Planes.from_f32(torch.stack([wg, wu], 1).reshape(-1, self.hidden), True),
                wd=Planes.from_f32(sd[p + "mlp.down_proj.weight"], True),
                # decode path: fp32 weights with the preceding RMSNorm weight folded in (W' = W diag(g))
                wqkv32=(wqkv * sd[p + "input_layernorm.weight"][None, :]).contiguous(),
                wo32=sd[p + "self_attn.o_proj.weight"].contiguous(),
                wg32=(wg * sd[p + "post_attention_layernorm.weight"][None, :]).contiguous(),
                wu32=(wu * sd[p + "post_attention_layernorm.weight"][None, :]).contiguous(),
                wd32=sd[p + "mlp.down_proj.weight"].contiguous()))
            L = layers[-1]       # product decode path: the same folded weights packed as fp16 {hi[4], lo[4]} groups
            L.update(wqkv_p=ops.lm_pack_weight(L["wqkv32"]), wo_p=ops.lm_pack_weight(L["wo32"]), wg_p=ops.lm_pack_weight(L["wg32"]),
                     wu_p=ops.lm_pack_weight(L["wu32"]), wd_p=ops.lm_pack_weight(L["wd32"]))
        self._w = dict(layers=layers, norm=sd["norm.weight"].contiguous(), head=Planes.from_f32(sd["output_head.weight"], True),
                       head32=(sd["output_head.weight"] * sd["norm.weight"][None, :]).contiguous(), emb=sd["codec_embedding.weight"].contiguous(),
                       adapter=Planes.from_f32(sd["adapter.weight"], True), adapter_b=sd["adapter.bias"].contiguous(),
                       cos=None, sin=None, rope_rows=0)
        self._w["head_p"] = ops.lm_pack_weight(self._w["head32"])
        self._ensure_rope(self.max_pos)
        return self._w

    def _ensure_rope(self, n_positions: int):
        """cos/sin tables cover positions [0, rope_rows).  The reference's HF rotary embedding computes them from position_ids
        on the fly and has no length limit (llm.py:187), so the table grows on demand (the kernels index it unguarded:
        every caller checks pos0 + L against it first).  Growing replaces the tensors, so captured decode graphs that hold
        the old pointers are dropped."""
        W = self._w
        if W["rope_rows"] >= n_positions:
            return
        rows = max(self.max_pos, -(-n_positions // 1024) * 1024)
        inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.int64).float() / 64))
        fr = torch.arange(rows).float()[:, None] * inv[None, :]
        emb = torch.cat((fr, fr), -1)
        W["cos"], W["sin"], W["rope_rows"] = emb.cos().to(self._dev()).contiguous(), emb.sin().to(self._dev()).contiguous(), rows
        self._gen_state = {}

    def _buf(self, name, shape, dtype=torch.float32):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self._dev())
            self._ws[key] = t
        return t

    def _planes(self, name, shape):
        key = ("P", name, tuple(shape))
        p = self._ws.get(key)
        if p is None:
            p = Planes.zeros(shape, True, self._dev())
            self._ws[key] = p
        return p

    # ------------------------------------------------------------------ transformer stack
    def _prefill(self, x: torch.Tensor, B: int, L: int, cache: StaticKVCache):
        """x [B*L, hidden] fp32, updated in place by the 12 layers; K/V written at cache.length..  cache None = a teacher-forced
        forward that nobody will decode from: no KV cache is allocated or written (tcgen05 attention only)."""
        W = self._prepare()
        H, heads, inter, M = self.hidden, self.heads, 4 * self.hidden, B * L
        pos0 = cache.length if cache is not None else 0
        if cache is not None:
            if pos0 + L > cache.Lmax:
                raise ValueError("KV cache too small")
            if cache.B != B or cache.heads != heads or cache.layers != self.n_layers:
                raise ValueError(f"KV cache built for batch {cache.B} x {cache.heads} heads x {cache.layers} layers, "
                                 f"got batch {B} x {heads} heads x {self.n_layers} layers")
        self._ensure_rope(pos0 + L)
        W = self._w
        t1 = self._planes("t1", (M, H))
        hid = self._planes("hid", (M, inter))
        qkv = self._buf("qkv", (M, 3 * H))
        q16 = self._buf("q32", (B, heads, L, 64)) if cache is not None else None
        xm = rowmap(x, H, M, 0)
        lin = lambda a, w, n, K, **kw: ops.gemm(a, w, n, a_batch=1, a_rows_per_batch=M, a_ld=K, m_per_batch=M, **kw)
        # a prefill from an empty cache (every call of llm_forward / forward / generate) attends within its own L positions: the causal
        # tcgen05 attention (csrc/attention_umma.cu) reads the qkv GEMM's output directly; lm_qkv_prep still fills the fp32 KV cache for
        # the decode steps.  A continuation (pos0 > 0) keeps the cache-reading mma.sync kernel.
        umma = pos0 == 0 and os.environ.get("QB_ATTENTION", "umma") != "legacy"
        if cache is None and not umma:
            raise RuntimeError("cache-less prefill needs the tcgen05 attention path")
        att_ws = self._buf("att5_ws", (ops.attention_umma_workspace_bytes(B, L, heads, 64, True),), torch.uint8) if umma else None
        for i, Lw in enumerate(W["layers"]):
            ops.rmsnorm(x, Lw["in_w"], M, H, t1)
            lin(t1, Lw["wqkv"], 3 * H, H, out_f32=rowmap(qkv, 3 * H, M, 0))
            if cache is not None:
                ops.lm_qkv_prep(qkv, B, L, heads, pos0, W["cos"], W["sin"], q16, cache.k[i], cache.v[i], cache.Lmax)
            if umma:
                ops.attention_umma(qkv, B, L, heads, 64, W["cos"], W["sin"], t1, att_ws, split=True, causal=True)
            else:
                ops.lm_flash_attn(q16, cache.k[i], cache.v[i], B, L, heads, pos0, cache.Lmax, t1)
            lin(t1, Lw["wo"], H, H, residual=xm, out_f32=xm)
            ops.rmsnorm(x, Lw["post_w"], M, H, t1)
            lin(t1, Lw["wgu"], 2 * inter, H, act=ACT_SWIGLU, out_planes=hid, out_planes_map=(inter, M, 0))
            lin(hid, Lw["wd"], H, inter, residual=xm, out_f32=xm)
        if cache is not None:
            cache.length = pos0 + L
            cache.pos.fill_(cache.length)

    def _decode_layers(self, x: torch.Tensor, B: int, cache: StaticKVCache):
        W = self._prepare()
        if self.decode_kernel == "tc":
            ops.lm_set_att_unroll(self.att_unroll)
        H, heads, inter = self.hidden, self.heads, 4 * self.hidden
        qb, ab, mb = self._buf("dq", (B, H)), self._buf("da", (B, H)), self._buf("dm", (B, inter))
        layer = ops.lm_decode_layer_tc if self.decode_kernel == "tc" else ops.lm_decode_layer
        for i, Lw in enumerate(W["layers"]):
            layer(x, B, H, heads, inter, Lw, cache.k[i], cache.v[i], cache.Lmax, cache.pos, W["cos"], W["sin"], qb, ab, mb)

    @torch.no_grad()
    def llm_forward(self, inputs_embeds, attention_mask=None, past_key_values: Optional[StaticKVCache] = None,
                    use_cache: bool = False, max_new_tokens: Optional[int] = None, **unused) -> LMOutput:
        """llm.py:150-228 (mask None + SDPA == causal).  `max_new_tokens` (optional) sizes a newly created cache; a cache
        that fills up is grown (reallocate + copy) like the reference's DynamicCache."""
        if attention_mask is not None:
            raise NotImplementedError("only the reference's causal (mask=None) path is implemented")
        W = self._prepare()
        B, L, H = inputs_embeds.shape
        cache = past_key_values
        if cache is None and not use_cache and os.environ.get("QB_ATTENTION", "umma") != "legacy":
            # teacher-forced forward (LLM_SFT.forward, llm_sft.py:93-135): nothing decodes from it, so no KV cache at all
            self._ensure_rope(L)
            x = inputs_embeds.float().reshape(B * L, H).contiguous().clone()
            self._prefill(x, B, L, None)
            out = torch.empty(B * L, H, device=x.device)
            ops.rmsnorm(x, self._w["norm"], B * L, H, out_f32=out)
            return LMOutput(out.reshape(B, L, H), None)
        if cache is None:
            extra = (max_new_tokens if max_new_tokens is not None else 1024) if use_cache else 0
            cache = StaticKVCache(self.n_layers, B, self.heads, max(64, -(-(L + extra) // 64) * 64), self._dev())
        if cache.B != B or cache.heads != self.heads or cache.layers != self.n_layers:
            raise ValueError(f"KV cache built for batch {cache.B} x {cache.heads} heads x {cache.layers} layers, "
                             f"got batch {B} x {self.heads} heads x {self.n_layers} layers")
        cache.reserve(cache.length + L)
        self._ensure_rope(cache.length + L)
        W = self._w
        x = inputs_embeds.float().reshape(B * L, H).contiguous().clone()
        if L == 1 and B <= 32 and cache.length > 0:
            self._decode_layers(x, B, cache)
            cache.length += 1
            cache.pos.fill_(cache.length)
        else:
            self._prefill(x, B, L, cache)
        out = torch.empty(B * L, H, device=x.device)
        ops.rmsnorm(x, W["norm"], B * L, H, out_f32=out)
        return LMOutput(out.reshape(B, L, H), cache if use_cache else None)

    # ------------------------------------------------------------------ conditioning prefix (llm_sft.py:58-78)
    def _adapter(self, feats):
        W = self._prepare()
        B, T, Fd = feats.shape
        fpad = (Fd + 63) // 64 * 64
        a = Planes.zeros((B * T, fpad), True, feats.device)
        ops.rows_to_planes(feats.float().contiguous(), 1, B * T, Fd, a, fpad, B * T, 0)
        wpad = W.get("adapter_pad")
        if wpad is None:
            w = torch.zeros(self.hidden, fpad, device=feats.device)
            w[:, :Fd] = self.adapter.weight.detach().float()
            wpad = W["adapter_pad"] = Planes.from_f32(w, True)
        out = torch.empty(B * T, self.hidden, device=feats.device)
        ops.gemm(a, wpad, self.hidden, a_batch=1, a_rows_per_batch=B * T, a_ld=fpad, m_per_batch=B * T, bias=W["adapter_b"],
                 out_f32=rowmap(out, self.hidden, B * T, 0))
        return out.reshape(B, T, self.hidden)

    def _prefix(self, task_name, enroll_feats, mix_feats):
        B = mix_feats.shape[0]
        e = lambda w: w.detach().float()
        task = e(self.task_embedding.weight)[self.task_map[task_name]][None, None].expand(B, 1, -1)
        mix_sos = e(self.mix_sos_embedding.weight)[0][None, None].expand(B, 1, -1)
        parts = [task]
        if enroll_feats is not None:
            parts += [e(self.enroll_sos_embedding.weight)[0][None, None].expand(B, 1, -1), self._adapter(enroll_feats)]
        parts += [mix_sos, self._adapter(mix_feats)]
        return torch.cat(parts, 1)

    # ------------------------------------------------------------------ teacher-forced forward (llm_sft.py:37-89)
    @torch.no_grad()
    def forward(self, task_name, enroll_mel, enroll_feats, mix_mel, mix_feats, global_ids, semantic_ids, return_logits=False):
        W = self._prepare()
        g = global_ids.long() + self.global_offset
        s = semantic_ids.long() + self.semantic_offset
        B = g.shape[0]
        col = lambda v: torch.full((B, 1), v, dtype=torch.long, device=g.device)
        input_ids = torch.cat([col(0), g, col(1), s], 1)
        target_ids = torch.cat([g, col(1), s, col(2)], 1)
        emb = torch.cat([self._prefix(task_name, enroll_feats if enroll_mel is not None else None, mix_feats),
                         W["emb"][input_ids]], 1)
        hs = self.llm_forward(emb).last_hidden_state[:, -target_ids.shape[1]:].contiguous()
        Lt, V = hs.shape[1], self.vocab_size
        M = B * Lt
        hp = Planes.zeros((M, self.hidden), True, hs.device)
        ops.split_f16(hs.reshape(M, self.hidden), hp)
        vpad = (V + 3) // 4 * 4
        logits = torch.empty(M, vpad, device=hs.device)
        ops.gemm(hp, W["head"], V, a_batch=1, a_rows_per_batch=M, a_ld=self.hidden, m_per_batch=M,
                 out_f32=rowmap(logits, vpad, M, 0))
        # label-smoothed KL + accuracy (llm.py:87-104) in one pass over the produced logits (csrc/llm.cu lm_loss_*)
        la = ops.lm_loss(logits, vpad, M, V, target_ids.reshape(-1).contiguous(), self.label_smoothing)
        loss, acc = la[0], la[1]
        if return_logits:
            logits = logits[:, :V].reshape(B, Lt, V)
        return (loss, acc, logits) if return_logits else (loss, acc)

    # ------------------------------------------------------------------ generate (llm_sft.py:93-195)
    @torch.no_grad()
    def generate(self, task_name, enroll_mel, enroll_feats, mix_mel, mix_feats, global_length: int = 32,
                 temperature: float = 0.8, top_k: int = 50, top_p: float = 0.95, do_sample: bool = True,
                 use_cuda_graph: bool = True, seed: Optional[int] = None):
        """llm_sft.py:93-195 with the reference's signature and defaults.  do_sample=True draws every token on the device
        (top-k -> top-p -> temperature -> multinomial, csrc/llm.cu lm_sample_embed_kernel); `seed` (default: torch's
        global generator) makes the draw reproducible.  Greedy decoding ignores temperature / top_k / top_p exactly as the
        reference's arg-max does (filters never remove the arg-max, llm.py:263-287)."""
        sampling = None
        if do_sample:
            if self.decode_kernel != "tc":
                raise NotImplementedError("sampled decoding runs on the tensor-core decode path only (QB_LM_DECODE=tc)")
            if not (0.0 < temperature <= 1.0):
                raise AssertionError("0 < temperature <= 1.0 (llm.py:278)")
            if top_k <= 0 or top_k > 1024:
                raise NotImplementedError("native sampling supports 1 <= top_k <= 1024 (reference default 50)")
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            sampling = dict(temperature=float(temperature), top_k=int(top_k), top_p=float(top_p), seed=int(seed))
        semantic_length = mix_mel.size(1)
        Ball = mix_feats.shape[0]
        starts = list(range(0, Ball, self.chunk))         # decode kernels keep <= 32 sequences' rows in registers
        n_lanes = min(self.lanes, len(starts))
        outs = []
        if n_lanes <= 1:
            for ci, b0 in enumerate(starts):
                sl = slice(b0, min(b0 + self.chunk, Ball))
                if sampling is not None:
                    sampling["call"] = ci
                outs.append(self._generate_chunk(task_name, None if enroll_mel is None else enroll_feats[sl], mix_feats[sl],
                                                 semantic_length, global_length, use_cuda_graph, sampling))
        else:
            # chunk ci runs on lane ci % n_lanes: the host enqueues one chunk after the other, the device overlaps the lanes
            n_pos = 2 + mix_feats.shape[1] + (0 if enroll_mel is None else 1 + enroll_feats.shape[1]) + global_length + 1 + semantic_length
            views = self._lanes(n_lanes, n_pos)
            cur = torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(cur)
            for ci, b0 in enumerate(starts):
                sl = slice(b0, min(b0 + self.chunk, Ball))
                view, stream = views[ci % n_lanes]
                if sampling is not None:
                    sampling["call"] = ci
                if ci < n_lanes:
                    stream.wait_event(ready)             # the inputs were produced on the caller's stream
                with torch.cuda.stream(stream):
                    outs.append(view._generate_chunk(task_name, None if enroll_mel is None else enroll_feats[sl], mix_feats[sl],
                                                     semantic_length, global_length, use_cuda_graph, sampling))
            for _, stream in views[:n_lanes]:
                cur.wait_stream(stream)
            for gi, si in outs:                           # allocated on a lane's stream, consumed on the caller's
                gi.record_stream(cur)
                si.record_stream(cur)
        return torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0)

    def _lanes(self, n: int, n_positions: int):
        """Lane = (shallow view of this module with its own workspace, decode state and captured graphs; its own stream).  The
        prepared weights and RoPE tables are shared read-only, so the tables are sized BEFORE the lanes exist (growing them
        replaces the tensors captured graphs point at)."""
        self._prepare()
        self._ensure_rope(n_positions)
        rows = self._w["rope_rows"]
        if self._lane_views is None or self._lane_views[0] != rows:
            self._lane_views = (rows, [])
        lanes = self._lane_views[1]
        while len(lanes) < n:
            v = copy.copy(self)
            v._ws, v._gen_state, v._lane_views = {}, {}, None
            v.lanes = 1
            v.att_unroll = self.lane_att_unroll
            lanes.append((v, torch.cuda.Stream(device=self._dev())))
        return lanes

    def _generate_chunk(self, task_name, enroll_feats, mix_feats, semantic_length, global_length, use_graph, sampling=None):
        W = self._prepare()
        dev = mix_feats.device
        prefix = self._prefix(task_name, enroll_feats, mix_feats)
        B, P, H = prefix.shape
        n_steps = global_length + 1 + semantic_length
        Lmax = -(-(P + n_steps) // 64) * 64
        self._ensure_rope(P + n_steps)
        W = self._w
        max_cols = max(self.global_size, self.semantic_size)
        samp_key = None if sampling is None else (sampling["temperature"], sampling["top_k"], sampling["top_p"])
        # Decode state (KV cache, counters, output ids) and the captured graphs are kept per shape: capturing and
        # instantiating ~560 kernel nodes costs the host 10-50 ms, as much as the whole generation takes on the device.
        key = (B, P, n_steps, bool(use_graph), self.decode_kernel, int(self.graph_steps), str(dev), samp_key)
        st = self._gen_state.get(key)
        if st is None:
            self._gen_state.clear()                    # one shape at a time (the cache is ~0.9 GB at B=32)
            st = dict(cache=StaticKVCache(self.n_layers, B, self.heads, Lmax, dev),
                      xs=torch.zeros(B, H, device=dev),
                      rng=torch.zeros(2, dtype=torch.int32, device=dev), slot=torch.zeros(2, dtype=torch.int32, device=dev),
                      out_ids=torch.zeros(B, n_steps, dtype=torch.int64, device=dev),
                      pv=torch.zeros(max_cols // 16 + 1, 32, device=dev),
                      pi=torch.zeros(max_cols // 16 + 1, 32, dtype=torch.int32, device=dev), g1=None, gk=None, captured=False,
                      logits=torch.zeros(B, max_cols, device=dev) if sampling is not None else None,
                      seed=torch.zeros(4, dtype=torch.int32, device=dev), dbg=torch.zeros(B, 4, device=dev))
            self._gen_state[key] = st
        cache, xs, rng, slot, out_ids, pv, pi = (st[k] for k in ("cache", "xs", "rng", "slot", "out_ids", "pv", "pi"))
        cache.length = 0
        cache.pos.zero_()
        slot.zero_()
        if sampling is not None:        # Philox key / call counter live in device memory: the captured graph is reused across seeds
            sd_ = sampling["seed"]
            to_i32 = lambda v: v - (1 << 32) if v >= (1 << 31) else v
            st["seed"].copy_(torch.tensor([to_i32(sd_ & 0xFFFFFFFF), to_i32((sd_ >> 32) & 0xFFFFFFFF), sampling["call"], 0],
                                          dtype=torch.int32), non_blocking=True)
        rng.copy_(torch.tensor([self.global_offset, self.global_offset + self.global_size], dtype=torch.int32), non_blocking=True)
        x = prefix.reshape(B * P, H).contiguous().clone()
        self._prefill(x, B, P, cache)

        def step():
            self._decode_layers(xs, B, cache)
            if sampling is not None:
                ops.lm_head_sample_tc(xs, B, H, W["head_p"], rng, max_cols, W["emb"], xs, out_ids, n_steps, cache.pos, slot, pv, pi,
                                      st["logits"], sampling["temperature"], sampling["top_k"], sampling["top_p"], st["seed"],
                                      st["dbg"])
            elif self.decode_kernel == "tc":
                ops.lm_head_argmax_tc(xs, B, H, W["head_p"], rng, max_cols, W["emb"], xs, out_ids, n_steps, cache.pos, slot,
                                      pv, pi)
            else:
                ops.lm_head_argmax(xs, B, H, W["norm"], W["head32"], rng, max_cols, W["emb"], xs, out_ids, n_steps, cache.pos,
                                   slot, pv, pi)

        # All decode state is on the device, so a graph may hold any number of consecutive steps: one single-step graph
        # plus one of `graph_steps` steps (fewer replays per generation).
        K = max(1, int(self.graph_steps))
        if use_graph and not st["captured"] and not (self.decode_kernel == "persistent" and sampling is None):
            # warm-up outside capture (one-time cudaFuncSetAttribute calls), then restore the mutated state
            xs.copy_(W["emb"][self.global_sos_token_id][None].expand(B, H))
            step()
            torch.cuda.synchronize()
            st["g1"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st["g1"]):
                step()
            if K > 1 and n_steps >= K:
                st["gk"] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(st["gk"]):
                    for _ in range(K):
                        step()
            st["captured"] = True
            cache.pos.fill_(cache.length)
            slot.zero_()
        g1, gk = (st["g1"], st["gk"]) if use_graph else (None, None)
        persistent = (self.decode_kernel == "persistent" and sampling is None and H == 512 and self.n_layers <= 16 and max_cols % 16 == 0)
        if persistent and "ptrs" not in st:
            Ls = W["layers"]
            st["ptrs"] = dict(n=self.n_layers, wqkv=ops.ptr_array([l["wqkv_p"] for l in Ls]), wo=ops.ptr_array([l["wo_p"] for l in Ls]),
                              wg=ops.ptr_array([l["wg_p"] for l in Ls]), wu=ops.ptr_array([l["wu_p"] for l in Ls]),
                              wd=ops.ptr_array([l["wd_p"] for l in Ls]), k=ops.ptr_array(cache.k), v=ops.ptr_array(cache.v))
            st["bar"] = torch.zeros(4, dtype=torch.int32, device=dev)
            st["dbuf"] = (torch.zeros(B, H, device=dev), torch.zeros(B, H, device=dev), torch.zeros(B, 4 * H, device=dev))

        def run_persistent(n):
            qb, ab, mb = st["dbuf"]
            ops.lm_decode_steps(xs, B, H, self.heads, 4 * H, st["ptrs"], cache.Lmax, W["head_p"], rng, max_cols, W["emb"], W["cos"], W["sin"],
                                qb, ab, mb, pv, pi, out_ids, n_steps, cache.pos, slot, n, st["bar"])

        def run(n):
            if persistent:
                return run_persistent(n)
            while n > 0:
                if gk is not None and n >= K:
                    gk.replay()
                    n -= K
                else:
                    g1.replay() if g1 is not None else step()
                    n -= 1

        xs.copy_(W["emb"][self.global_sos_token_id][None].expand(B, H))
        run(global_length + 1)
        rng.copy_(torch.tensor([self.semantic_offset, self.semantic_offset + self.semantic_size], dtype=torch.int32))
        xs.copy_(W["emb"][self.semantic_sos_token_id][None].expand(B, H))
        run(semantic_length)
        cache.length += n_steps
        global_ids = out_ids[:, :global_length] - self.global_offset          # (new tensors: out_ids is reused by the next call)
        semantic_ids = out_ids[:, global_length + 1:] - self.semantic_offset
        return global_ids, semantic_ids
