#!/usr/bin/env python
"""bench.py - H-Codec-2.0 encode + RVQ + decode throughput on B200 (BASELINE.json configs[1]).

A "step" = one pass of the hot path (Codec.encode -> Codec.decode) over one synthetic batch of
B=64 clips x 10 s at the shipped 48 kHz configuration (480 000 samples / clip, 125 tokens / stream),
seeded random weights of the shipped architecture (large_12.5hz_config.yaml).  No pretrained
weights / datasets exist offline, hence `data: synthetic`.

  python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (CUDA kernels via the C ABI)
  python bench.py --impl reference --steps 2 --warmup 1     # the reference's CPU path (oracle port)
  torchrun ... bench.py --gpus N ...                        # one rank per GPU, weak scaling

Prints ONE JSON line (see README / DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "hcodec2_encode_rvq_decode_samples_per_s"
UNIT = "samples/s"
# algorithmic FLOPs per 50 Hz frame (multiply-add = 2), encoder+semantic+RVQ+decoder
FLOP_PER_FRAME = 2245.8e6

H2_FULL = dict(
    sampling_rate=48000,
    encoder_config=dict(dim=1536, intermediate_dim=4608, dimension=512, n_fft=1920, hop_length=960,
                        convnext_layers=24, transformer_layers=2, target_frame_rate=12.5, causal=False),
    decoder_config=dict(input_channels=1024, dim=1536, intermediate_dim=4608, convnext_layers=32, n_fft=1920,
                        hop_length=960, transformer_layers=2, target_frame_rate=12.5, causal=False),
    quantizer_config=dict(dim=512, codebook_size=1024, num_quantizers=16, decay=0.99, kmeans_init=True,
                          kmeans_iters=50, quantize_dropout=False),
    semantic_encoder_config=dict(input_channels=768, encode_channels=1536, out_channels=512,
                                 channel_ratios=[1, 1, 1], strides=[2, 1, 2]),
    semantic_decoder_config=dict(code_dim=512, output_channels=768, decode_channels=1536,
                                 channel_ratios=[1, 1, 1], strides=[2, 1, 2]),
)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


def random_init_(model, seed: int):
    """Seeded random weights of the shipped architecture, generated on the module's device."""
    dev = next(model.parameters()).device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "rnn." in name:
                H = p.shape[-1] if p.dim() == 2 else p.shape[0] // 4
                p.copy_((torch.rand(p.shape, generator=g, device=dev) * 2 - 1) / H ** 0.5)
            elif name.endswith("gamma"):
                n_layers = 24 if name.startswith("encoder.") else 32
                p.copy_((1.0 / n_layers) * (1 + 0.2 * torch.randn(p.shape, generator=g, device=dev)))
            elif p.dim() >= 2:
                fan = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * fan ** -0.5)
            elif "norm" in name and name.endswith("weight") or name.endswith("prior_net.7.weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g, device=dev))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g, device=dev))
        for q in (model.quantizer, model.semantic_quantizer):
            cb = torch.stack([torch.randn(q.codebook_size, q.dim, generator=g, device=dev) * 0.35 * 0.85 ** i
                              for i in range(q.num_quantizers)], 0)
            q.set_codebooks(cb)
    model._w = None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


_BEST_THREADS = None


def _best_threads(sd, cfg):
    """Pick the host thread count that makes the reference's CPU path fastest (short calibration on a
    0.8 s clip): on a many-core shared host `all cores` is often far slower than a moderate count."""
    global _BEST_THREADS
    if _BEST_THREADS is None:
        from oracle import hcodec2
        try:
            avail = len(os.sched_getaffinity(0))
        except Exception:
            avail = os.cpu_count() or 1
        g = torch.Generator().manual_seed(3)
        wav = 0.1 * torch.randn(1, 38400, generator=g)
        feat = torch.randn(1, 768, 40, generator=g)
        best = (1e30, 1)
        for n in sorted({min(avail, c) for c in (8, 16, 32, 64, avail)}):
            torch.set_num_threads(n)
            hcodec2.codec_encode(sd, cfg, wav, feat)   # warm
            t0 = time.perf_counter()
            hcodec2.codec_encode(sd, cfg, wav, feat)
            dt = time.perf_counter() - t0
            if dt < best[0]:
                best = (dt, n)
        _BEST_THREADS = best[1]
    return _BEST_THREADS


def cpu_baseline(model_sd_cpu, cfg, seconds, clips):
    """The reference's own PyTorch CPU path (oracle port, pinned bit-exact against the reference
    modules) on this box's host cores, on a bounded sample of the same workload."""
    from oracle import hcodec2
    torch.set_num_threads(_best_threads(model_sd_cpu, cfg))
    T = int(seconds * cfg["sampling_rate"])
    g = torch.Generator().manual_seed(7)
    wav = 0.1 * torch.randn(clips, T, generator=g)
    f = torch.randn(clips, 768, T // 960, generator=g)
    feat = torch.sign(f) * f.abs() ** 0.3
    t0 = time.perf_counter()
    ac, sc = hcodec2.codec_encode(model_sd_cpu, cfg, wav, feat)
    hcodec2.codec_decode(model_sd_cpu, cfg, ac, sc)
    dt = time.perf_counter() - t0
    return clips * T / dt, dt


def run_reference(args, cfg):
    """--impl reference: the CPU path alone, same metric / config (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import weights
    torch.manual_seed(0)
    sd = weights.make_h2_state_dict(cfg, 0)
    clips = args.ref_clips
    times = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_baseline(sd, cfg, args.seconds, clips)
        if i >= args.warmup:
            times.append(dt)
    T = int(args.seconds * cfg["sampling_rate"])
    ms = 1e3 * sum(times) / len(times)
    value = clips * T / (ms / 1e3)
    cores = _BEST_THREADS
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=f"HCodec-2.0 (48 kHz shipped config) {args.seconds:g} s clips, encode+RVQ+decode",
                            batch_per_step=clips, samples_per_clip=T, cpu_threads=cores),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{clips} clip(s) x {args.seconds:g} s per step, oracle port of the reference "
                                         f"(pinned bit-exact against the reference modules), torch CPU fp32"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def run_lm(args):
    """Secondary line: UniSE SR AR-LM greedy generate (prefix 252 + 33 + 250 cached steps), B=32 per GPU.
    tokens/s counts generated tokens (283 per sequence)."""
    LM = dict(num_tasks=3, task_map=dict(se=0, tse=1, rtse=2), feats_dim=768,
              llm_base_config=dict(cond_dim=80, global_size=4096, semantic_size=8192, hidden_size=512, num_layers=12,
                                   num_attention_heads=8, dropout_p=0.1, max_position_embeddings=4096, label_smoothing=0.1))
    B, T = 32, 250
    if args.impl == "reference" and int(os.environ.get("RANK", "0")) != 0:
        return
    if args.impl == "reference":
        from oracle import llama
        sd = llama.make_lm_state_dict(LM, 7, 2.0)
        Bc = 4
        mix = torch.randn(Bc, T, 768)
        torch.set_num_threads(min(32, os.cpu_count() or 1))
        t0 = time.perf_counter()
        llama.sft_generate(sd, LM, "se", None, mix, T)
        dt = time.perf_counter() - t0
        v = Bc * 283 / dt
        print(json.dumps(dict(metric="unise_sr_arlm_generate_tokens_per_s", value=v, unit="tokens/s", n_gpus=args.gpus,
                              steps=1, warmup=0, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                              dtype="f32", data="synthetic", impl="reference",
                              config=dict(workload="UniSE SR AR-LM greedy generate, prefix 252 + 283 cached steps", batch=Bc),
                              cpu_baseline=dict(value=v, unit="tokens/s", cores=torch.get_num_threads(), kind="port",
                                                sample=f"{Bc} sequences x 283 tokens, oracle port pinned against transformers.LlamaModel"),
                              e2e=dict(value=v, unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)))
        return
    from unified_audio_b200 import ops
    from unified_audio_b200.llm import LLM_SFT
    from unified_audio_b200.parallel import gather_tokens
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:       # sequences are independent: B per rank (weak scaling), one all_gather of the generated ids
        import torch.distributed as dist_
        dist = dist_
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    m = LLM_SFT(num_tasks=3, task_map=LM["task_map"], feats_dim=768, llm_base_config=LM["llm_base_config"]).to(dev)
    g = torch.Generator(device=dev).manual_seed(7)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() >= 2 and "embedding" not in n:
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * (2.0 / p.shape[-1] ** 0.5))
            elif p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g, device=dev))
            else:
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g, device=dev))
    m._w = None
    mix_h = torch.randn(B, T, 768, generator=torch.Generator().manual_seed(100 + rank)).pin_memory()
    mix = mix_h.to(dev)

    def step(src):
        gi, si = m.generate("se", None, None, src, src, do_sample=False)
        if dist is not None:
            gather_tokens(torch.cat([gi, si], 1), world * B)
        return gi, si


    def timed(fn):
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(args.steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    if args.workload == "lm_forward":
        # teacher-forced UniSE LM forward (llm_sft.py:37-89): prefix 252 + 284 code tokens per sequence, logits over the
        # 12291-entry vocabulary, loss + accuracy - the "AR-LM forward" of north_star.  Tensor-bound; every GEMM is a 3-term split.
        gt = torch.Generator().manual_seed(7 + rank)
        gids_h = torch.randint(0, 4096, (B, 32), generator=gt).pin_memory()
        sids_h = torch.randint(0, 8192, (B, T), generator=gt).pin_memory()
        gids, sids = gids_h.to(dev), sids_h.to(dev)
        L = 2 + T + 32 + 1 + T + 1            # task + mix_sos + feats, then sos/global/sos/semantic (+ eos target)
        fwd = lambda a, b_, c_: m("se", None, None, a, a, b_, c_)
        for _ in range(max(args.warmup, 3)):
            fwd(mix, gids, sids)
        ops.launch_count_reset()
        ms = timed(lambda: fwd(mix, gids, sids))
        launches = ops.launch_count()

        def e2e_fwd():
            loss, acc = fwd(mix_h.to(dev, non_blocking=True), gids_h.to(dev, non_blocking=True), sids_h.to(dev, non_blocking=True))
            loss.cpu(); acc.cpu()
        ms_e2e = timed(e2e_fwd)
        if rank == 0:
            peaks = load_peaks()
            per_tok = 12 * 2 * (4 * 512 * 512 + 3 * 512 * 2048)
            flops = world * B * (L * per_tok + 12 * 4 * 512 * L * L / 2 + 2 * 768 * 512 * T + (T + 34) * 2 * 512 * 12291)
            tf = flops / (ms * 1e-3) / 1e12
            print(json.dumps(dict(
                metric="unise_lm_forward_tokens_per_s", value=world * B * L / (ms * 1e-3), unit="tokens/s", n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f16x3 split tensor-core (fp32-grade), f32 accumulate", data="synthetic",
                config=dict(workload=f"UniSE LM teacher-forced forward, {B} sequences x {L} positions per GPU, logits + loss", batch=B * world,
                            positions=L, parallelism=f"dp{world}"),
                e2e=dict(value=world * B * L / (ms_e2e * 1e-3), unit="tokens/s", ms_per_step=ms_e2e,
                         h2d_bytes_per_step=int(mix_h.numel() * 4 + gids_h.numel() * 8 + sids_h.numel() * 8) * world, d2h_bytes_per_step=8 * world),
                gpu_launches=int(launches),
                roofline=dict(bound="tensor", achieved=tf, peak=peaks["tf_sus"] * world, unit="TFLOP/s", frac=tf / (peaks["tf_sus"] * world),
                              traffic=None, kernel="whole forward, algorithmic FLOPs (every GEMM and the attention issued 3x: ceiling 1/3)"))))
        if dist is not None:
            dist.destroy_process_group()
        return

    for _ in range(max(args.warmup, 1)):
        step(mix)
    ops.launch_count_reset()
    ms = timed(lambda: step(mix))
    launches = ops.launch_count() // args.steps

    def e2e_step():
        gi, si = step(mix_h.to(dev, non_blocking=True))
        gi.cpu(); si.cpu()
    ms_e2e = timed(e2e_step)
    B = B * world                       # whole-job totals from here on
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    # bytes per decode step: fp32 layer weights + head slice + fp32 KV read, summed over the 283 steps
    w_bytes = 12 * (4 * 512 * 512 + 3 * 512 * 2048) * 4
    kv = sum(B * 2 * 12 * 512 * 4 * (252 + i + 1) for i in range(283))
    head = 33 * 4096 * 512 * 4 + 250 * 8192 * 512 * 4
    total_bytes = world * (283 * w_bytes + head) + kv      # every rank streams its replica of the weights
    gbs = total_bytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(metric="unise_sr_arlm_generate_tokens_per_s", value=B * 283 / (ms * 1e-3), unit="tokens/s", n_gpus=world,
                          steps=args.steps, warmup=max(args.warmup, 1), ms_per_step=ms, higher_is_better=True, scaling="weak",
                          vs_baseline=None, dtype="f16x3 split tensor-core (fp32-grade), f32 accumulate / f32 KV cache", data="synthetic",
                          config=dict(workload="UniSE SR AR-LM greedy generate (prefill 252 + 33 + 250 cached steps), batch=32 per GPU",
                                      batch=B, semantic_length=T, parallelism=f"dp{world} (sequences sharded, one NCCL all_gather of ids)",
                                      launches="decode steps replay one captured CUDA graph (62 kernels); gpu_launches counts eager launches + the capture"),
                          e2e=dict(value=B * 283 / (ms_e2e * 1e-3), unit="tokens/s", h2d_bytes_per_step=int(mix_h.numel() * 4) * world,
                                   d2h_bytes_per_step=B * 282 * 8),
                          gpu_launches=int(launches),
                          roofline=dict(bound="hbm", achieved=gbs, peak=peaks["hbm"] * world, unit="GB/s", frac=gbs / (peaks["hbm"] * world),
                                        traffic=None, kernel="decode step (lm_skinny<QKV|RESID|GATEUP|HEAD> + lm_decode_attn2), algorithmic bytes = 4 B/param packed "
                                        "weights + head slice + fp32 KV read per step, whole generate incl. prefill"))))
    if dist is not None:
        dist.destroy_process_group()


def bicodec_flops_per_clip(cfg, T):
    """algorithmic FLOPs of BiCodec.detokenize for one clip of T tokens (2 x MACs, true channel counts and taps)"""
    p, d = cfg["prenet"], cfg["decoder"]
    dim, inter = p["vocos_dim"], p["vocos_intermediate_dim"]
    blocks = 2 * len(p["sample_ratios"]) + p["vocos_num_layers"]
    backbones = len(p["sample_ratios"]) + 1
    f = 2.0 * T * (p["input_channels"] * dim + backbones * dim * dim * 7 + blocks * 2 * dim * inter + dim * p["out_channels"])
    ch = d["channels"]
    f += 2.0 * T * d["input_channel"] * ch * 7
    Tc = T
    for i, (k, r) in enumerate(zip(d["kernel_sizes"], d["rates"])):
        cin, cout = ch // 2 ** i, ch // 2 ** (i + 1)
        f += 2.0 * Tc * cin * cout * k                     # transposed conv: k taps per INPUT frame
        Tc *= r
        f += 3 * 2.0 * Tc * cout * cout * 8                # 3 residual units: dilated k7 + 1x1
    f += 2.0 * Tc * (ch // 2 ** len(d["rates"])) * 7
    return f


def run_bicodec(args):
    """Secondary line: BiCodec.detokenize, the decoder UniSE feeds its AR-LM tokens to (configs[2] back half):
    B=32 clips x 250 semantic tokens + 32 global tokens -> 5 s @ 16 kHz each."""
    B, T = 32, 250
    if args.impl == "reference" and int(os.environ.get("RANK", "0")) != 0:
        return
    if args.impl == "reference":
        from oracle import bicodec as ob
        cfg = ob.BICODEC_FULL
        sd = ob.make_state_dict(cfg, 5)
        sem, glob = ob.synth_tokens(cfg, 1, T, 9)
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        ob.detokenize(sd, cfg, sem[:, :25], glob)
        t0 = time.perf_counter()
        ob.detokenize(sd, cfg, sem, glob)
        dt = time.perf_counter() - t0
        v = T * 320 / dt
        print(json.dumps(dict(metric="bicodec_detokenize_samples_per_s", value=v, unit="samples/s", n_gpus=args.gpus, steps=1,
                              warmup=1, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                              dtype="f32", data="synthetic", impl="reference",
                              config=dict(workload="BiCodec detokenize, 1 clip x 250 tokens -> 5 s @ 16 kHz", batch=1),
                              cpu_baseline=dict(value=v, unit="samples/s", cores=torch.get_num_threads(), kind="port",
                                                sample="1 clip x 5 s, oracle port pinned against the reference's BiCodec classes"),
                              e2e=dict(value=v, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)))
        return
    from unified_audio_b200 import ops
    from unified_audio_b200.bicodec import BICODEC_CONFIG, BiCodec
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:       # clips are independent: B per rank, no data-path collective (the waveforms stay on their rank)
        import torch.distributed as dist_
        dist = dist_
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    cfg = BICODEC_CONFIG
    m = BiCodec(cfg).to(dev)
    g = torch.Generator(device=dev).manual_seed(5)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("alpha"):
                p.copy_(1 + 0.3 * torch.rand(p.shape, generator=g, device=dev))
            elif n.endswith(("weight_g",)):
                p.fill_(1.0)
            elif p.dim() >= 2:
                fan = p[0].numel() if "block.1.weight_v" not in n else 2 * p.shape[0]
                p.copy_(torch.randn(p.shape, generator=g, device=dev) / fan ** 0.5)
            elif n.endswith(("gamma",)):
                p.fill_(1.0 / 12)
            elif n.endswith(("norm.weight", "scale.bias", "final_layer_norm.weight")):
                p.fill_(1.0)
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=dev))
        sdm = m.state_dict()
        for n in list(sdm):     # weight-norm gains ~ ||v||, residual branches damped (as oracle.bicodec.make_state_dict)
            if n.endswith("weight_g"):
                v = sdm[n[:-1] + "v"]
                gain = 0.3 if ".block.3." in n else (0.1 if n.startswith("decoder.model.6.") else 1.0)
                sdm[n].copy_(gain * v.reshape(v.shape[0], -1).norm(dim=1).reshape(sdm[n].shape))
    m._w = None
    gt = torch.Generator().manual_seed(50 + rank)
    sem_h = torch.randint(0, cfg["quantizer"]["codebook_size"], (B, T), generator=gt).pin_memory()
    glob_h = torch.randint(0, 4096, (B, 1, 32), generator=gt).pin_memory()
    sem, glob = sem_h.to(dev), glob_h.to(dev)
    wav_h = torch.empty(B, 1, T * 320).pin_memory()

    def timed(fn):
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(args.steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for _ in range(max(args.warmup, 3)):
        m.detokenize(sem, glob)
    if args.quick:
        ms = timed(lambda: m.detokenize(sem, glob))
        if rank == 0:
            print(json.dumps(dict(quick=True, ms_per_step=ms)))
        return
    ops.launch_count_reset()
    ms = timed(lambda: m.detokenize(sem, glob))
    launches = ops.launch_count()

    def e2e_step():
        w = m.detokenize(sem_h.to(dev, non_blocking=True), glob_h.to(dev, non_blocking=True))
        wav_h.copy_(w, non_blocking=True)
    ms_e2e = timed(e2e_step)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    samples = world * B * T * 320
    tf = world * B * bicodec_flops_per_clip(cfg, T) / (ms * 1e-3) / 1e12
    cpu = None
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "bicodec", "--impl", "reference"],
                           capture_output=True, text=True, timeout=600)
        cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as e:      # the baseline is reported, never required for the GPU line
        cpu = dict(value=None, unit="samples/s", cores=0, kind="port", sample=f"failed: {e}")
    print(json.dumps(dict(
        metric="bicodec_detokenize_samples_per_s", value=samples / (ms * 1e-3), unit="samples/s", n_gpus=world, steps=args.steps,
        warmup=max(args.warmup, 3), ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f16x3 split tensor-core (fp32-grade), f32 accumulate", data="synthetic",
        config=dict(workload="BiCodec detokenize (UniSE's decoder): 32 clips x 250 semantic + 32 global tokens -> 5 s @ 16 kHz",
                    batch_per_gpu=B, tokens=T, precision_policy="accurate",
                    l2="activations per step (~10 GB) exceed the 126 MB L2; no flush needed",
                    parallelism=f"dp{world} (clips sharded, no collective)"),
        e2e=dict(value=samples / (ms_e2e * 1e-3), unit="samples/s", ms_per_step=ms_e2e,
                 h2d_bytes_per_step=int(sem_h.numel() * 8 + glob_h.numel() * 8) * world,
                 d2h_bytes_per_step=int(wav_h.numel() * 4) * world),
        gpu_launches=int(launches),
        roofline=dict(bound="tensor", achieved=tf, peak=peaks["tf_sus"] * world, unit="TFLOP/s", frac=tf / (peaks["tf_sus"] * world),
                      traffic=None, kernel="whole detokenize path, algorithmic FLOPs (true channels / taps; every GEMM issued 3x in the "
                      "`accurate` policy) against the sustained bf16 peak"),
        cpu_baseline=cpu)))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying the CUDA graph")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--precision", default="mixed")
    ap.add_argument("--ref-clips", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="codec", choices=["codec", "lm", "lm_forward", "bicodec"],
                    help="codec = BASELINE configs[1] (default, the driver's line); lm = UniSE SR AR-LM generate (configs[2]); "
                         "bicodec = BiCodec detokenize, the decoder UniSE feeds the LM tokens to")
    ap.add_argument("--quick", action="store_true", help="profiling aid: W warm-up + K steps only, no e2e/roofline/cpu legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    cfg = H2_FULL
    if args.workload in ("lm", "lm_forward"):
        return run_lm(args)
    if args.workload == "bicodec":
        return run_bicodec(args)
    if args.impl == "reference":
        return run_reference(args, cfg)

    from unified_audio_b200 import ops
    from unified_audio_b200.codec import Codec
    from unified_audio_b200.parallel import gather_tokens

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    peaks = load_peaks()

    model = Codec(cfg["encoder_config"], cfg["decoder_config"], cfg["quantizer_config"],
                  cfg["semantic_encoder_config"], cfg["semantic_decoder_config"], precision=args.precision).to(dev)
    random_init_(model, 1234)
    B, T = args.batch, int(args.seconds * cfg["sampling_rate"])
    T -= T % 3840
    F_ = T // 960
    g = torch.Generator().manual_seed(2000 + rank)
    wav_h = (0.1 * torch.randn(B, T, generator=g)).pin_memory()
    f = torch.randn(B, 768, F_, generator=g)
    feat_h = (torch.sign(f) * f.abs() ** 0.3).pin_memory()
    wav_d, feat_d = wav_h.to(dev), feat_h.to(dev)

    # the public fixed-shape entry point: encode -> decode captured once in a CUDA graph (Codec.graphed), replayed per step
    graphed = None
    if not args.no_graph:
        try:
            graphed = model.graphed("roundtrip", wav_d, feat_d)
        except Exception as e:      # same kernels either way: fall back to launching them one by one
            print(f"[bench] CUDA-graph capture failed ({e!r}); launching kernel by kernel", file=sys.stderr)
            torch.cuda.synchronize()

    def step_device():
        if graphed is not None:
            ac, sc, rec = graphed()                      # static inputs already hold this rank's batch (HBM-resident)
        else:
            ac, sc = model.encode(wav_d, feat_d)
            rec = model.decode(ac, sc)
        if dist is not None:   # the path's single exchange: gather the int64 tokens
            gather_tokens(torch.stack([ac, sc], 1), world * B)
        return ac, sc, rec

    codes_h = torch.empty(2, B, 16, T // 3840, dtype=torch.int64).pin_memory()
    rec_h = torch.empty(B, T).pin_memory()

    def step_e2e():
        if graphed is not None:
            ac, sc, rec = graphed(wav_h, feat_h)         # pinned host -> static device inputs (H2D inside the timed region)
        else:
            w = wav_h.to(dev, non_blocking=True)
            ft = feat_h.to(dev, non_blocking=True)
            ac, sc = model.encode(w, ft)
            rec = model.decode(ac, sc)
        codes_h[0].copy_(ac, non_blocking=True)
        codes_h[1].copy_(sc, non_blocking=True)
        rec_h.copy_(rec, non_blocking=True)
        return rec

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        ev[0].record()
        for i in range(steps):
            fn()
            ev[i + 1].record()
        barrier()
        total_ms = ev[0].elapsed_time(ev[-1])
        if dist is not None:
            t = torch.tensor([total_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t)
        return total_ms / steps

    if args.quick:
        for _ in range(args.warmup):
            step_device()
        ms = timed(step_device, args.steps)
        if rank == 0:
            print(json.dumps(dict(quick=True, ms_per_step=ms, value=world * B * T / (ms * 1e-3))))
        return
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.launch_count_reset()
    ms = timed(step_device, args.steps)
    launches = ops.launch_count() + (graphed.launches_per_replay * args.steps if graphed is not None else 0)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel: the ConvNeXt pointwise GEMM (tcgen05), timed alone
    M, C, I = B * F_, 1536, 4608
    blk = model._prepare()["enc"]["convnext"][0]
    t1 = model._planes("cnx_t1", (M, C), model.policy["convnext"])
    hid = model._planes("cnx_hid", (M, I), model.policy["convnext"])
    reps = 10
    for _ in range(3):
        model._linear(t1, blk["w1"], I, M, C, bias=blk["b1"], act=ops.ACT_GELU, out_planes=hid, out_planes_map=(I, M, 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        model._linear(t1, blk["w1"], I, M, C, bias=blk["b1"], act=ops.ACT_GELU, out_planes=hid, out_planes_map=(I, M, 0))
    e1.record()
    torch.cuda.synchronize()
    gemm_ms = e0.elapsed_time(e1) / reps
    gemm_tf = 2.0 * M * I * C / (gemm_ms * 1e-3) / 1e12

    if rank != 0:
        return
    samples = world * B * T
    value = samples / (ms * 1e-3)
    path_tf = world * B * F_ * FLOP_PER_FRAME / (ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16x3/f16 tensor-core, f32 accumulate",
        data="synthetic",
        config=dict(workload="HCodec-2.0 batch=64 x 10 s (48 kHz shipped config, 480000 samples/clip) encode+RVQ+decode",
                    batch_per_gpu=B, samples_per_clip=T, tokens_per_stream=T // 3840, precision_policy=args.precision,
                    l2="working set per step (~3 GB activations + 4.6 GB weights) exceeds the 126 MB L2; no flush needed",
                    parallelism=f"dp{world} (clips sharded, one NCCL all_gather of tokens)",
                    launch="one CUDA graph replay per step (Codec.graphed('roundtrip')); gpu_launches = library kernels in the "
                           "graph x steps" if graphed is not None else "kernel by kernel"),
        e2e=dict(value=samples / (ms_e2e * 1e-3), unit=UNIT, ms_per_step=ms_e2e,
                 h2d_bytes_per_step=int(wav_h.numel() * 4 + feat_h.numel() * 4),
                 d2h_bytes_per_step=int(codes_h.numel() * 8 + rec_h.numel() * 4)),
        gpu_launches=int(launches),
        clocks=clocks,
        roofline=dict(bound="tensor", achieved=gemm_tf, peak=peaks["tf_burst"], unit="TFLOP/s", frac=gemm_tf / peaks["tf_burst"],
                      traffic=traffic, kernel="gemm_tc2_kernel<256,1,6> (cta_group::2) ConvNeXt pwconv1 [32000x4608x1536] fp16 + GELU epilogue, timed alone",
                      peak_source=f"{peaks['src']} dense bf16 burst (fp16 shares the pipe)",
                      path_algorithmic_tflops=path_tf, path_frac_of_sustained=path_tf / peaks["tf_sus"]),
    )
    line["cpu_baseline"] = None                  # timed on rank 0 at N = 1 only (the N > 1 lines carry the key, empty)
    if not args.no_cpu_baseline and world == 1:
        sd_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        v, dt = cpu_baseline(sd_cpu, cfg, args.seconds, 1)
        line["cpu_baseline"] = dict(value=v, unit=UNIT, cores=_BEST_THREADS, kind="port",
                                    sample=f"1 clip x {args.seconds:g} s encode+RVQ+decode ({dt:.1f} s), oracle port of "
                                           "the reference's PyTorch CPU path, same weights")
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
