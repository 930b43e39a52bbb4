"""Pin oracle/llama.py's restatement of `LLM_SFT.forward / generate / sample_logits / loss_function` against the REFERENCE'S OWN CODE.

TEST INFRASTRUCTURE.  Run in the build container only:  python -m oracle.make_golden_lm_reference

SURVEY 7.6: `QuarkAudio-UniSE/model/llm` does not import here as shipped - `conformer.py:17` needs x_transformers (absent), and
`llm.py:79,183-206` are written against the transformers release the reference pins (`LlamaModel._update_causal_mask`, decoder layers
taking `past_key_value=` and returning tuples), not the 5.5 of this image.  Three shims, none touching the reference's files, make its
classes run:
  1. a stub `x_transformers.x_transformers` (RotaryEmbedding / apply_rotary_pos_emb) for the conformer - built by `__init__`, never
     executed on the inference path (llm_sft.py:62-65,112-115 are commented out);
  2. `LlamaModel._update_causal_mask` given back as a no-op attribute so that `CustomLlamaModel.__init__` (llm.py:79) finishes;
  3. `llm_forward` (llm.py:150-228) bound to an equivalent over THE REFERENCE OBJECT'S OWN `layers / norm / rotary_emb` through
     transformers 5.5's `LlamaModel.forward` (mask None + SDPA == causal, `DynamicCache`): the one function whose body is
     version-specific - and the one oracle/make_golden_lm.py already pins against `transformers.LlamaModel` separately.
Everything else runs UNMODIFIED: `LLM_SFT.__init__`, the conditioning prefix, the teacher-forced `forward`, `loss_function`, the two
decoding loops of `generate` with their range masks, `sample_logits`.  The script checks the oracle against it (tokens, loss, accuracy,
logits, the filtered support of `sample_logits`) and writes tests/golden/lm_reference.npz + lm_reference_pinning_report.json;
tests/test_host.py re-checks the oracle against the fixture without the reference.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/QuarkAudio-UniSE/model/llm"


def import_reference_llm():
    import transformers
    xt = types.ModuleType("x_transformers")
    xtx = types.ModuleType("x_transformers.x_transformers")

    class RotaryEmbedding(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
    xtx.RotaryEmbedding = RotaryEmbedding
    xtx.apply_rotary_pos_emb = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("conformer is dead code on this path"))
    xt.x_transformers = xtx
    sys.modules.setdefault("x_transformers", xt)
    sys.modules.setdefault("x_transformers.x_transformers", xtx)
    if not hasattr(transformers.LlamaModel, "_update_causal_mask"):
        transformers.LlamaModel._update_causal_mask = lambda self, *a, **k: None
    pkg = types.ModuleType("refllm")
    pkg.__path__ = [REF]
    sys.modules["refllm"] = pkg
    mods = {}
    for name in ("conformer", "llm", "llm_sft"):
        spec = importlib.util.spec_from_file_location(f"refllm.{name}", os.path.join(REF, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refllm.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["llm_sft"].LLM_SFT


def build_reference(cfg, sd):
    from transformers import DynamicCache, LlamaModel
    from transformers.modeling_outputs import BaseModelOutputWithPast
    LLM_SFT = import_reference_llm()
    ref = LLM_SFT(num_tasks=cfg["num_tasks"], task_map=cfg["task_map"], feats_dim=cfg["feats_dim"],
                  llm_base_config=cfg["llm_base_config"]).eval()
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("cond_", "rotary_emb.")) for k in missing), missing     # the dead conformer keeps its own init
    hf = LlamaModel(ref.config).eval()
    hf.layers, hf.norm, hf.rotary_emb = ref.layers, ref.norm, ref.rotary_emb

    def llm_forward(inputs_embeds, attention_mask=None, past_key_values=None, use_cache=False, **unused):
        assert attention_mask is None
        if use_cache and past_key_values is None:
            past_key_values = DynamicCache()                                         # llm.py:169-170
        out = hf(inputs_embeds=inputs_embeds, past_key_values=past_key_values, use_cache=use_cache)
        return BaseModelOutputWithPast(last_hidden_state=out.last_hidden_state, past_key_values=out.past_key_values if use_cache else None)
    ref.llm_forward = llm_forward
    return ref


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def main():
    from oracle import llama
    cfg = llama.lm_small()
    seed, gain = 5, 4.0
    sd = llama.make_lm_state_dict(cfg, seed, gain)
    ref = build_reference(cfg, sd)
    b = cfg["llm_base_config"]
    g = torch.Generator().manual_seed(11)
    B, T, Te = 3, 12, 7
    mix = torch.randn(B, T, cfg["feats_dim"], generator=g)
    enr = torch.randn(B, Te, cfg["feats_dim"], generator=g)
    mel = torch.zeros(B, T, 80)
    emel = torch.zeros(B, Te, 80)
    gids = torch.randint(0, b["global_size"], (B, 32), generator=g)
    sids = torch.randint(0, b["semantic_size"], (B, T), generator=g)
    report, out = {}, dict(mix=mix.numpy(), enroll=enr.numpy(), gids=gids.numpy(), sids=sids.numpy())
    with torch.no_grad():
        for task, e_mel, e_feats in (("se", None, None), ("tse", emel, enr), ("rtse", emel, enr)):
            # ---- teacher-forced forward (llm_sft.py:37-89, llm.py:87-104), unmodified
            loss_r, acc_r = ref(task, e_mel, e_feats, mel, mix, gids, sids)
            loss_o, acc_o, logits_o = llama.sft_forward(sd, cfg, task, e_feats, mix, gids, sids, return_logits=True)
            # ---- greedy generate (llm_sft.py:93-195, llm.py:253-289), unmodified
            gg_r, ss_r = ref.generate(task, e_mel, e_feats, mel, mix, do_sample=False)
            gg_o, ss_o, margins = llama.sft_generate(sd, cfg, task, e_feats, mix, T, return_margins=True)
            same = bool(torch.equal(gg_r, gg_o) and torch.equal(ss_r, ss_o))
            report[task] = dict(loss_reference=float(loss_r), loss_oracle=float(loss_o), loss_rel=abs(float(loss_r) - float(loss_o)) / abs(float(loss_r)),
                                acc_reference=float(acc_r), acc_oracle=float(acc_o), tokens_identical=same, min_margin=float(margins.min()),
                                global_shape=list(gg_r.shape), semantic_shape=list(ss_r.shape))
            print(task, report[task])
            assert report[task]["loss_rel"] < 1e-5 and float(acc_r) == float(acc_o) and same
            out.update({f"{task}.loss": np.float64(loss_r), f"{task}.acc": np.float64(acc_r), f"{task}.gen_global": gg_r.numpy(),
                        f"{task}.gen_semantic": ss_r.numpy()})
        # ---- sample_logits (llm.py:253-289), unmodified: it filters its argument in place; what survives, divided by the
        # temperature, is the distribution torch.multinomial draws from
        V = 3 + b["global_size"] + b["semantic_size"]
        lg = 3.0 * torch.randn(4, V, generator=g)
        lg[:, :3] = float("-inf")
        for top_k, top_p, temp in ((50, 0.95, 0.8), (5, 0.5, 1.0), (20, 1.0, 0.3)):
            work = lg.clone()
            greedy = ref.sample_logits(work, temperature=temp, top_k=top_k, top_p=top_p, do_sample=False)
            probs_o = llama.sample_filter(lg.clone(), temperature=temp, top_k=top_k, top_p=top_p)      # the oracle returns the probabilities
            probs_r = torch.softmax(work / temp, -1)                                     # what the reference hands to torch.multinomial
            sup_r, sup_o = torch.isfinite(work), probs_o > 0
            assert torch.equal(sup_r, sup_o), "filtered support differs"
            e = float((probs_r - probs_o).abs().max())
            assert e < 1e-6 and torch.equal(greedy[:, 0], probs_o.argmax(-1))
            torch.manual_seed(123)
            draw = ref.sample_logits(lg.clone(), temperature=temp, top_k=top_k, top_p=top_p, do_sample=True)
            assert bool(sup_o.gather(1, draw).all())                                 # every draw lies inside the oracle's support
            report[f"sample_k{top_k}_p{top_p}_t{temp}"] = dict(support_sizes=sup_r.sum(1).tolist(), max_abs_diff=e)
            out[f"sample.k{top_k}.p{top_p}.t{temp}.support"] = np.packbits(sup_r.numpy(), axis=1)
            out[f"sample.k{top_k}.p{top_p}.t{temp}.probs_max"] = probs_r.max(-1).values.numpy()
        out["sample.logits"] = lg.numpy()
    # ---- the shipped architecture (hidden 512, 12 layers, 8 heads, vocab 12 291): report only
    full = llama.LM_FULL
    sdf = llama.make_lm_state_dict(full, 7, 2.0)
    reff = build_reference(full, sdf)
    gf = torch.Generator().manual_seed(12)
    mixf, enrf = torch.randn(2, 9, full["feats_dim"], generator=gf), torch.randn(2, 5, full["feats_dim"], generator=gf)
    gidf = torch.randint(0, full["llm_base_config"]["global_size"], (2, 32), generator=gf)
    sidf = torch.randint(0, full["llm_base_config"]["semantic_size"], (2, 9), generator=gf)
    with torch.no_grad():
        lr, ar = reff("tse", torch.zeros(2, 5, 80), enrf, torch.zeros(2, 9, 80), mixf, gidf, sidf)
        lo, ao = llama.sft_forward(sdf, full, "tse", enrf, mixf, gidf, sidf)
        gr, sr = reff.generate("tse", torch.zeros(2, 5, 80), enrf, torch.zeros(2, 9, 80), mixf, do_sample=False)
        go, so, mg = llama.sft_generate(sdf, full, "tse", enrf, mixf, 9, return_margins=True)
    report["full_config_tse"] = dict(loss_reference=float(lr), loss_oracle=float(lo), acc_equal=float(ar) == float(ao),
                                     tokens_identical=bool(torch.equal(gr, go) and torch.equal(sr, so)), min_margin=float(mg.min()))
    print("full", report["full_config_tse"])
    assert abs(float(lr) - float(lo)) < 1e-5 * abs(float(lr)) and report["full_config_tse"]["tokens_identical"]
    meta = dict(cfg=cfg, seed=seed, gain=gain, T=T, Te=Te, shims=["x_transformers stub (dead conformer)", "LlamaModel._update_causal_mask no-op",
                                                                 "llm_forward via transformers 5.5 LlamaModel.forward over the reference object's own layers"],
                reference="QuarkAudio-UniSE/model/llm/llm_sft.py:13-195, llm.py:13-148,253-289 (unmodified)")
    np.savez_compressed(os.path.join(GOLD, "lm_reference.npz"), meta=np.array(json.dumps(meta)), **out)
    json.dump(report, open(os.path.join(GOLD, "lm_reference_pinning_report.json"), "w"), indent=1)
    print("wrote lm_reference.npz", os.path.getsize(os.path.join(GOLD, "lm_reference.npz")), "bytes")


if __name__ == "__main__":
    main()
