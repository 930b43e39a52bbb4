"""Pin oracle/llama.py against `transformers.LlamaModel` (the engine the reference builds its layers from,
U/model/llm/llm.py:63-79) and write tests/golden/lm_small.npz.  Build container only.
Usage: python -m oracle.make_golden_lm"""
import json, os
import numpy as np
import torch
from oracle import llama

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

def hf_model(cfg, sd):
    from transformers import LlamaConfig, LlamaModel
    b = cfg["llm_base_config"]
    V = 3 + b["global_size"] + b["semantic_size"]
    c = LlamaConfig(vocab_size=V, hidden_size=b["hidden_size"], num_hidden_layers=b["num_layers"],
                    num_attention_heads=b["num_attention_heads"], intermediate_size=b["hidden_size"] * 4,
                    attention_dropout=b["dropout_p"], max_position_embeddings=b["max_position_embeddings"])
    m = LlamaModel(c).eval()
    msd = {k: v for k, v in sd.items() if k.startswith("layers.") or k == "norm.weight"}
    missing, unexpected = m.load_state_dict(msd, strict=False)
    assert not unexpected and all("embed_tokens" in k or "rotary" in k for k in missing), (missing, unexpected)
    assert c.rms_norm_eps == 1e-6 and c.num_key_value_heads == b["num_attention_heads"]
    return m

def rel(a, b): return float((a - b).abs().max() / b.abs().max())

def main():
    torch.manual_seed(0)
    report = {}
    for name, cfg, gain in (("small", llama.lm_small(), 4.0), ("full", llama.LM_FULL, 1.0)):
        sd = llama.make_lm_state_dict(cfg, 5, gain)
        hf = hf_model(cfg, sd)
        H = cfg["llm_base_config"]["hidden_size"]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(2, 9, H, generator=g)
        with torch.no_grad():
            from transformers import DynamicCache
            ref_full = hf(inputs_embeds=x).last_hidden_state
            cache = DynamicCache()
            o1 = hf(inputs_embeds=x[:, :5], past_key_values=cache, use_cache=True).last_hidden_state
            steps = [hf(inputs_embeds=x[:, i:i + 1], past_key_values=cache, use_cache=True).last_hidden_state for i in range(5, 9)]
            ref_inc = torch.cat([o1] + steps, 1)
        mine_full, _ = llama.llm_forward(sd, cfg, x)
        m1, c = llama.llm_forward(sd, cfg, x[:, :5])
        outs = [m1]
        for i in range(5, 9):
            o, c = llama.llm_forward(sd, cfg, x[:, i:i + 1], c)
            outs.append(o)
        mine_inc = torch.cat(outs, 1)
        report[name] = dict(full_vs_hf=rel(mine_full, ref_full), inc_vs_hf=rel(mine_inc, ref_inc), inc_vs_full=rel(mine_inc, mine_full))
        print(name, report[name])
        assert report[name]["full_vs_hf"] < 5e-6 and report[name]["inc_vs_hf"] < 5e-6
    # golden fixture (small config): teacher-forced logits + greedy generation
    cfg = llama.lm_small()
    sd = llama.make_lm_state_dict(cfg, 5, 4.0)
    g = torch.Generator().manual_seed(3)
    B, T = 3, 10
    mix = torch.randn(B, T, cfg["feats_dim"], generator=g)
    enr = torch.randn(B, 6, cfg["feats_dim"], generator=g)
    gids = torch.randint(0, cfg["llm_base_config"]["global_size"], (B, 32), generator=g)
    sids = torch.randint(0, cfg["llm_base_config"]["semantic_size"], (B, T), generator=g)
    loss, acc, logits = llama.sft_forward(sd, cfg, "tse", enr, mix, gids, sids, return_logits=True)
    gg, ss, margins = llama.sft_generate(sd, cfg, "se", None, mix, T, return_margins=True)
    np.savez_compressed(os.path.join(GOLD, "lm_small.npz"), mix=mix.numpy(), enroll=enr.numpy(), gids=gids.numpy(),
                        sids=sids.numpy(), loss=loss.numpy(), acc=acc.numpy(), logits=logits.numpy(), gen_global=gg.numpy(),
                        gen_semantic=ss.numpy(), gen_margins=margins.numpy(),
                        meta=np.array(json.dumps(dict(cfg=cfg, seed=5, gain=4.0, report=report))))
    print("min generation margin", float(margins.min()))

if __name__ == "__main__":
    main()
