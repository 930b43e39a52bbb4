"""ncu target: UniSE LM teacher-forced forward, B=32 x 536 positions (two calls; the second is the one to read)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unified_audio_b200.llm import LLM_SFT
LM = dict(cond_dim=80, global_size=4096, semantic_size=8192, hidden_size=512, num_layers=12, num_attention_heads=8,
          dropout_p=0.1, max_position_embeddings=4096, label_smoothing=0.1)
m = LLM_SFT(num_tasks=3, task_map=dict(se=0, tse=1, rtse=2), feats_dim=768, llm_base_config=LM).cuda()
with torch.no_grad():
    for n, p in m.named_parameters():
        p.copy_(torch.randn_like(p) * (0.05 if p.dim() >= 2 else 1.0))
m._w = None
B, T = 32, 250
mix = torch.randn(B, T, 768, device="cuda")
g, s = torch.randint(0, 4096, (B, 32), device="cuda"), torch.randint(0, 8192, (B, T), device="cuda")
for _ in range(2):
    m("se", None, None, mix, mix, g, s)
    torch.cuda.synchronize()
    print("forward done")
