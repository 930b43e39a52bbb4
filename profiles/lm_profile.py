"""ncu target: UniSE LM generate, B=32, short (T=8), no CUDA graph so every kernel is listed."""
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
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mix = torch.randn(32, 250, 768, device="cuda")
class Mel:  # only .size(1) is consumed (llm_sft.py:108)
    def __init__(self, t): self.t = t
    def size(self, d): return self.t
m.generate("se", None, None, Mel(T), mix, do_sample=False, use_cuda_graph=False)
torch.cuda.synchronize()
