"""Text tower alone vs the oracle (bf16 and fp32 modes) with a linear loss: isolates tower backward accuracy from the
InfoNCE amplification. B=2, L=32."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import oracle as O
from oracle.weights import dual_encoder_spec, fill_seeded
from oa_transformer_b200.engine import TextEngine
from oa_transformer_b200.functional import run_tower

def rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))

B, L = int(os.environ.get("TB", 2)), int(os.environ.get("TL", 32))
spec = {k: v for k, v in dual_encoder_spec(frames=2, depth=1).items() if k.startswith(("text_model.", "txt_proj."))}
w = fill_seeded(spec, 91, 0.02)
g = torch.Generator().manual_seed(5)
text = O.synth_text(B, L, g, ragged=bool(int(os.environ.get("RAGGED", 0))))
coef = torch.randn(B, 256, generator=g)

def oracle_run(bf16):
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in w.items()}
    te = O.compute_text(text, p, O.OracleCfg(bf16=bf16))
    (te * coef).sum().backward()
    return te.detach(), {k: v.grad for k, v in p.items() if v.is_floating_point() and v.grad is not None}

t16, g16 = oracle_run(True)
t32, g32 = oracle_run(False)
dev = torch.device("cuda")
params = {k: v.to(dev).clone().requires_grad_(v.is_floating_point()) for k, v in w.items()}
tnamed = [(k, v) for k, v in params.items() if v.is_floating_point()]
te = run_tower(TextEngine(dev, heads=12), tnamed, input_ids=text["input_ids"].to(dev), attention_mask=text["attention_mask"].to(dev))
(te * coef.to(dev)).sum().backward()
gc = {k: v.grad.detach().cpu() for k, v in params.items() if v.is_floating_point() and v.grad is not None}
print("emb err ours-vs-bf16 %.3e  bf16-vs-fp32 %.3e" % (rel(te.detach().cpu(), t16), rel(t16, t32)))
rows = []
for k in g16:
    if k.endswith("k_lin.bias") or k not in gc:
        continue
    rows.append((rel(gc[k], g16[k]), rel(g16[k], g32[k]), k))
rows.sort(reverse=True)
for r in rows[:14]:
    print("ours-vs-bf16 %.3e   floor(bf16-vs-fp32) %.3e   %s" % r)
import statistics
print("median ours %.3e floor %.3e" % (statistics.median(r[0] for r in rows), statistics.median(r[1] for r in rows)))
