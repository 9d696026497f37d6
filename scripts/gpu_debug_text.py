"""Text tower alone vs the oracle (bf16 and fp32 modes) with a linear loss, several input seeds: gradient error next to
the number of CLS features whose ReLU mask (oa_model.py:68) differs between the implementations. B=2, L=32."""
import os, sys, statistics
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
dev = torch.device("cuda")
for seed in range(5, 11):
    g = torch.Generator().manual_seed(seed)
    text = O.synth_text(B, L, g)
    coef = torch.randn(B, 256, generator=g)

    def oracle_run(bf16):
        p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in w.items()}
        cfg = O.OracleCfg(bf16=bf16)
        hid = O.distilbert(text["input_ids"], text["attention_mask"], p, cfg)[:, 0]
        te = O.linear(torch.relu(hid.float()), p["txt_proj.1.weight"], p["txt_proj.1.bias"], cfg)
        (te * coef).sum().backward()
        return hid.detach(), {k: v.grad for k, v in p.items() if v.is_floating_point() and v.grad is not None}

    h16, g16 = oracle_run(True)
    h32, g32 = oracle_run(False)
    params = {k: v.to(dev).clone().requires_grad_(v.is_floating_point()) for k, v in w.items()}
    tnamed = [(k, v) for k, v in params.items() if v.is_floating_point()]
    eng = TextEngine(dev, heads=12)
    te = run_tower(eng, tnamed, input_ids=text["input_ids"].to(dev), attention_mask=text["attention_mask"].to(dev))
    (te * coef.to(dev)).sum().backward()
    hc = eng.saved["last32"].view(B, L, -1)[:, 0].detach().cpu()
    gc = {k: v.grad.detach().cpu() for k, v in params.items() if v.is_floating_point() and v.grad is not None}
    keys = [k for k in g16 if not k.endswith("k_lin.bias") and k in gc and k.startswith("text_model.transformer")]
    ours = statistics.median(rel(gc[k], g16[k]) for k in keys)
    floor = statistics.median(rel(g16[k], g32[k]) for k in keys)
    print("seed %d: grad err ours-vs-bf16 %.3e, bf16-vs-fp32 %.3e | relu mask flips ours-vs-bf16 %d, bf16-vs-fp32 %d | hidden rel err %.2e / %.2e"
          % (seed, ours, floor, int(((hc > 0) != (h16 > 0)).sum()), int(((h16 > 0) != (h32 > 0)).sum()), rel(hc, h16), rel(h16, h32)))
