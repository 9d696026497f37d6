"""Host-side schedules of the two towers (engine.VideoEngine / engine.TextEngine) on CPU: the operator layer is replaced,
for this test only, by the plain-torch stand-in tests/fake_liboat.py, so that buffer management, the frozen-in-time
residual wiring, operand packing and the hand-written backward bookkeeping (which gradient lands in which tensor, the
fused bias-gradient reductions, the CLS-row-only final LayerNorm, q/k/v packing of the text tower) are checked against
the oracle without a GPU. The CUDA kernels themselves are checked on the GPU (tests/test_*_gpu.py)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_liboat  # noqa: E402

from oracle import oracle as O  # noqa: E402
from oracle.weights import fill_seeded, text_tower_spec, video_tower_spec  # noqa: E402


def rel(a, b):
    d, n = float((a.double() - b.double()).norm()), float(b.double().norm())
    return 0.0 if d <= 1e-9 else d / max(n, 1e-30)


@pytest.fixture()
def engine_on_fake_ops(monkeypatch):
    from oa_transformer_b200 import engine
    monkeypatch.setattr(engine, "ops", fake_liboat)
    monkeypatch.setattr(engine, "SIDE_STREAM", False)
    return engine


@pytest.mark.parametrize("frames,objects,dim,heads", [(2, 3, 128, 2), (3, 0, 128, 2), (2, 2, 256, 4)])
def test_video_engine_schedule_matches_oracle(engine_on_fake_ops, frames, objects, dim, heads):
    """dim 256: the backward hands delta = rowsum(dO * O) from the projection dgrad (gemm act 4) to the attention backward
    (the stand-in's attn_bwd asserts that the delta it receives belongs to that attention's dO and O)."""
    engine = engine_on_fake_ops
    depth, B = 2, 2
    spec = video_tower_spec(depth=depth, dim=dim, frames=frames, grid=2, patch=16, objects=objects > 0)
    spec["vid_proj.0.weight"], spec["vid_proj.0.bias"] = (32, dim), (32,)
    w = fill_seeded(spec, 3, 0.05)
    g = torch.Generator().manual_seed(4)
    video = torch.randn(B, frames, 3, 32, 32, generator=g)
    objs = O.synth_objects(B, frames, objects, g) if objects else None
    coef = torch.randn(B, 32, generator=g)

    p = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    ref = O.compute_video(video, p, O.OracleCfg(heads=heads, bf16=True), objs)
    (ref * coef).sum().backward()

    eng = engine.VideoEngine(torch.device("cpu"), heads=heads)
    params = {k: v.clone() for k, v in w.items()}
    out = eng.forward(params, video, objs)
    assert rel(out, ref.detach()) < 1e-3      # same storage points and split-bf16 rows (a missing split row costs 4e-3)
    named = [(k, torch.nn.Parameter(v)) for k, v in params.items()]
    book = engine.GradBook(named, torch.device("cpu"))
    eng.backward(params, book, coef.clone())
    for name, ref_p in p.items():
        assert ref_p.grad is not None, name
        assert rel(book[name], ref_p.grad) < 3e-2, (name, rel(book[name], ref_p.grad))


def test_qkv_bias_gradient_from_softmax_identities(engine_on_fake_ops, monkeypatch):
    """engine.QKV_BIAS_IDENTITY: the qkv bias gradient of the divided attentions (video_transformer.py:102) from a column sum
    over dq only - the dk columns sum to zero (rows of dS sum to zero), the dv columns to db_proj . W_proj (rows of P sum to
    one, dO = dY_proj . W_proj). Against the fp32 oracle it must be at least as close as the column sum over all of dqkv,
    and its k part exactly zero (the full column sum leaves bf16 rounding noise there; the true value is ~1e-7)."""
    engine = engine_on_fake_ops
    dim, heads, depth, B, frames, objects = 256, 4, 2, 2, 2, 2
    spec = video_tower_spec(depth=depth, dim=dim, frames=frames, grid=2, patch=16, objects=True)
    spec["vid_proj.0.weight"], spec["vid_proj.0.bias"] = (32, dim), (32,)
    w = fill_seeded(spec, 3, 0.05)
    g = torch.Generator().manual_seed(4)
    video = torch.randn(B, frames, 3, 32, 32, generator=g)
    objs = O.synth_objects(B, frames, objects, g)
    coef = torch.randn(B, 32, generator=g)
    p = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    (O.compute_video(video, p, O.OracleCfg(heads=heads, bf16=False), objs) * coef).sum().backward()
    err = {}
    for ident in (True, False):
        monkeypatch.setattr(engine, "QKV_BIAS_IDENTITY", ident)
        eng = engine.VideoEngine(torch.device("cpu"), heads=heads)
        params = {k: v.clone() for k, v in w.items()}
        eng.forward(params, video, objs)
        book = engine.GradBook([(k, torch.nn.Parameter(v)) for k, v in params.items()], torch.device("cpu"))
        eng.backward(params, book, coef.clone())
        for name in w:
            if name.endswith("qkv.bias"):
                err[(ident, name)] = rel(book[name], p[name].grad)
                if ident:
                    assert float(book[name][dim:2 * dim].abs().max()) == 0.0, name
                    assert rel(book[name][2 * dim:], p[name].grad[2 * dim:]) < 2e-2, name
    names = sorted(n for (i, n) in err if i)
    assert len(names) == 2 * depth
    for n in names:
        assert err[(True, n)] <= err[(False, n)] * 1.05 + 1e-4, (n, err[(True, n)], err[(False, n)])


@pytest.mark.parametrize("tokens,region_layer,depth", [("final", None, 2), ("region", 1, 2), ("region", 2, 2), ("region", 2, 3)])
def test_video_engine_token_features_schedule(engine_on_fake_ops, tokens, region_layer, depth):
    """forward_features' second result (video_transformer.py:351: x[:, 1:] after the final norm) and the region variant's
    region_norm(x after K blocks)[:, 1:] (oa_video_transformer_region.py:364-376), forward and backward: the gradient
    through the token features joins the CLS path at the right block, and the fused bias-gradient sums stay exact."""
    engine = engine_on_fake_ops
    dim, heads, B, frames = 128, 2, 2, 2
    spec = video_tower_spec(depth=depth, dim=dim, frames=frames, grid=2, patch=16)
    spec["vid_proj.0.weight"], spec["vid_proj.0.bias"] = (32, dim), (32,)
    if tokens == "region":
        spec["video_model.region_norm.weight"], spec["video_model.region_norm.bias"] = (dim,), (dim,)
    w = fill_seeded(spec, 13, 0.05)
    g = torch.Generator().manual_seed(14)
    video = torch.randn(B, frames, 3, 32, 32, generator=g)
    T = 1 + frames * 4
    coef = torch.randn(B, 32, generator=g)
    ctok = torch.randn(B, T - 1, dim, generator=g)

    p = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    cfg = O.OracleCfg(heads=heads, bf16=True)
    cls_ref, tok_ref = O.video_tower(video, p, cfg, return_tokens=True, region_layer=region_layer)
    emb_ref = O.linear(cls_ref, p["vid_proj.0.weight"], p["vid_proj.0.bias"], cfg, True)
    ((emb_ref * coef).sum() + (tok_ref * ctok).sum()).backward()

    eng = engine.VideoEngine(torch.device("cpu"), heads=heads)
    params = {k: v.clone() for k, v in w.items()}
    out, tok = eng.forward(params, video, tokens=tokens, region_layer=region_layer or 6)
    assert rel(out, emb_ref.detach()) < 1e-3 and rel(tok[:, 1:], tok_ref.detach()) < 1e-3
    named = [(k, torch.nn.Parameter(v)) for k, v in params.items()]
    book = engine.GradBook(named, torch.device("cpu"))
    dtok = torch.zeros(B, T, dim)
    dtok[:, 1:] = ctok
    eng.backward(params, book, coef.clone(), dtok)
    for name, ref_p in p.items():
        assert ref_p.grad is not None, name
        assert rel(book[name], ref_p.grad) < 3e-2, (name, rel(book[name], ref_p.grad))


def test_text_engine_schedule_matches_oracle(engine_on_fake_ops):
    engine = engine_on_fake_ops
    dim, heads, layers, B, L = 128, 2, 2, 3, 6
    spec = text_tower_spec(layers=layers, dim=dim, hidden=256, vocab=50, max_pos=16)
    spec["txt_proj.1.weight"], spec["txt_proj.1.bias"] = (32, dim), (32,)
    w = fill_seeded(spec, 5, 0.05)
    g = torch.Generator().manual_seed(6)
    ids = torch.randint(1, 50, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, 4:] = 0
    mask[2, 3:] = 0
    coef = torch.randn(B, 32, generator=g)

    p = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    ref = O.compute_text({"input_ids": ids, "attention_mask": mask}, p, O.OracleCfg(heads=heads, bf16=True, text_layers=layers))
    (ref * coef).sum().backward()

    eng = engine.TextEngine(torch.device("cpu"), heads=heads)
    params = {k: v.clone() for k, v in w.items()}
    out = eng.forward(params, ids, mask)
    assert rel(out, ref.detach()) < 1e-3
    named = [(k, torch.nn.Parameter(v)) for k, v in params.items()]
    book = engine.GradBook(named, torch.device("cpu"))
    eng.backward(params, book, coef.clone())
    for name, ref_p in p.items():
        if name.endswith("k_lin.bias"):          # softmax is invariant to the key bias: the gradient is exactly zero
            continue
        if ref_p.grad is None:
            continue
        assert rel(book[name], ref_p.grad) < 8e-2, (name, rel(book[name], ref_p.grad))


def test_tower_gradients_alias_the_flat_book_and_still_accumulate(engine_on_fake_ops, monkeypatch):
    """functional.TowerRunner returns fresh views of the flat gradient book: autograd adopts them as p.grad without a copy
    (so an in-place all-reduce of the book IS the all-reduce of p.grad - bench.py / ADVICE r1), and a second backward
    without zero_grad still accumulates (the aliasing p.grad is detached from the book before the book is zeroed)."""
    engine = engine_on_fake_ops
    from oa_transformer_b200 import functional
    dim, heads, layers, B, L = 128, 2, 1, 2, 4
    spec = text_tower_spec(layers=layers, dim=dim, hidden=256, vocab=50, max_pos=16)
    spec["txt_proj.1.weight"], spec["txt_proj.1.bias"] = (32, dim), (32,)
    w = fill_seeded(spec, 5, 0.05)
    named = [(k, torch.nn.Parameter(v.clone())) for k, v in w.items()]
    ids = torch.randint(1, 50, (B, L), generator=torch.Generator().manual_seed(1))
    eng = engine.TextEngine(torch.device("cpu"), heads=heads)
    out = functional.run_tower(eng, named, input_ids=ids, attention_mask=None)
    out.sum().backward()
    book = eng._gradbook
    first = {n: p.grad.clone() for n, p in named}
    assert all(p.grad.data_ptr() == book.views[n].data_ptr() for n, p in named)
    out = functional.run_tower(eng, named, input_ids=ids, attention_mask=None)
    out.sum().backward()
    for n, p in named:
        assert torch.allclose(p.grad, 2 * first[n], rtol=1e-6, atol=1e-8), n


def _text_dropout_multipliers(ops_mod, B, L, D, H, layers, pd, pa, seed, device=None):
    """site -> multiplier tensor, from the same draws the operator layer makes (ops.dropout_mask)."""
    drop = {}
    def mult(shape, p, site):
        n = 1
        for d in shape:
            n *= d
        return ops_mod.dropout_mask(n, p, seed, site, device).view(shape).float().cpu() / (1.0 - p)
    drop[0] = mult((B, L, D), pd, 0)
    for i in range(layers):
        drop[1 + 3 * i] = mult((B, H, L, L), pa, 1 + 3 * i)
        drop[2 + 3 * i] = mult((B, L, D), pd, 2 + 3 * i)
    return drop


def test_text_engine_training_dropout_schedule(engine_on_fake_ops):
    """DistilBERT's training-mode dropout (embedding, attention weights, FFN output - the reference keeps
    text_model.train(), model/oa_model.py:28) in the text schedule, forward and backward, vs the oracle with the SAME
    masks injected."""
    engine = engine_on_fake_ops
    dim, heads, layers, B, L = 128, 2, 2, 3, 6
    spec = text_tower_spec(layers=layers, dim=dim, hidden=256, vocab=50, max_pos=16)
    spec["txt_proj.1.weight"], spec["txt_proj.1.bias"] = (32, dim), (32,)
    w = fill_seeded(spec, 5, 0.05)
    g = torch.Generator().manual_seed(6)
    ids = torch.randint(1, 50, (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, 4:] = 0
    coef = torch.randn(B, 32, generator=g)
    pd, pa, seed = 0.1, 0.1, 1234567
    drop = _text_dropout_multipliers(fake_liboat, B, L, dim, heads, layers, pd, pa, seed)
    p = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    cfg = O.OracleCfg(heads=heads, bf16=True, text_layers=layers)
    ref = O.compute_text({"input_ids": ids, "attention_mask": mask}, p, cfg, drop=drop)
    ref_eval = O.compute_text({"input_ids": ids, "attention_mask": mask}, w, cfg)
    (ref * coef).sum().backward()
    eng = engine.TextEngine(torch.device("cpu"), heads=heads)
    params = {k: v.clone() for k, v in w.items()}
    out = eng.forward(params, ids, mask, dropout={"p": pd, "p_attn": pa, "seed": seed})
    assert rel(out, ref.detach()) < 1e-3 and rel(out, ref_eval.detach()) > 5e-2      # dropout really changes the output
    named = [(k, torch.nn.Parameter(v)) for k, v in params.items()]
    book = engine.GradBook(named, torch.device("cpu"))
    eng.backward(params, book, coef.clone())
    for name, ref_p in p.items():
        if name.endswith("k_lin.bias") or ref_p.grad is None:
            continue
        assert rel(book[name], ref_p.grad) < 8e-2, (name, rel(book[name], ref_p.grad))
