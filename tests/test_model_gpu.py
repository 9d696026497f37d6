"""End-to-end parity of the CUDA dual encoder (towers + sim matrix + InfoNCE, forward and backward) on the GPU:
  * against OUTPUTS OF THE REFERENCE (tests/golden/dual_small.pt, cfg1_full.pt: fp32 PyTorch);
  * against the pinned oracle in bf16-operand mode on the same inputs.

Tolerances (BASELINE.json north_star: 1e-3 on bf16 logits/grads):
  LOGIT_TOL = 1e-3 absolute on cosine logits in [-1, 1], against the REFERENCE's fp32 outputs and against the oracle.
    Plain bf16 operands everywhere cannot guarantee that (oracle bf16 mode without split rows vs oracle fp32: 7.6e-4
    on cfg1, 1.5e-3 on the small fixture, 1.3e-3 at the benchmark geometry), so the rows the logits depend on
    directly - CLS rows, text tower, projections - take split-bf16 (three-term) products in the forward pass
    (engine.py docstring; oracle cfg.split mirrors it): 3.9e-4 / 6.5e-4 / 1.4e-4 on the same cases.
  Gradients (gate_grads): bf16 operands put a floor under per-tensor gradient error that no kernel can beat - the
    distance between the oracle in bf16-operand mode and the fp32 oracle on the same inputs (`noise_floor`, measured
    in every test; 1-4 % per tensor for InfoNCE at T = 0.05 on random-init towers, whose embeddings are nearly
    identical across samples so that dL/dembedding is a difference of almost equal terms; 0.5-0.7 % under a linear
    loss). The CUDA path must be no further from the fp32 truth (the reference's outputs, or the fp32 oracle) than
    the bf16 oracle is: 1x the floor, with 10 % slack on the median over tensors and 1.5x on the single worst
    tensor. Against the bf16 oracle itself the expected distance is sqrt(2) x the floor (two independent roundings
    of the same computation: the kernels round unnormalised softmax weights, the oracle normalised ones), gated at
    1.5x / 2x. Mathematically-zero gradients (softmax is invariant to the key bias: *.k_lin.bias) are excluded.
"""
import json
import os

import pytest
import torch

from oracle import oracle as O
from oracle.weights import dual_encoder_spec, fill_seeded

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def rel(a, b, floor=1e-6):
    d, n = float((a.double() - b.double()).norm()), float(b.double().norm())
    return 0.0 if d <= floor else d / max(n, 1e-30)


LOGIT_TOL = 1e-3


def cuda_dual(p_cpu, video, ids, mask, heads, objects=None, want_grads=True, temperature=0.05):
    """Run the CUDA path through the public functional layer with leaf parameters built from a weight dict."""
    from oa_transformer_b200.engine import TextEngine, VideoEngine
    from oa_transformer_b200.functional import norm_softmax_loss, run_tower, sim_matrix
    dev = torch.device("cuda")
    params = {k: v.to(dev).clone().requires_grad_(v.is_floating_point()) for k, v in p_cpu.items()}
    vnamed = [(k, v) for k, v in params.items() if k.startswith("video_model.") or k.startswith("vid_proj.")]
    tnamed = [(k, v) for k, v in params.items() if (k.startswith("text_model.") or k.startswith("txt_proj."))
              and v.is_floating_point()]
    ve = run_tower(VideoEngine(dev, heads=heads), vnamed, video=video.to(dev),
                   objects=None if objects is None else objects.to(dev))
    te = run_tower(TextEngine(dev, heads=heads), tnamed, input_ids=ids.to(dev),
                   attention_mask=None if mask is None else mask.to(dev))
    sims = sim_matrix(te, ve)
    loss = norm_softmax_loss(sims, temperature)
    grads = {}
    if want_grads:
        loss.backward()
        grads = {k: v.grad.detach().cpu() for k, v in params.items() if v.is_floating_point() and v.grad is not None}
    torch.cuda.synchronize()
    return te.detach().cpu(), ve.detach().cpu(), sims.detach().cpu(), float(loss), grads


def oracle_dual(p_cpu, video, ids, mask, cfg, objects=None, temperature=0.05):
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in p_cpu.items()}
    data = {"video": video, "text": {"input_ids": ids, "attention_mask": mask}}
    if objects is not None:
        data["object"] = objects
    te, ve = O.dual_encoder(data, p, cfg)
    sims = O.sim_matrix(te, ve)
    loss = O.norm_softmax_loss(sims, temperature)
    loss.backward()
    grads = {k: v.grad for k, v in p.items() if torch.is_tensor(v) and v.is_floating_point() and v.grad is not None}
    return te.detach(), ve.detach(), sims.detach(), float(loss), grads


def summarize(tag, sims, sims_ref, loss, loss_ref, grads, grads_ref):
    errs = {k: rel(grads[k], grads_ref[k]) for k in grads_ref if k in grads and not k.endswith("k_lin.bias")}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    rep = {"case": tag, "logit_max_abs_err": float((sims - sims_ref).abs().max()), "loss": loss, "loss_ref": loss_ref,
           "grad_rel_err_max": max(errs.values()) if errs else None,
           "grad_rel_err_median": sorted(errs.values())[len(errs) // 2] if errs else None, "worst": worst,
           "missing": [k for k in grads_ref if k not in grads][:5]}
    os.makedirs(REPORT, exist_ok=True)
    with open(os.path.join(REPORT, "parity_%s.json" % tag), "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))
    return rep


def gate_grads(floor, vs_truth=None, vs_bf16_oracle=None):
    """See the module docstring: 1x the measured floor against the fp32 truth, sqrt(2)x against the bf16 oracle."""
    if vs_truth is not None:
        assert vs_truth["grad_rel_err_median"] < 1.1 * floor["grad_rel_err_median"] + 1e-3, (vs_truth, floor)
        assert vs_truth["grad_rel_err_max"] < 1.5 * floor["grad_rel_err_max"] + 1e-2, (vs_truth, floor)
    if vs_bf16_oracle is not None:
        assert vs_bf16_oracle["grad_rel_err_median"] < 1.5 * floor["grad_rel_err_median"] + 1e-3, (vs_bf16_oracle, floor)
        assert vs_bf16_oracle["grad_rel_err_max"] < 2.0 * floor["grad_rel_err_max"] + 1e-2, (vs_bf16_oracle, floor)


def noise_floor(w, video, ids, mask, cfg16, cfg32, objects=None, temperature=0.05, tag=""):
    """bf16-operand oracle vs fp32 oracle on the same inputs: the error that operand rounding alone introduces."""
    _, _, s16, l16, g16 = oracle_dual(w, video, ids, mask, cfg16, objects, temperature)
    _, _, s32, l32, g32 = oracle_dual(w, video, ids, mask, cfg32, objects, temperature)
    rep = summarize("noise_floor_%sT%g" % (tag, temperature), s16, s32, l16, l32, g16, g32)
    return (s16, l16, g16), rep


def test_small_dual_encoder_vs_reference_golden_and_bf16_oracle():
    g = torch.load(os.path.join(GOLD, "dual_small.pt"), map_location="cpu", weights_only=False)
    w = g["weights"]
    te, ve, sims, loss, grads = cuda_dual(w, g["video"], g["input_ids"], g["attention_mask"], heads=2)
    (osims, oloss, ograds), floor = noise_floor(w, g["video"], g["input_ids"], g["attention_mask"],
                                                O.OracleCfg(heads=2, text_layers=2, bf16=True),
                                                O.OracleCfg(heads=2, text_layers=2), tag="small_")
    # (1) vs the reference's fp32 outputs
    rep = summarize("small_vs_reference", sims, g["sims"], loss, float(g["loss"]), grads, g["grads"])
    assert rep["logit_max_abs_err"] < LOGIT_TOL and not rep["missing"]
    assert abs(loss - float(g["loss"])) < 1e-2
    # (2) vs the pinned oracle in bf16-operand mode
    rep16 = summarize("small_vs_bf16_oracle", sims, osims, loss, oloss, grads, ograds)
    assert rep16["logit_max_abs_err"] < LOGIT_TOL
    gate_grads(floor, vs_truth=rep, vs_bf16_oracle=rep16)


def test_cfg1_full_size_vs_reference_golden_and_bf16_oracle():
    """BASELINE.json configs[0]: 2-frame 64x64, 0 objects, 8-token text, batch 4, full ViT-B/16 + DistilBERT."""
    g = torch.load(os.path.join(GOLD, "cfg1_full.pt"), map_location="cpu", weights_only=False)
    w = fill_seeded(g["shapes"], g["weight_seed"], g["weight_scale"])
    te, ve, sims, loss, grads = cuda_dual(w, g["video"], g["input_ids"], g["attention_mask"], heads=12)
    rep = summarize("cfg1_vs_reference", sims, g["sims"], loss, float(g["loss"]), grads, g["grads_subset"])
    assert rep["logit_max_abs_err"] < LOGIT_TOL
    assert abs(loss - float(g["loss"])) < 1e-2
    (osims, oloss, ograds), floor = noise_floor(w, g["video"], g["input_ids"], g["attention_mask"],
                                                O.OracleCfg(bf16=True), O.OracleCfg(), tag="cfg1_")
    rep16 = summarize("cfg1_vs_bf16_oracle", sims, osims, loss, oloss, grads, ograds)
    assert rep16["logit_max_abs_err"] < LOGIT_TOL
    assert abs(loss - oloss) < 2e-3 * max(1.0, abs(oloss))
    # the fixture keeps the reference's gradients for a subset of tensors only: the floor is taken over the same subset
    sub = {k: v for k, v in ograds.items() if k in g["grads_subset"]}
    _, _, s32, l32, g32 = oracle_dual(w, g["video"], g["input_ids"], g["attention_mask"], O.OracleCfg())
    floor_sub = summarize("noise_floor_cfg1_subset", osims, s32, oloss, l32, sub, g32)
    gate_grads(floor_sub, vs_truth=rep)
    gate_grads(floor, vs_bf16_oracle=rep16)


def test_cfg1_gradients_without_temperature_amplification():
    """Same network and inputs with T = 1 (the logit noise is no longer multiplied by 20 inside the loss): the CUDA
    gradients must sit within the operand-rounding noise floor measured the same way (bf16 oracle vs fp32 oracle).
    Bias gradients are sums over the batch of nearly cancelling rows, which is where that floor is largest."""
    g = torch.load(os.path.join(GOLD, "cfg1_full.pt"), map_location="cpu", weights_only=False)
    w = fill_seeded(g["shapes"], g["weight_seed"], g["weight_scale"])
    te, ve, sims, loss, grads = cuda_dual(w, g["video"], g["input_ids"], g["attention_mask"], heads=12, temperature=1.0)
    (osims, oloss, ograds), floor = noise_floor(w, g["video"], g["input_ids"], g["attention_mask"],
                                                O.OracleCfg(bf16=True), O.OracleCfg(), temperature=1.0, tag="cfg1_")
    _, _, s32, l32, g32 = oracle_dual(w, g["video"], g["input_ids"], g["attention_mask"], O.OracleCfg(),
                                      temperature=1.0)
    rep32 = summarize("cfg1_T1_vs_fp32_oracle", sims, s32, loss, l32, grads, g32)
    rep16 = summarize("cfg1_T1_vs_bf16_oracle", sims, osims, loss, oloss, grads, ograds)
    assert abs(loss - oloss) < 1e-3 * max(1.0, abs(oloss))
    gate_grads(floor, vs_truth=rep32, vs_bf16_oracle=rep16)


def test_object_tokens_224_vs_bf16_oracle():
    """Spec-by-extension rows X1-X3: 224x224, 2 frames, 4 object regions per frame, ragged text."""
    spec = dual_encoder_spec(frames=2, objects=True)
    w = fill_seeded(spec, 77, 0.02)
    g = torch.Generator().manual_seed(78)
    B, Fr, Oo, L = 2, 2, 4, 12
    video = torch.randn(B, Fr, 3, 224, 224, generator=g)
    objects = O.synth_objects(B, Fr, Oo, g)
    text = O.synth_text(B, L, g, ragged=True)
    te, ve, sims, loss, grads = cuda_dual(w, video, text["input_ids"], text["attention_mask"], heads=12,
                                          objects=objects)
    (osims, oloss, ograds), floor = noise_floor(w, video, text["input_ids"], text["attention_mask"],
                                                O.OracleCfg(bf16=True), O.OracleCfg(), objects=objects, tag="objects_")
    _, _, s32, l32, g32 = oracle_dual(w, video, text["input_ids"], text["attention_mask"], O.OracleCfg(), objects)
    rep32 = summarize("objects_vs_fp32_oracle", sims, s32, loss, l32, grads, g32)
    rep16 = summarize("objects_vs_bf16_oracle", sims, osims, loss, oloss, grads, ograds)
    assert rep16["logit_max_abs_err"] < LOGIT_TOL and rep32["logit_max_abs_err"] < LOGIT_TOL
    gate_grads(floor, vs_truth=rep32, vs_bf16_oracle=rep16)
    assert "video_model.object_embed.weight" in grads


@pytest.mark.parametrize("frames,tag", [(4, "cfg3"), (16, "cfg5")])
def test_config_shaped_frames_and_objects_vs_bf16_oracle(frames, tag):
    """BASELINE configs[2] / configs[4] geometry (4 and 16 frames of 224x224, 36 object regions per frame, 32-token text:
    T = 929 / 3713 tokens per video) on a 2-block video tower, batch 2: logits vs the bf16 oracle, and the backward of
    both towers under a LINEAR loss on the embeddings. (With two pairs and random weights the InfoNCE gradient of one
    tower is proportional to the DIFFERENCE of the other tower's two nearly identical embeddings, which turns 1e-3
    embedding noise into 10 % gradient noise - a property of the toy problem, not of the kernels; the InfoNCE coupling
    itself is covered at batch 4 by the cfg1 tests.) Exercises the time kernels at Fp = 4 and Fp = 16 and the tcgen05
    space kernels at n = 232 inside the full schedule; gates are relative to the bf16-vs-fp32 oracle noise floor."""
    from oa_transformer_b200.engine import TextEngine, VideoEngine
    from oa_transformer_b200.functional import run_tower, sim_matrix
    spec = dual_encoder_spec(frames=frames, objects=True, depth=2)
    w = fill_seeded(spec, 91, 0.02)
    g = torch.Generator().manual_seed(92 + frames)
    B, Oo, L = 2, 36, 32
    video = torch.randn(B, frames, 3, 224, 224, generator=g)
    objects = O.synth_objects(B, frames, Oo, g)
    text = O.synth_text(B, L, g)
    ct, cv = torch.randn(B, 256, generator=g), torch.randn(B, 256, generator=g)

    def oracle_run(bf16):
        p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in w.items()}
        te, ve = O.dual_encoder({"video": video, "object": objects, "text": text}, p, O.OracleCfg(bf16=bf16))
        ((te * ct).sum() + (ve * cv).sum()).backward()
        return O.sim_matrix(te, ve).detach(), {k: v.grad for k, v in p.items()
                                               if v.is_floating_point() and v.grad is not None}

    s16, g16 = oracle_run(True)
    s32, g32 = oracle_run(False)
    dev = torch.device("cuda")
    params = {k: v.to(dev).clone().requires_grad_(v.is_floating_point()) for k, v in w.items()}
    vnamed = [(k, v) for k, v in params.items() if k.startswith(("video_model.", "vid_proj."))]
    tnamed = [(k, v) for k, v in params.items() if k.startswith(("text_model.", "txt_proj.")) and v.is_floating_point()]
    ve = run_tower(VideoEngine(dev, heads=12), vnamed, video=video.to(dev), objects=objects.to(dev))
    te = run_tower(TextEngine(dev, heads=12), tnamed, input_ids=text["input_ids"].to(dev),
                   attention_mask=text["attention_mask"].to(dev))
    sims = sim_matrix(te, ve).detach().cpu()
    ((te * ct.to(dev)).sum() + (ve * cv.to(dev)).sum()).backward()
    torch.cuda.synchronize()
    grads = {k: v.grad.detach().cpu() for k, v in params.items() if v.is_floating_point() and v.grad is not None}
    # video tower (smooth: GELU only): gated on its own bf16-vs-fp32 noise floor. Text tower: the ReLU in front of txt_proj
    # (oa_model.py:68) makes its gradient discontinuous - a rounding-level difference that flips the sign of a few of the
    # 768 CLS features changes ~0.5 % of the gradient energy, i.e. ~7 % relative error in every text tensor (seen in the
    # bf16-vs-fp32 oracle pair itself for one of the two seeds) - so it only gets an absolute gate here; its kernels are
    # gated tightly in test_ops_gpu.py and, without mask flips, by the cfg1 / golden cases above.
    def part(d, prefixes):
        return {k: v for k, v in d.items() if k.startswith(prefixes)}
    vid = ("video_model.", "vid_proj.")
    txt = ("text_model.", "txt_proj.")
    floor = summarize("noise_floor_%s_linear_video" % tag, s16, s32, 0.0, 0.0, part(g16, vid), part(g32, vid))
    floort = summarize("noise_floor_%s_linear_text" % tag, s16, s32, 0.0, 0.0, part(g16, txt), part(g32, txt))
    rep32 = summarize("%s_depth2_linear_video_vs_fp32_oracle" % tag, sims, s32, 0.0, 0.0, part(grads, vid), part(g32, vid))
    rep = summarize("%s_depth2_linear_video_vs_bf16_oracle" % tag, sims, s16, 0.0, 0.0, part(grads, vid), part(g16, vid))
    rept32 = summarize("%s_depth2_linear_text_vs_fp32_oracle" % tag, sims, s32, 0.0, 0.0, part(grads, txt), part(g32, txt))
    rept = summarize("%s_depth2_linear_text_vs_bf16_oracle" % tag, sims, s16, 0.0, 0.0, part(grads, txt), part(g16, txt))
    assert rep["logit_max_abs_err"] < LOGIT_TOL and rep32["logit_max_abs_err"] < LOGIT_TOL
    gate_grads(floor, vs_truth=rep32, vs_bf16_oracle=rep)
    gate_grads(floort, vs_truth=rept32, vs_bf16_oracle=rept)
    assert not rep["missing"] and not rept["missing"]


def test_cfg4_benchmark_geometry_full_depth_vs_oracles():
    """The configuration bench.py times (BASELINE configs[3] per-GPU shard: 8 frames of 224x224, 36 object regions per
    frame, 32-token text, ViT-B/16 space-time with all 12 blocks + DistilBERT-base, InfoNCE at T = 0.05) at batch 4:
    logits within 1e-3 of the fp32 oracle (= what the reference computes) and of the bf16 oracle; every parameter
    gradient compared per tensor with both oracles, gated on the measured operand-rounding floor (bf16 oracle vs fp32
    oracle on the same inputs)."""
    spec = dual_encoder_spec(frames=8, objects=True)
    w = fill_seeded(spec, 0, 0.02)
    g = torch.Generator().manual_seed(1234)
    B, Fr, Oo, L = 4, 8, 36, 32
    video = torch.randn(B, Fr, 3, 224, 224, generator=g)
    objects = O.synth_objects(B, Fr, Oo, g)
    text = O.synth_text(B, L, g)
    te, ve, sims, loss, grads = cuda_dual(w, video, text["input_ids"], text["attention_mask"], heads=12, objects=objects)
    torch.cuda.empty_cache()
    (s16, l16, g16), floor = noise_floor(w, video, text["input_ids"], text["attention_mask"], O.OracleCfg(bf16=True),
                                         O.OracleCfg(), objects=objects, tag="cfg4_")
    _, _, s32, l32, g32 = oracle_dual(w, video, text["input_ids"], text["attention_mask"], O.OracleCfg(), objects)
    rep16 = summarize("cfg4_vs_bf16_oracle", sims, s16, loss, l16, grads, g16)
    rep32 = summarize("cfg4_vs_fp32_oracle", sims, s32, loss, l32, grads, g32)
    assert not rep16["missing"]
    assert rep16["logit_max_abs_err"] < LOGIT_TOL and rep32["logit_max_abs_err"] < LOGIT_TOL
    assert abs(loss - l32) < 2e-3 * max(1.0, abs(l32))
    gate_grads(floor, vs_truth=rep32, vs_bf16_oracle=rep16)


def _region_model(g):
    from oa_transformer_b200.model.oa_model_region_mem import FrozenInTime
    m = FrozenInTime({"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224", "num_frames": 1,
                      "pretrained": True, "time_init": "rand", "allow_missing_vit": True, "img_size": 32, "depth": 6,
                      "embed_dim": 128, "num_heads": 2},
                     {"model": "", "input_objects": False},
                     {"model": "distilbert-base-uncased", "pretrained": True, "random_init": True,
                      "config": dict(dim=128, hidden_dim=256, n_heads=2, n_layers=2, vocab_size=200,
                                     max_position_embeddings=16)})
    w = fill_seeded(g["shapes"], g["weight_seed"], g["weight_scale"])
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not unexpected and all(k.startswith("video_model.head.") for k in missing), (missing, unexpected)
    return m.cuda(), w


def test_region_variant_vs_reference_golden():
    """the region-sensitive variant through the plugin surface (model.oa_model_region_mem.FrozenInTime,
    trainer.trainer_region_mem.region_loss) against OUTPUTS OF THE REFERENCE (tests/golden/region_small.pt: embeddings,
    region_sim, both loss terms, gradients), and against the oracle in bf16 mode for the floor."""
    from oa_transformer_b200.model import NormSoftmaxLoss, sim_matrix
    from oa_transformer_b200.trainer.trainer_region_mem import region_loss
    g = torch.load(os.path.join(GOLD, "region_small.pt"), map_location="cpu", weights_only=False)
    m, w = _region_model(g)
    m.eval()            # the fixture was generated with the reference in eval mode (DistilBERT dropout off)
    data = {"video": g["video"].cuda(), "text": {"input_ids": g["input_ids"].cuda(),
                                                  "attention_mask": g["attention_mask"].cuda()},
            "text_region_embedding": g["text_region_embedding"].cuda(), "patch_masks": g["patch_masks"].cuda()}
    te, ve, rs = m(data, aug=True)
    t2v = NormSoftmaxLoss(0.05)(sim_matrix(te, ve))
    rl = region_loss(rs, data["patch_masks"].squeeze(1).float())
    (t2v + rl).backward()
    torch.cuda.synchronize()
    grads = {k: v.grad.detach().cpu() for k, v in m.named_parameters() if v.grad is not None}
    sims = sim_matrix(te, ve).detach().cpu()
    from oracle import oracle as OO
    ref_sims = OO.sim_matrix(g["text_embeds"], g["video_embeds"])
    rep = summarize("region_vs_reference", sims, ref_sims, float(t2v + rl), float(g["loss"]), grads, g["grads_subset"])
    assert rep["logit_max_abs_err"] < LOGIT_TOL and not rep["missing"]
    # floor: the oracle in bf16 mode vs the reference on the same tensors
    p = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    odata = {"video": g["video"], "text": {"input_ids": g["input_ids"], "attention_mask": g["attention_mask"]},
             "text_region_embedding": g["text_region_embedding"]}
    ote, ove, ors = OO.region_mem_forward(odata, p, OO.OracleCfg(heads=2, text_layers=2, bf16=True))
    (OO.norm_softmax_loss(OO.sim_matrix(ote, ove)) + OO.region_loss(ors, g["patch_masks"].squeeze(1).float())).backward()
    og = {k: v.grad for k, v in p.items() if v.grad is not None and k in g["grads_subset"]}
    floor = summarize("noise_floor_region", OO.sim_matrix(ote, ove).detach(), ref_sims, 0.0, 0.0, og, g["grads_subset"])
    gate_grads(floor, vs_truth=rep)
    # region similarities (patch tokens and the 512 -> 256 projection run on plain bf16 operands): as close to the
    # reference as the bf16 oracle is
    rs_floor = float((ors.detach() - g["region_sim"]).abs().max())
    rs_err = float((rs.detach().cpu() - g["region_sim"]).abs().max())
    print(json.dumps({"case": "region_sim", "max_abs_err_vs_reference": rs_err, "bf16_oracle_floor": rs_floor,
                      "region_loss": float(rl), "region_loss_ref": float(g["region_loss"])}))
    assert rs_err < 1.5 * rs_floor + 1e-4
    assert abs(float(rl) - float(g["region_loss"])) < 5e-3 * float(g["region_loss"])
    # the unused parameters of the reference's forward stay without gradient contributions
    assert "video_model.region_norm.weight" in grads and "txt_proj_2.1.weight" in grads


def test_global_local_loss_head_vs_reference_golden():
    """trainer/trainer_global_local.py:187-208 (3-term InfoNCE with the mean-pooled fine-grained term) and the mask
    pooling einsum of model/oa_model_global_local.py:178, forward and backward, vs the reference-generated fixture."""
    from oa_transformer_b200.model import NormSoftmaxLoss
    from oa_transformer_b200.trainer.trainer_global_local import global_local_loss, pooled_region_features
    g = torch.load(os.path.join(GOLD, "global_local_loss.pt"), map_location="cpu", weights_only=False)
    names = ("text_embeds", "pad_text_embeds", "video_embeds", "region_feat", "tags_feat")
    t = {k: g[k].cuda().requires_grad_(True) for k in names}
    loss, terms = global_local_loss(NormSoftmaxLoss(0.05), t["text_embeds"], t["pad_text_embeds"], t["video_embeds"],
                                    t["region_feat"], t["tags_feat"])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])
    for k in ("st2sv", "lt2sv", "fine_grained"):
        assert abs(float(terms[k]) - float(g["terms"][k])) < 1e-4 * abs(float(g["terms"][k])), k
    for k in names:
        assert rel(t[k].grad.cpu(), g["grads"][k]) < 1e-3, (k, rel(t[k].grad.cpu(), g["grads"][k]))
    pf = g["patch_feats"].cuda().requires_grad_(True)
    pooled = pooled_region_features(g["patch_masks"].cuda(), pf)
    (pooled * g["pool_probe"].cuda()).sum().backward()
    assert rel(pooled.detach().cpu(), g["pooled"]) < 1e-5 and rel(pf.grad.cpu(), g["pool_grad"]) < 1e-5


def test_train_dist_region_mem_entry_point_synthetic(tmp_path, monkeypatch):
    """The region variant's entry point (train_dist_region_mem.py wiring: oa_model_region_mem + trainer_region_mem) for
    one tiny epoch with validation and a checkpoint."""
    import json as _json
    from oa_transformer_b200 import train_dist_region_mem
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = _json.load(open(os.path.join(root, "oa_transformer_b200", "configs", "pt", "cc3m_webvid",
                                       "synthetic-region-mem.json")))
    cfg["trainer"]["save_dir"] = str(tmp_path)
    cfg["trainer"]["save_period"] = 1
    cfg["data_loader"][0]["args"].update({"batch_size": 2, "n_samples": 4})
    cfg["arch"]["args"]["video_params"]["depth"] = 6
    cfg["trainer"]["max_samples_per_epoch"] = 4
    p = tmp_path / "cfg.json"
    p.write_text(_json.dumps(cfg))
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    train_dist_region_mem.main(["-c", str(p)])
    ckpts = list(tmp_path.rglob("checkpoint-epoch1.pth"))
    assert len(ckpts) == 1
    ck = torch.load(str(ckpts[0]), map_location="cpu", weights_only=False)
    assert "txt_proj_2.1.weight" in ck["state_dict"] and "video_model.region_norm.weight" in ck["state_dict"]


def test_text_tower_training_dropout_vs_oracle_with_same_masks():
    """DistilBERT's dropout stays active in training (oa_model.py:28). Full-size text tower,
    forward and backward, with the three dropout sites on, against the oracle given exactly the masks the kernels draw
    (ops.dropout_mask); and model.eval() / model.train() switch it through the module surface."""
    from oa_transformer_b200 import ops
    from oa_transformer_b200.engine import TextEngine
    from oa_transformer_b200.functional import run_tower
    from oracle.weights import text_tower_spec
    spec = text_tower_spec()
    w = fill_seeded(spec, 51, 0.02)
    g = torch.Generator().manual_seed(52)
    B, L, H, D, layers = 4, 32, 12, 768, 6
    text = O.synth_text(B, L, g, ragged=True)
    # the tower's own output (last_hidden_state[:, 0]) under a linear loss: the ReLU of txt_proj would add its mask-flip
    # noise (see the config-shaped test below) to what this test is about
    coef = torch.randn(B, D, generator=g)
    pd, pa, seed = 0.1, 0.1, 20261017
    drop = {}

    def mult(shape, p, site):
        n = 1
        for d in shape:
            n *= d
        return ops.dropout_mask(n, p, seed, site, "cuda").view(shape).float().cpu() / (1.0 - p)
    drop[0] = mult((B, L, D), pd, 0)
    for i in range(layers):
        drop[1 + 3 * i] = mult((B, H, L, L), pa, 1 + 3 * i)
        drop[2 + 3 * i] = mult((B, L, D), pd, 2 + 3 * i)

    def oracle_run(cfg, d):
        p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in w.items()}
        te = O.distilbert(text["input_ids"], text["attention_mask"], p, cfg, drop=d)[:, 0]
        (te * coef).sum().backward()
        return te.detach(), {k: v.grad for k, v in p.items() if v.is_floating_point() and v.grad is not None}

    t16, g16 = oracle_run(O.OracleCfg(bf16=True), drop)
    t32, g32 = oracle_run(O.OracleCfg(), drop)
    t_eval, _ = oracle_run(O.OracleCfg(), None)
    dev = torch.device("cuda")
    params = {k: v.to(dev).clone().requires_grad_(v.is_floating_point()) for k, v in w.items()}
    named = [(k, v) for k, v in params.items() if v.is_floating_point()]
    te = run_tower(TextEngine(dev, heads=H), named, input_ids=text["input_ids"].to(dev),
                   attention_mask=text["attention_mask"].to(dev), dropout={"p": pd, "p_attn": pa, "seed": seed},
                   proj=None)
    (te * coef.to(dev)).sum().backward()
    torch.cuda.synchronize()
    grads = {k: v.grad.detach().cpu() for k, v in params.items() if v.is_floating_point() and v.grad is not None}
    out = te.detach().cpu()
    assert rel(out, t32) < 2e-3 and rel(out, t16) < 2e-3, (rel(out, t32), rel(out, t16))
    assert rel(t32, t_eval) > 5e-2                       # the masks matter: training output != eval output
    floor = summarize("noise_floor_text_dropout", out, t16, 0.0, 0.0, g16, g32)
    rep32 = summarize("text_dropout_vs_fp32_oracle", out, t32, 0.0, 0.0, grads, g32)
    rep16 = summarize("text_dropout_vs_bf16_oracle", out, t16, 0.0, 0.0, grads, g16)
    gate_grads(floor, vs_truth=rep32, vs_bf16_oracle=rep16)
    assert not rep32["missing"]


def test_video_tower_cuda_graph_replay_matches_eager(monkeypatch):
    """engine._GraphedSchedule: the video tower's forward and backward replayed as CUDA graphs (third call on) give what
    the eager launches give (first call), with new input values copied into the same tensors and updated weights."""
    from oa_transformer_b200 import engine, ops
    from oa_transformer_b200.functional import run_tower
    from oracle.weights import video_tower_spec
    monkeypatch.setattr(engine, "GRAPH", True)
    spec = video_tower_spec(depth=2, frames=4, objects=True)
    spec["vid_proj.0.weight"], spec["vid_proj.0.bias"] = (256, 768), (256,)
    w = fill_seeded(spec, 61, 0.02)
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(62)
    B, Fr, Oo = 2, 4, 36
    vids = [torch.randn(B, Fr, 3, 224, 224, generator=g) for _ in range(2)]
    objs = [O.synth_objects(B, Fr, Oo, g) for _ in range(2)]
    coef = torch.randn(B, 256, generator=g).to(dev)

    def run(eng, params, video, objects):
        for p in params.values():
            p.grad = None
        named = list(params.items())
        ve = run_tower(eng, named, video=video, objects=objects)
        (ve * coef).sum().backward()
        torch.cuda.synchronize()
        return ve.detach().clone(), {k: v.grad.detach().clone() for k, v in params.items()}

    def fresh():
        return {k: v.to(dev).clone().requires_grad_(True) for k, v in w.items()}

    # reference: eager engine (graphs off), the two inputs, and a weight update in between
    monkeypatch.setattr(engine, "GRAPH", False)
    eng0, p0 = engine.VideoEngine(dev, heads=12), fresh()
    ref = []
    for i in range(2):
        ref.append(run(eng0, p0, vids[i].to(dev), objs[i].to(dev)))
        with torch.no_grad():
            for v in p0.values():
                v.mul_(1.01)
    monkeypatch.setattr(engine, "GRAPH", True)
    eng1, p1 = engine.VideoEngine(dev, heads=12), fresh()
    vbuf, obuf = vids[0].to(dev), objs[0].to(dev)
    run(eng1, p1, vbuf, obuf)                    # 1st call: eager
    run(eng1, p1, vbuf, obuf)                    # 2nd call: captured + replayed
    n_replays = ops.GRAPH_REPLAYS
    got = []
    for i in range(2):                           # later calls: replays, inputs refreshed in place, weights updated in place
        vbuf.copy_(vids[i])
        obuf.copy_(objs[i])
        got.append(run(eng1, p1, vbuf, obuf))
        with torch.no_grad():
            for v in p1.values():
                v.mul_(1.01)
    assert ops.GRAPH_REPLAYS - n_replays == 4            # forward + backward of both steps were graph replays
    for (ve_r, g_r), (ve_g, g_g) in zip(ref, got):
        assert rel(ve_g, ve_r) < 2e-3                    # split-K reductions: equal to rounding, not bit for bit
        for k in g_r:
            assert rel(g_g[k], g_r[k]) < 2e-2, (k, rel(g_g[k], g_r[k]))


def test_space_time_transformer_returns_patch_tokens_like_the_reference():
    """video_transformer.py:346-351 returns (x[:, 0], x[:, 1:]) after the final norm: the module mirror with
    return_tokens=True against the reference's own outputs and gradients (tests/golden/video_small.pt, whose loss mixes
    the CLS feature and the patch tokens)."""
    from oa_transformer_b200.model import SpaceTimeTransformer
    g = torch.load(os.path.join(GOLD, "video_small.pt"), map_location="cpu", weights_only=False)
    m = SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=3,
                             time_init="rand")
    m.head = torch.nn.Identity()
    missing, unexpected = m.load_state_dict(g["weights"], strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    m = m.cuda()
    cls, tokens = m(g["video"].cuda(), return_tokens=True)
    # (this fixture's weights are 4x trained scale: bf16 operands put ~5e-3 on the raw 128-d features)
    assert rel(cls.detach().cpu(), g["cls"]) < 1e-2 and rel(tokens.detach().cpu(), g["tokens"]) < 1e-2
    ((cls * g["probe"].cuda()).sum() + (tokens * g["probe_tokens"].cuda()).sum()).backward()
    torch.cuda.synchronize()
    errs = {k: rel(dict(m.named_parameters())[k].grad.cpu(), v) for k, v in g["grads"].items()}
    worst = max(errs.values())
    assert worst < 3e-2, sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    cls2, none = m(g["video_short"].cuda())
    assert none is None and rel(cls2.detach().cpu(), g["cls_short"]) < 1e-2


def test_frozen_in_time_module_surface():
    """The nn.Module mirror: constructor, forward(data) -> (text, video) embeddings, backward into .grad."""
    from oa_transformer_b200.model import FrozenInTime, NormSoftmaxLoss, sim_matrix
    torch.manual_seed(0)
    m = FrozenInTime({"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224", "num_frames": 2,
                      "pretrained": True, "time_init": "zeros", "allow_missing_vit": True},
                     {"model": "", "input_objects": False},
                     {"model": "distilbert-base-uncased", "pretrained": True, "random_init": True}).cuda()
    g = torch.Generator().manual_seed(1)
    data = {"video": torch.randn(2, 2, 3, 224, 224, generator=g).cuda(),
            "text": {k: v.cuda() for k, v in O.synth_text(2, 8, g).items()}}
    te, ve = m(data, aug=True)
    assert te.shape == (2, 256) and ve.shape == (2, 256)
    loss = NormSoftmaxLoss(0.05)(sim_matrix(te, ve))
    loss.backward()
    assert m.vid_proj[0].weight.grad is not None and m.text_model.embeddings.word_embeddings.weight.grad is not None
    assert torch.isfinite(loss)
    with torch.no_grad():
        s = m(data, return_embeds=False)
    assert s.shape == (2, 2)
    # training mode keeps DistilBERT's dropout on (oa_model.py:28): two calls differ, the video tower (no dropout) does
    # not; eval() makes the text embeddings deterministic again
    with torch.no_grad():
        t1, v1 = m(data)
        t2, v2 = m(data)
        # (the video tower has no dropout; its split-K reductions make it reproducible to rounding, not bit for bit)
        assert float((t1 - t2).abs().max()) > 1e-2 and float((v1 - v2).abs().max()) < 5e-3
        m.eval()
        t3, _ = m(data)
        t4, _ = m(data)
        assert torch.equal(t3, t4)
        m.train()


def test_train_dist_multi_entry_point_synthetic(tmp_path, monkeypatch):
    """The plugin surface end to end on one GPU: ConfigParser -> FrozenInTime / NormSoftmaxLoss / loaders ->
    Multi_Trainer_dist._train_epoch (model, AllGather_multi, sim_matrix, loss, backward, AdamW step) -> validation
    metrics -> checkpoint with the reference's dictionary layout."""
    import json as _json
    from oa_transformer_b200 import train_dist_multi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = _json.load(open(os.path.join(root, "oa_transformer_b200", "configs", "pt", "cc3m_webvid",
                                       "synthetic-objects.json")))
    cfg["trainer"]["save_dir"] = str(tmp_path)
    cfg["trainer"]["save_period"] = 1
    cfg["data_loader"][0]["args"].update({"batch_size": 2, "n_samples": 4, "num_objects": 4})
    cfg["data_loader"][0]["args"]["video_params"]["num_frames"] = 2
    cfg["arch"]["args"]["video_params"]["num_frames"] = 2
    cfg["trainer"]["max_samples_per_epoch"] = 4
    p = tmp_path / "cfg.json"
    p.write_text(_json.dumps(cfg))
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    train_dist_multi.main(["-c", str(p)])
    ckpts = list(tmp_path.rglob("checkpoint-epoch1.pth"))
    assert len(ckpts) == 1
    ck = torch.load(str(ckpts[0]), map_location="cpu", weights_only=False)
    assert set(ck.keys()) == {"arch", "epoch", "state_dict", "optimizer", "monitor_best", "config"}
    assert ck["arch"] == "FrozenInTime" and "video_model.object_embed.weight" in ck["state_dict"]
