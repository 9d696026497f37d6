"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.pt by RUNNING THE UNMODIFIED REFERENCE in the authoring
container (python oracle/make_golden.py). The fixtures hold seeded inputs, the weights (small models) or the weight
seed (full-size cfg1), and the reference's outputs / gradients. tests/test_oracle_vs_golden.py pins oracle/oracle.py
against them on any machine; the GPU parity tests then compare the CUDA path with the pinned oracle.

Reference entry points exercised (paths relative to /root/reference/OATrans):
  model/video_transformer.py:SpaceTimeTransformer            -> video_small.pt, cfg1_full.pt
  model/oa_model.py:FrozenInTime (+ HF DistilBertModel)      -> dual_small.pt, cfg1_full.pt
  model/model.py:sim_matrix, model/loss.py:NormSoftmaxLoss   -> loss.pt (+ every dual fixture)
  trainer/trainer_dist.py:AllGather_multi (2-rank gloo)      -> allgather2.pt
  base/base_dataset_global_local.py:patch_all_masks_from_bbox-> patch_masks.pt
  model/metric.py:t2v_metrics / v2t_metrics                  -> metrics.pt
  model/oa_model_region_mem.py:FrozenInTime over model/oa_video_transformer_region.py:SpaceTimeTransformer,
  trainer/trainer_region_mem.py:157-167 (region BCE loss)    -> region_small.pt
  trainer/trainer_global_local.py:187-208 (3-term loss), model/oa_model_global_local.py:178 -> global_local_loss.pt
  base/base_dataset_region_mem.py:233-247, base/base_dataset_global_local.py:395-405,
  base/base_dataset.py:593-650 (bbox / tag / region-feature bookkeeping)                    -> bookkeeping.pt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_shim  # noqa: E402

torch = ref_shim.install()
import numpy as np  # noqa: E402
from torch import nn  # noqa: E402


def seeded_weights(state_dict, seed, scale=0.02):
    """Deterministic weights by parameter name order: N(0, scale) for matrices / embeddings / biases,
    1 + N(0, 0.1) for LayerNorm gains. Shared with the tests (oracle/weights.py re-implements it without the reference)."""
    from oracle.weights import fill_seeded
    return fill_seeded(state_dict, seed, scale)


def grads_of(model, loss):
    model.zero_grad()
    loss.backward()
    return {k: v.grad.detach().clone() for k, v in model.named_parameters() if v.grad is not None}


def make_video_small():
    from OATrans.model.video_transformer import SpaceTimeTransformer
    torch.manual_seed(1)
    m = SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=3,
                             time_init="rand")
    m.head = nn.Identity()
    m.pre_logits = nn.Identity()
    sd = seeded_weights(m.state_dict(), seed=11, scale=0.08)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(12)
    video = torch.randn(3, 3, 3, 32, 32, generator=g)
    video2 = torch.randn(2, 2, 3, 32, 32, generator=g)      # fewer frames than num_frames (:73, :323-324)
    cls, tokens = m(video)
    probe = torch.randn(cls.shape, generator=g)
    probe_t = torch.randn(tokens.shape, generator=g) * 0.1
    loss = (cls * probe).sum() + (tokens * probe_t).sum()
    grads = grads_of(m, loss)
    cls2, tokens2 = m(video2)
    torch.save({"weights": sd, "video": video, "cls": cls.detach(), "tokens": tokens.detach(), "probe": probe,
                "probe_tokens": probe_t, "grads": grads, "video_short": video2, "cls_short": cls2.detach(),
                "cfg": {"heads": 2, "img": 32}}, os.path.join(GOLD, "video_small.pt"))
    print("video_small ok", cls.shape, tokens.shape)


def _build_frozen(num_frames, video_model, text_cfg_kwargs, cwd_seed):
    """FrozenInTime built through its own constructor (oa_model.py:11-92) with a scratch pretrained/ dir, then the
    video tower swapped for the requested geometry (FrozenInTime never forwards img_size)."""
    from transformers import DistilBertConfig, DistilBertModel
    from OATrans.model.oa_model import FrozenInTime
    d = ref_shim.scratch_cwd(cwd_seed)
    if text_cfg_kwargs:
        torch.manual_seed(cwd_seed)
        DistilBertModel(DistilBertConfig(**text_cfg_kwargs)).save_pretrained(
            os.path.join(d, "pretrained", "distilbert-base-uncased"))
    m = FrozenInTime(
        video_params={"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224", "num_frames": num_frames,
                      "pretrained": True, "time_init": "zeros"},
        object_params={"model": "", "input_objects": False},
        text_params={"model": "pretrained/distilbert-base-uncased", "pretrained": True, "input": "text"},
        projection_dim=256, projection="minimal")
    if video_model is not None:
        video_model.head = nn.Identity()
        video_model.pre_logits = nn.Identity()
        video_model.fc = nn.Identity()
        ftr = video_model.embed_dim
        m.video_model = video_model
        m.vid_proj = nn.Sequential(nn.Linear(ftr, 256))
    m.eval()   # DistilBERT dropout off; video tower has no dropout
    return m


def _run_dual(m, data, T=0.05):
    from OATrans.model.model import sim_matrix
    from OATrans.model.loss import NormSoftmaxLoss
    text_e, video_e = m(data)                       # FrozenInTime.forward, oa_model.py:97-104
    sims = sim_matrix(text_e, video_e)              # trainer_dist.py:161
    loss = NormSoftmaxLoss(T)(sims)                 # trainer_dist.py:162
    grads = grads_of(m, loss)
    return text_e.detach(), video_e.detach(), sims.detach(), loss.detach(), grads


def make_dual_small():
    from OATrans.model.video_transformer import SpaceTimeTransformer
    torch.manual_seed(2)
    vm = SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=2,
                              time_init="rand")
    m = _build_frozen(2, vm, dict(dim=128, hidden_dim=256, n_heads=2, n_layers=2, vocab_size=200,
                                  max_position_embeddings=16), cwd_seed=3)
    sd = seeded_weights(m.state_dict(), seed=21, scale=0.08)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(22)
    video = torch.randn(4, 2, 3, 32, 32, generator=g)
    ids = torch.randint(5, 200, (4, 8), generator=g)
    mask = torch.ones(4, 8, dtype=torch.long)
    mask[1, 5:] = 0
    mask[3, 3:] = 0
    ids = ids * mask
    data = {"video": video, "text": {"input_ids": ids, "attention_mask": mask}}
    te, ve, sims, loss, grads = _run_dual(m, data)
    torch.save({"weights": sd, "video": video, "input_ids": ids, "attention_mask": mask, "text_embeds": te,
                "video_embeds": ve, "sims": sims, "loss": loss, "grads": grads,
                "cfg": {"heads": 2, "text_layers": 2}}, os.path.join(GOLD, "dual_small.pt"))
    print("dual_small ok loss", float(loss))


def make_cfg1_full():
    """BASELINE.json configs[0]: 1-clip 2-frame 64x64, 0 objects, 8-token text, batch 4, full-size ViT-B/16 +
    DistilBERT-base. Weights are too large to commit: they are regenerated from seed 31 by name (oracle/weights.py)."""
    from OATrans.model.video_transformer import SpaceTimeTransformer
    torch.manual_seed(4)
    vm = SpaceTimeTransformer(img_size=64, num_frames=2, time_init="rand")
    m = _build_frozen(2, vm, None, cwd_seed=5)
    sd = seeded_weights(m.state_dict(), seed=31, scale=0.02)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(32)
    video = torch.randn(4, 2, 3, 64, 64, generator=g)
    ids = torch.randint(1000, 30522, (4, 8), generator=g)
    ids[:, 0] = 101
    ids[:, -1] = 102
    mask = torch.ones(4, 8, dtype=torch.long)
    data = {"video": video, "text": {"input_ids": ids, "attention_mask": mask}}
    te, ve, sims, loss, grads = _run_dual(m, data)
    keep = ["vid_proj.0.weight", "txt_proj.1.bias", "video_model.blocks.0.attn.qkv.bias",
            "video_model.blocks.11.timeattn.proj.bias", "video_model.temporal_embed", "video_model.cls_token",
            "video_model.blocks.5.norm3.weight", "text_model.transformer.layer.0.attention.q_lin.bias",
            "text_model.embeddings.LayerNorm.weight", "video_model.patch_embed.proj.bias"]
    small = {k: grads[k] for k in keep}
    norms = {k: float(v.norm()) for k, v in grads.items()}
    torch.save({"weight_seed": 31, "weight_scale": 0.02, "names": list(sd.keys()),
                "shapes": {k: tuple(v.shape) for k, v in sd.items()}, "video": video, "input_ids": ids,
                "attention_mask": mask, "text_embeds": te, "video_embeds": ve, "sims": sims, "loss": loss,
                "grads_subset": small, "grad_norms": norms}, os.path.join(GOLD, "cfg1_full.pt"))
    print("cfg1_full ok loss", float(loss))


def make_loss():
    from OATrans.model.model import sim_matrix
    from OATrans.model.loss import NormSoftmaxLoss
    g = torch.Generator().manual_seed(41)
    out = {}
    for name, (n, dim) in {"b8": (8, 256), "b33": (33, 256), "b64": (64, 256)}.items():
        a = torch.randn(n, dim, generator=g, requires_grad=True)
        b = torch.randn(n, dim, generator=g, requires_grad=True)
        with torch.no_grad():
            if name == "b8":
                a[2].zero_()          # exercises the eps clamp of sim_matrix (model.py:168-170)
        sims = sim_matrix(a, b)
        loss = NormSoftmaxLoss(0.05)(sims)
        ga, gb = torch.autograd.grad(loss, [a, b])
        out[name] = {"a": a.detach(), "b": b.detach(), "sims": sims.detach(), "loss": loss.detach(), "ga": ga, "gb": gb}
    torch.save(out, os.path.join(GOLD, "loss.pt"))
    print("loss ok")


def _ag_worker(rank, world, port, B, ret):
    import torch.distributed as dist
    from types import SimpleNamespace
    torch = ref_shim.install()
    from OATrans.trainer.trainer_dist import AllGather_multi
    from OATrans.model.model import sim_matrix
    from OATrans.model.loss import NormSoftmaxLoss
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(51)
    text = torch.randn(world * B, 256, generator=g)
    video = torch.randn(world * B, 256, generator=g)
    t = text[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    v = video[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    args = SimpleNamespace(rank=rank, world_size=world, local_rank=rank)
    vg = AllGather_multi.apply(v, world, args)          # trainer_dist.py:159-160
    tg = AllGather_multi.apply(t, world, args)
    loss = NormSoftmaxLoss(0.05)(sim_matrix(tg, vg))
    loss.backward()
    ret[rank] = {"t_grad": t.grad.clone(), "v_grad": v.grad.clone(), "loss": loss.detach().clone()}
    if rank == 0:
        ret["text"], ret["video"] = text, video
    dist.destroy_process_group()


def make_allgather():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ag_worker, args=(2, 29611, 3, ret), nprocs=2, join=True)
    out = {"text": ret["text"], "video": ret["video"], "B": 3, "world": 2,
           "ranks": [dict(ret[0]), dict(ret[1])]}
    torch.save(out, os.path.join(GOLD, "allgather2.pt"))
    print("allgather2 ok")


def make_patch_masks():
    import importlib
    mod = importlib.import_module("OATrans.base.base_dataset_global_local")
    fn = None
    for name in dir(mod):
        obj = getattr(mod, name)
        if isinstance(obj, type) and "patch_all_masks_from_bbox" in obj.__dict__:
            fn = obj.__dict__["patch_all_masks_from_bbox"]
            break
    assert fn is not None
    rng = np.random.RandomState(61)
    boxes = []
    for _ in range(64):
        x1, y1 = rng.uniform(0, 0.7, 2)
        w, h = rng.uniform(0.05, 0.3, 2)
        boxes.append([x1, y1, x1 + w, y1 + h, w, h])
    # edge cases: full image, zero-area, exactly on grid lines
    boxes += [[0, 0, 1, 1, 1, 1], [0.5, 0.5, 0.5, 0.5, 0, 0], [2 / 14, 3 / 14, 5 / 14, 7 / 14, 3 / 14, 4 / 14],
              [0.999, 0.999, 1.0, 1.0, 0.001, 0.001]]
    boxes = np.array(boxes, dtype=np.float64)
    masks = fn(None, boxes.copy())
    torch.save({"boxes": torch.from_numpy(boxes), "masks": torch.from_numpy(masks)},
               os.path.join(GOLD, "patch_masks.pt"))
    print("patch_masks ok", masks.shape, masks.sum())


def make_metrics():
    metric = ref_shim.load_file_module("ref_metric", "OATrans/model/metric.py")
    rng = np.random.RandomState(71)
    out = {}
    for name, n in (("n50", 50), ("n200", 200)):
        sims = rng.randn(n, n).astype(np.float32)
        sims[np.arange(n), np.arange(n)] += 1.5
        if name == "n50":
            sims[3, 7] = sims[3, 3]       # tie on a text query
            sims[9, 4] = sims[4, 4]       # tie on a video query
            sims[20, :] = 0.25            # constant row
        t2v = {k: float(v) for k, v in metric.t2v_metrics(sims.copy()).items()}
        v2t = {k: float(v) for k, v in metric.v2t_metrics(sims.copy()).items()}
        out[name] = {"sims": torch.from_numpy(sims), "t2v": t2v, "v2t": v2t}
    torch.save(out, os.path.join(GOLD, "metrics.pt"))
    print("metrics ok", out["n50"]["t2v"])


def make_region_small():
    """The region-sensitive variant end to end through the reference: model/oa_model_region_mem.py:FrozenInTime.forward
    (:105-123: anchor/video clip split, vid_proj on CLS and region features, txt_proj_2, mean-pool blend, region_sim)
    over model/oa_video_transformer_region.py:SpaceTimeTransformer (region_norm at layer 6), then the loss of
    trainer/trainer_region_mem.py:157-167 (NormSoftmaxLoss + 0.1 * BCELoss(sum) / rows)."""
    from transformers import DistilBertConfig, DistilBertModel
    if not hasattr(nn.init, "xavier_uniform"):            # deprecated alias the reference still calls (:13)
        nn.init.xavier_uniform = nn.init.xavier_uniform_
    from OATrans.model.oa_model_region_mem import FrozenInTime
    from OATrans.model.oa_video_transformer_region import SpaceTimeTransformer
    from OATrans.model.model import sim_matrix
    from OATrans.model.loss import NormSoftmaxLoss
    d = ref_shim.scratch_cwd(6)
    torch.manual_seed(6)
    DistilBertModel(DistilBertConfig(dim=128, hidden_dim=256, n_heads=2, n_layers=2, vocab_size=200,
                                     max_position_embeddings=16)).save_pretrained(
        os.path.join(d, "pretrained", "distilbert-base-uncased"))
    m = FrozenInTime(
        video_params={"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224", "num_frames": 1,
                      "pretrained": True, "time_init": "zeros"},
        object_params={"model": "", "input_objects": False},
        text_params={"model": "pretrained/distilbert-base-uncased", "pretrained": True, "input": "text"},
        projection_dim=256, projection="minimal")
    vm = SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=128, depth=6, num_heads=2, num_frames=1,
                              time_init="rand")
    vm.head = nn.Identity()
    vm.pre_logits = nn.Identity()
    vm.fc = nn.Identity()
    m.video_model = vm
    m.vid_proj = nn.Sequential(nn.Linear(128, 256))
    m.eval()
    sd = seeded_weights(m.state_dict(), seed=81, scale=0.08)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(82)
    B, K, L = 4, 5, 4
    video = torch.randn(B, 2, 3, 32, 32, generator=g)          # frame 0 = anchor image, frame 1 = the (1-frame) clip
    ids = torch.randint(5, 200, (B, 8), generator=g)
    mask = torch.ones(B, 8, dtype=torch.long)
    mask[2, 6:] = 0
    ids = ids * mask
    tre = 0.05 * torch.randn(B, K, 512, generator=g)          # keeps the region logits of O(1): unsaturated sigmoids
    patch_masks = (torch.rand(B, 1, K, L, generator=g) > 0.5).double()     # float64-from-numpy in the loaders
    data = {"video": video, "text": {"input_ids": ids, "attention_mask": mask}, "text_region_embedding": tre,
            "patch_masks": patch_masks}
    text_e, video_e, region_sim = m(data, aug=True)
    output = sim_matrix(text_e, video_e)
    loss = NormSoftmaxLoss(0.05)(output)
    pm = patch_masks.squeeze(1).float()
    rs, pmv = region_sim.view(-1, region_sim.size(-1)), pm.view(-1, pm.size(-1))
    r_loss = 0.1 * nn.BCELoss(reduction="sum")(rs, pmv) / rs.size(0)
    total = loss + r_loss
    grads = grads_of(m, total)
    # weights are regenerated from the seed by name (oracle/weights.py); every 1-D gradient and a spread of matrices
    # are kept, the norm of every gradient beside them
    keep2d = ("vid_proj.0.weight", "txt_proj.1.weight", "txt_proj_2.1.weight", "video_model.blocks.0.attn.qkv.weight",
              "video_model.blocks.5.timeattn.proj.weight", "video_model.blocks.3.mlp.fc1.weight",
              "video_model.patch_embed.proj.weight", "video_model.pos_embed", "video_model.cls_token",
              "text_model.transformer.layer.0.attention.q_lin.weight", "text_model.embeddings.word_embeddings.weight")
    small = {k: v for k, v in grads.items() if v.dim() == 1 or k in keep2d}
    torch.save({"weight_seed": 81, "weight_scale": 0.08, "shapes": {k: tuple(v.shape) for k, v in sd.items()},
                "video": video, "input_ids": ids, "attention_mask": mask, "text_region_embedding": tre,
                "patch_masks": patch_masks, "text_embeds": text_e.detach(), "video_embeds": video_e.detach(),
                "region_sim": region_sim.detach(), "loss": total.detach(), "t2v_loss": loss.detach(),
                "region_loss": r_loss.detach(), "grads_subset": small,
                "grad_norms": {k: float(v.norm()) for k, v in grads.items()},
                "cfg": {"heads": 2, "text_layers": 2}}, os.path.join(GOLD, "region_small.pt"))
    print("region_small ok loss", float(total), "region", float(r_loss), "grads", len(grads))


def make_global_local_loss():
    """The loss arithmetic of trainer/trainer_global_local.py:187-208 (the trainer's model cannot be constructed - SURVEY
    fact 3 - so the three terms are evaluated with the reference's sim_matrix / NormSoftmaxLoss on seeded features),
    plus the mask pooling einsum of model/oa_model_global_local.py:178."""
    from OATrans.model.model import sim_matrix
    from OATrans.model.loss import NormSoftmaxLoss
    g = torch.Generator().manual_seed(91)
    B, R, P, L = 4, 8, 256, 196
    names = ("text_embeds", "pad_text_embeds", "video_embeds")
    t = {k: torch.randn(B, P, generator=g, requires_grad=True) for k in names}
    region_feat = torch.randn(B, R, P, generator=g, requires_grad=True)
    tags_feat = torch.randn(B, R, P, generator=g, requires_grad=True)
    loss_fn = NormSoftmaxLoss(0.05)
    st2sv = loss_fn(sim_matrix(t["text_embeds"], t["video_embeds"]))
    lt2sv = loss_fn(sim_matrix(t["pad_text_embeds"], t["video_embeds"]))
    fine = loss_fn(sim_matrix(torch.mean(region_feat, dim=1), torch.mean(tags_feat, dim=1)))
    loss = st2sv + lt2sv + fine
    gs = torch.autograd.grad(loss, [t[k] for k in names] + [region_feat, tags_feat])
    out = {k: v.detach() for k, v in t.items()}
    out.update(region_feat=region_feat.detach(), tags_feat=tags_feat.detach(), loss=loss.detach(),
               terms={"st2sv": st2sv.detach(), "lt2sv": lt2sv.detach(), "fine_grained": fine.detach()},
               grads={k: gv for k, gv in zip(names + ("region_feat", "tags_feat"), gs)})
    patch_masks = (torch.rand(B, R, L, generator=g) > 0.8).float()
    patch_feats = torch.randn(B, L, P, generator=g, requires_grad=True)
    pooled = torch.einsum('b o l, b l c -> b o c', patch_masks, patch_feats)
    probe = torch.randn(B, R, P, generator=g)
    (gp,) = torch.autograd.grad((pooled * probe).sum(), [patch_feats])
    out.update(patch_masks=patch_masks, patch_feats=patch_feats.detach(), pooled=pooled.detach(), pool_probe=probe,
               pool_grad=gp)
    torch.save(out, os.path.join(GOLD, "global_local_loss.pt"))
    print("global_local_loss ok", float(loss))


def make_bookkeeping():
    """Integer / index bookkeeping in front of the path, outputs of the reference's own functions:
      base/base_dataset_region_mem.py:233-247  patch_all_masks_from_bbox (random pick of 5 boxes, same-class union)
      base/base_dataset_global_local.py:395-405 object_tags_masks (running end offsets of the tag tokens)
      base/base_dataset.py:593-650              read_object_from_disk (confidence order, v=2 class de-duplication,
                                                'edge' padding to top_k, box geometry scaled to [0, 1])"""
    import importlib
    import random
    import tempfile
    from types import SimpleNamespace
    out = {}
    # ---- region_mem masks
    mod = importlib.import_module("OATrans.base.base_dataset_region_mem")
    fn = None
    for name in dir(mod):
        obj = getattr(mod, name)
        if isinstance(obj, type) and "patch_all_masks_from_bbox" in obj.__dict__:
            fn = obj.__dict__["patch_all_masks_from_bbox"]
            break
    assert fn is not None
    rng = np.random.RandomState(101)
    cases = []
    for case in range(6):
        n = int(rng.randint(5, 21))
        xy = rng.uniform(0, 0.7, (n, 2))
        wh = rng.uniform(0.05, 0.3, (n, 2))
        boxes = np.concatenate([xy, xy + wh, wh], axis=1).astype(np.float64)
        if case == 0:
            boxes[0, :4] = [0, 0, 1, 1]
            boxes[1, :4] = [3 / 14, 2 / 14, 6 / 14, 9 / 14]
        classes = [int(c) for c in rng.randint(0, 6, n)]       # few classes -> same-class unions happen
        random.seed(200 + case)
        indexs = random.sample(range(0, n), 5)
        random.seed(200 + case)
        masks, sel = fn(None, boxes.copy(), list(classes))
        cases.append({"boxes": torch.from_numpy(boxes), "classes": torch.tensor(classes, dtype=torch.int32),
                      "indexs": torch.tensor(indexs, dtype=torch.int32), "masks": torch.from_numpy(masks),
                      "sel_objects": [int(x) for x in sel]})
    out["region_mem_masks"] = cases
    # ---- object_tags_masks
    gl = importlib.import_module("OATrans.base.base_dataset_global_local")
    otm = None
    for name in dir(gl):
        obj = getattr(gl, name)
        if isinstance(obj, type) and "object_tags_masks" in obj.__dict__:
            otm = obj.__dict__["object_tags_masks"]
            break
    assert otm is not None
    lens = rng.randint(1, 5, 1601).astype(np.float64)           # np.loadtxt gives float64 (:279)
    tags = []
    for k in (1, 7, 20):
        idx = [int(i) for i in rng.randint(0, 1601, k)]
        mask, total = otm(SimpleNamespace(object_token_lens=lens), idx)
        tags.append({"indices": torch.tensor(idx, dtype=torch.int64), "mask": mask.clone(), "total": int(total)})
    out["object_tags"] = {"lens": torch.from_numpy(lens), "cases": tags}
    # ---- region features from an extractor .npz
    bd = importlib.import_module("OATrans.base.base_dataset")
    feats = []
    tmp = tempfile.mkdtemp(prefix="oat_npz_")
    for case, (n, top_k, v) in enumerate([(20, 10, 1), (6, 10, 1), (20, 10, 2), (12, 36, 1), (9, 10, 2)]):
        x = np.abs(rng.randn(n, 2048)).astype(np.float32)
        w, h = int(rng.randint(200, 800)), int(rng.randint(200, 800))
        x1 = rng.uniform(0, 0.7 * w, n)
        y1 = rng.uniform(0, 0.7 * h, n)
        bbox = np.stack([x1, y1, x1 + rng.uniform(10, 0.3 * w, n), y1 + rng.uniform(10, 0.3 * h, n)], axis=1).astype(np.float32)
        conf = rng.permutation(n).astype(np.float32) / n + 0.01            # distinct confidences
        ids = rng.randint(0, 8, n).astype(np.int64)
        path = os.path.join(tmp, "%d.npz" % case)
        np.savez(path, x=x, bbox=bbox, info={"objects_conf": conf, "objects_id": ids, "image_w": w, "image_h": h})
        feat = bd.read_object_from_disk(path, top_k=top_k, v=v)
        feats.append({"x": torch.from_numpy(x), "bbox": torch.from_numpy(bbox), "conf": torch.from_numpy(conf),
                      "ids": torch.from_numpy(ids), "image_w": w, "image_h": h, "top_k": top_k, "v": v,
                      "feat": feat.clone()})
        print("  read_object_from_disk n=%d top_k=%d v=%d ->" % (n, top_k, v), tuple(feat.shape), feat.dtype)
    out["region_features"] = feats
    torch.save(out, os.path.join(GOLD, "bookkeeping.pt"))
    print("bookkeeping ok")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    assert ref_shim.available(), "reference not mounted"
    which = sys.argv[1:] or ["video_small", "dual_small", "cfg1_full", "loss", "allgather", "patch_masks", "metrics"]
    for w in which:
        globals()["make_" + w]()
