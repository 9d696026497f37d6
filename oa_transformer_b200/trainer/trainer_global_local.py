"""The loss of OATrans/trainer/trainer_global_local.py:171-211 on liboat: six AllGather_multi
(video, pad-text, pad-video, text embeddings, region features (B, R, P), tag features (B, R, P)), then

    loss = NormSoftmaxLoss(sim_matrix(text, video))                                   short text -> short video  (:187-188)
         + NormSoftmaxLoss(sim_matrix(pad_text, video))                               tag-padded text -> video   (:190,198)
         + NormSoftmaxLoss(sim_matrix(mean(region_feat, 1), mean(tags_feat, 1)))      fine-grained term          (:207-208)

The global-local MODEL cannot be constructed in the reference (CrossModalityFusion / SpaceTimeObjectTransformer are
referenced but never defined: model/oa_model_global_local.py:38,40,143 - SURVEY.md fact 3), so what is on the path is
this loss head; the mask pooling that produces `region_feat` (einsum 'b o l, b l c -> b o c', ibid. :178) is
functional.object_patch_attention(mode="mask")."""
from .. import functional as OF
from ..model.model import sim_matrix
from .trainer_dist import AllGather_multi


def global_local_loss(loss_fn, text_embeds, pad_text_embeds, video_embeds, region_feat, tags_feat, n_gpu=1, args=None,
                      pad_video_embeds=None):
    """Returns (loss, dict of the three terms). `args` needs .rank / .world_size when n_gpu > 1."""
    if args is not None and getattr(args, "world_size", 1) > 1:
        gather = lambda t: AllGather_multi.apply(t, n_gpu, args)                       # noqa: E731
        video_embeds, pad_text_embeds, text_embeds = gather(video_embeds), gather(pad_text_embeds), gather(text_embeds)
        if pad_video_embeds is not None:
            pad_video_embeds = gather(pad_video_embeds)                                 # gathered, unused (:173,191)
        region_feat, tags_feat = gather(region_feat), gather(tags_feat)
    st2sv = loss_fn(sim_matrix(text_embeds, video_embeds))
    lt2sv = loss_fn(sim_matrix(pad_text_embeds, video_embeds))
    fine = loss_fn(sim_matrix(OF.token_pool(None, region_feat, 0.0, 1.0), OF.token_pool(None, tags_feat, 0.0, 1.0)))
    return st2sv + lt2sv + fine, {"st2sv": st2sv, "lt2sv": lt2sv, "fine_grained": fine}


def pooled_region_features(patch_masks, patch_feats):
    """region_feat = einsum('b o l, b l c -> b o c', patch_masks, patch_feats) (model/oa_model_global_local.py:178)."""
    _, out = OF.object_patch_attention(None, None, patch_feats, mode="mask", masks=patch_masks.float())
    return out
