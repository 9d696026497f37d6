"""Retrieval metrics on a similarity matrix, same definitions and tie rules as OATrans/model/metric.py
(t2v_metrics :16-121 breaks ties optimistically, v2t_metrics :123-212 averages tied ranks, cols2metrics :281-291).
numpy on the host, as in the reference (`_valid_epoch` moves the embeddings to the CPU first,
trainer/trainer_dist.py:237-264). Square matrices (one caption per video) and the query_masks-free path only.
A CUDA tensor takes the device path instead: the O(n^2) rank-of-the-diagonal counting runs in `oat_retrieval_ranks`
(same tie rules, bit-identical ranks) and only the n ranks come back to the host for cols2metrics."""
import numpy as np


def _is_cuda_tensor(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def _device_ranks(sims):
    from .. import ops
    return ops.retrieval_ranks(sims.detach().float().contiguous())


def cols2metrics(cols, num_queries):
    metrics = {
        "R1": 100 * float(np.sum(cols == 0)) / num_queries,
        "R5": 100 * float(np.sum(cols < 5)) / num_queries,
        "R10": 100 * float(np.sum(cols < 10)) / num_queries,
        "R50": 100 * float(np.sum(cols < 50)) / num_queries,
        "MedR": np.median(cols) + 1,
        "MeanR": np.mean(cols) + 1,
    }
    stats = np.array([metrics["R1"], metrics["R5"], metrics["R10"]], dtype=np.float64)
    metrics["geometric_mean_R1-R5-R10"] = float(np.exp(np.log(stats).mean())) if np.all(stats > 0) else 0.0
    return metrics


def t2v_metrics(sims, query_masks=None):
    assert query_masks is None, "query_masks path is outside the hot-path scope"
    if _is_cuda_tensor(sims):
        assert sims.dim() == 2 and sims.shape[0] == sims.shape[1], "one caption per video expected"
        return cols2metrics(_device_ranks(sims)[0].cpu().numpy().astype(np.int64), sims.shape[0])
    sims = np.asarray(sims)
    assert sims.ndim == 2, "expected a matrix"
    num_queries, num_vids = sims.shape
    assert num_queries == num_vids, "one caption per video expected"
    dists = -sims
    sorted_dists = np.sort(dists, axis=1)
    gt = np.diag(dists)[:, np.newaxis]
    rows, cols = np.where((sorted_dists - gt) == 0)
    if rows.size > num_queries:                       # ties: keep the best (first) position per query
        _, idx = np.unique(rows, return_index=True)
        cols = cols[idx]
    assert cols.size == num_queries
    return cols2metrics(cols, num_queries)


def v2t_metrics(sims, query_masks=None):
    assert query_masks is None, "query_masks path is outside the hot-path scope"
    if _is_cuda_tensor(sims):
        assert sims.dim() == 2 and sims.shape[0] == sims.shape[1], "one caption per video expected"
        return cols2metrics(_device_ranks(sims)[1].cpu().numpy().astype(np.float64), sims.shape[0])
    sims = np.asarray(sims).T
    assert sims.ndim == 2, "expected a matrix"
    num_queries, num_caps = sims.shape
    assert num_queries == num_caps, "one caption per video expected"
    dists = -sims
    ranks = np.empty(num_queries)
    for ii in range(num_queries):
        row = dists[ii]
        ranks[ii] = np.where((np.sort(row) - row[ii]) == 0)[0].mean()     # tied ranks are averaged
    return cols2metrics(ranks, num_queries)
