"""world_size-2 gloo tests (CPU) of the N>1 host logic: the embedding all-gather with slice-backward semantics
(OATrans/trainer/trainer_dist.py:29-45) against the reference-generated fixture tests/golden/allgather2.pt."""
import os

import torch
import torch.multiprocessing as mp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _worker(rank, world, port, ret, packed=False):
    import torch.distributed as dist
    from oa_transformer_b200.functional import AllGatherSlice
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.load(os.path.join(GOLD, "allgather2.pt"), map_location="cpu", weights_only=False)
    B = g["B"]
    t = g["text"][rank * B:(rank + 1) * B].clone().requires_grad_(True)
    v = g["video"][rank * B:(rank + 1) * B].clone().requires_grad_(True)
    if packed:      # one packed collective for both tensors (functional.AllGatherPairSlice)
        from oa_transformer_b200.functional import AllGatherPairSlice
        vg, tg = AllGatherPairSlice.apply(v, t, rank, world)
    else:
        vg = AllGatherSlice.apply(v, rank, world)
        tg = AllGatherSlice.apply(t, rank, world)
    assert torch.equal(vg.detach(), g["video"]) and torch.equal(tg.detach(), g["text"])   # rank-order concatenation
    loss = O.norm_softmax_loss(O.sim_matrix(tg, vg))
    loss.backward()
    ret[rank] = (float(loss), t.grad.clone(), v.grad.clone())
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("packed", [False, True])
def test_allgather_slice_two_ranks_matches_reference(packed):
    g = torch.load(os.path.join(GOLD, "allgather2.pt"), map_location="cpu", weights_only=False)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29643 + int(packed), ret, packed), nprocs=2, join=True)
    for r in range(2):
        loss, tg, vg = ret[r]
        ref = g["ranks"][r]
        assert abs(loss - float(ref["loss"])) < 1e-5
        # local slice of the global gradient, NOT reduced over ranks
        assert torch.allclose(tg, ref["t_grad"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(vg, ref["v_grad"], rtol=1e-5, atol=1e-7)


def test_allgather_single_process_is_identity_with_full_gradient():
    from oa_transformer_b200.functional import AllGatherSlice
    x = torch.randn(4, 8, requires_grad=True)
    y = AllGatherSlice.apply(x, 0, 1)
    assert torch.equal(y, x)
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
