"""N-rank check of the SHIPPED multi-GPU path (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
        scripts/ddp_trainer_check.py [--blocks 12] [--batch 4] [--steps 2]

The model goes through Multi_BaseTrainer_dist's DistributedDataParallel wrap (find_unused_parameters=True,
base/base_trainer.py:19-23) and the six hot lines of Multi_Trainer_dist._train_epoch (trainer_dist.py:158-163).
Checked: (1) the loss is the same number on every rank; (2) after backward p.grad is bit-identical on every rank;
(3) p.grad equals the mean over ranks of the local gradients obtained under no_sync() (DDP averages, the all-gather
backward only slices - SURVEY.md fact 7); (4) the step time through DDP. Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=12)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    from types import SimpleNamespace
    import bench
    from oa_transformer_b200.model import FrozenInTime, NormSoftmaxLoss
    from oa_transformer_b200.model.model import sim_matrix
    from oa_transformer_b200.synth import fill_seeded
    from oa_transformer_b200.trainer.trainer_dist import allgather_pair
    torch.manual_seed(0)
    m = FrozenInTime(
        video_params={"model": "SpaceTimeObjectTransformer", "arch_config": "base_patch16_224", "num_frames": bench.FRAMES,
                      "pretrained": True, "time_init": "rand", "allow_missing_vit": True, "depth": args.blocks},
        object_params={"model": "", "input_objects": True},
        text_params={"model": "distilbert-base-uncased", "pretrained": True, "random_init": True, "input": "text"})
    m.load_state_dict(fill_seeded({k: v for k, v in m.state_dict().items()}, 0, 0.02))
    m = m.to(device)
    m.train()
    # exactly the wrap of Multi_BaseTrainer_dist.__init__ (base/base_trainer.py:19-23)
    ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[local], find_unused_parameters=True)
    targs = SimpleNamespace(rank=rank, world_size=world, local_rank=local)
    loss_fn = NormSoftmaxLoss(0.05)
    host = bench.synth_batch(args.batch, rank, pinned=False)
    data = {"video": host["video"].to(device), "object": host["object"].to(device),
            "text": {k: v.to(device) for k, v in host["text"].items()}}
    params = [p for p in m.parameters() if p.requires_grad]

    def hot_lines():
        text_embeds, video_embeds = ddp(data, aug=True)
        video_embeds, text_embeds = allgather_pair(video_embeds, text_embeds, world, targs)
        loss = loss_fn(sim_matrix(text_embeds, video_embeds))
        loss.backward()
        return loss

    for p in params:
        p.grad = None
    torch.manual_seed(4321)              # the text tower draws its dropout seed from torch's CPU generator: same masks
    loss = hot_lines()                   # in the DDP evaluation and in the no_sync() one below
    torch.cuda.synchronize()
    losses = [torch.zeros(1, device=device) for _ in range(world)]
    dist.all_gather(losses, loss.detach().reshape(1))
    probe = [p for p in params if p.grad is not None]
    sums = torch.stack([p.grad.double().sum() for p in probe] + [p.grad.double().abs().sum() for p in probe])
    allsums = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(allsums, sums)
    reduced = [p.grad.detach().clone() for p in probe]
    for p in params:
        p.grad = None
    torch.manual_seed(4321)
    with ddp.no_sync():
        hot_lines()
    torch.cuda.synchronize()
    num2 = den2 = 0.0
    per = []
    for p, r in zip(probe, reduced):
        loc = p.grad.detach().clone()
        dist.all_reduce(loc, op=dist.ReduceOp.SUM)
        loc /= world
        per.append((float((loc - r).norm()), float(loc.norm())))
        num2 += per[-1][0] ** 2
        den2 += per[-1][1] ** 2
    gnorm = den2 ** 0.5
    # tensors that carry a visible share of the gradient (the key biases' gradient is mathematically zero: noise)
    worst = max(e / d for e, d in per if d > 1e-4 * gnorm)
    whole = (num2 / max(den2, 1e-300)) ** 0.5
    # step time through DDP
    for p in params:
        p.grad = None
    hot_lines()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for p in params:
            p.grad = None
        hot_lines()
    torch.cuda.synchronize()
    dist.barrier()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    if rank == 0:
        print(json.dumps({"check": "DDP path of Multi_BaseTrainer_dist + trainer_dist hot lines", "ranks": world,
                          "blocks": args.blocks, "batch_per_gpu": args.batch,
                          "loss": float(loss), "loss_identical_across_ranks": all(bool(torch.equal(losses[0], x)) for x in losses),
                          "p_grad_identical_across_ranks": all(bool(torch.equal(allsums[0], x)) for x in allsums),
                          "p_grad_tensors": len(probe), "params_without_grad": len(params) - len(probe),
                          "ddp_grad_vs_mean_of_local_grads_rel": whole,
                          "ddp_grad_vs_mean_of_local_grads_max_rel_per_tensor": worst, "ms_per_step_through_ddp": ms}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
