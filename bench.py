#!/usr/bin/env python
"""bench.py - video-text pairs/s, forward + backward, of the OA-Transformer dual-encoder hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch-per-gpu B]
  (N > 1: launched by torch.distributed.run, one rank per GPU, NCCL)

Workload (BASELINE.json metric / configs[3]): 8-frame 224x224 video + 36 object regions per frame + 32-token text,
per-GPU batch 32 (global 256 on 8 GPUs, weak scaling), ViT-B/16 space-time tower + DistilBERT-base, bf16 operands /
fp32 accumulation, synthetic inputs, seeded random weights. One step = text tower fwd, video tower fwd, embedding
all-gather, sim matrix + symmetric InfoNCE, backward of all of it, and (N > 1) the gradient all-reduce that the
reference's DDP wrapper performs (base/base_trainer.py:23). No optimizer step (as BASELINE.md defines the metric).

value  : device-timed (CUDA events, max over ranks) with inputs resident in HBM.
e2e    : the same step through the public plugin API (model.FrozenInTime -> AllGather -> sim_matrix ->
         NormSoftmaxLoss), inputs in pinned HOST memory, H2D copies and the loss read-back inside the timed region.
roofline / roofline_attention: per-launch CUDA-event timings of one extra instrumented step (never the timed steps).
cpu_baseline: the CPU oracle (a port of the reference algorithm, oracle/oracle.py) on the host cores, bounded sample.
--impl reference: the same CPU oracle as its own arm (the reference is pure PyTorch; /root/reference is absent on the
         GPU box, so the pinned port is what runs - `kind: "port"`).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "video-text pairs/sec (8-frame 224^2, 36 obj, 32-tok) fwd+bwd"
UNIT = "pairs/s"
FRAMES, OBJECTS, TEXT_LEN, IMG = 8, 36, 32, 224


def pair_flops(frames=FRAMES, objects=OBJECTS, text_len=TEXT_LEN, n_patches=196):
    """Algorithmic fwd FLOPs per pair (SURVEY.md section 8d formulas)."""
    n = n_patches + objects
    T = 1 + frames * n
    video = 12 * (T * 18874368 + 4 * 768 * (frames * n * (1 + n) + frames * n * (1 + frames) + 2 * T)) \
        + 2 * frames * n_patches * 768 * 768 + 2 * frames * objects * 2054 * 768 + 2 * 768 * 256
    text = 6 * (text_len * 14155776 + 4 * text_len * text_len * 768) + 393216
    return video + text


def load_gemm_traffic():
    """Measured DRAM bytes per GEMM launch (ncu --set full, profiles/r1_kernel_traffic.json made by
    scripts/ncu_traffic_table.py from one launch of every per-layer GEMM shape at the bench geometry), averaged over the
    18 GEMM launches of one transformer layer; beside it the algorithmic operand + output bytes of the same launches."""
    path = traffic_table_path()
    if path is None:
        return None, None
    with open(path) as f:
        t = json.load(f)
    count = {"gemm_fwd_qkv": 2, "gemm_fwd_proj": 2, "gemm_fwd_fc1": 1, "gemm_fwd_fc2": 1, "gemm_dgrad_fc2": 1,
             "gemm_dgrad_fc1": 1, "gemm_dgrad_qkv": 2, "gemm_dgrad_proj": 2, "gemm_wgrad_proj": 2, "gemm_wgrad_qkv": 2,
             "gemm_wgrad_fc1": 1, "gemm_wgrad_fc2": 1}
    if any(k not in t for k in count):
        return None, None
    n = sum(count.values())
    meas = sum(c * (t[k]["dram_read_bytes"] + t[k]["dram_write_bytes"]) for k, c in count.items()) / n
    return meas, "average over the 18 GEMM launches of one layer at M=59424 (B=32); per shape in profiles/" + os.path.basename(path)


def traffic_table_path():
    """Newest committed ncu traffic table (scripts/ncu_traffic_table.py output)."""
    for name in ("r2d_kernel_traffic.json", "r2c_kernel_traffic.json", "r2_kernel_traffic.json", "r1_kernel_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            return path
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [x.strip() for x in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def cpu_oracle_step_fn(batch, threads=None):
    """fwd+bwd of the pinned CPU oracle (fp32, what the reference computes) on `batch` synthetic pairs."""
    import torch
    from oracle import oracle as O
    from oracle.weights import dual_encoder_spec, fill_seeded
    if threads:
        torch.set_num_threads(threads)
    w = fill_seeded(dual_encoder_spec(frames=FRAMES, objects=True), 0, 0.02)
    p = {k: v.requires_grad_(True) for k, v in w.items()}
    g = torch.Generator().manual_seed(1234)
    video = torch.randn(batch, FRAMES, 3, IMG, IMG, generator=g)
    objects = O.synth_objects(batch, FRAMES, OBJECTS, g)
    text = O.synth_text(batch, TEXT_LEN, g)
    cfg = O.OracleCfg()

    def step():
        for v in p.values():
            v.grad = None
        te, ve = O.dual_encoder({"video": video, "object": objects, "text": text}, p, cfg)
        loss = O.norm_softmax_loss(O.sim_matrix(te, ve))
        loss.backward()
        return float(loss.detach())

    return step


class _StdoutToStderr:
    """Everything written to fd 1 while the run is in flight (NCCL's version banner, library chatter) goes to stderr, so
    that stdout carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def run_reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 2
    step = cpu_oracle_step_fn(batch, cores)
    warm = max(1, min(args.warmup, 2))
    for _ in range(warm):
        step()
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = batch / dt
    sample = "%d timed fwd+bwd steps of %d pairs each (8x224^2 frames, 36 objects, 32 tokens), fp32, %d threads; " \
             "steps capped at 10 and warm-up at 2 to bound the run" % (steps, batch, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "8-frame 224^2 + 36 obj/frame + 32-token text, fwd+bwd, CPU oracle port of the "
                                   "reference path, batch %d per step" % batch},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return line


# ------------------------------------------------------------------------------------------------ GPU arm
def build_model(device, seed=0):
    import torch
    from oa_transformer_b200.model import FrozenInTime
    from oa_transformer_b200.synth import fill_seeded
    torch.manual_seed(seed)
    m = FrozenInTime(
        video_params={"model": "SpaceTimeObjectTransformer", "arch_config": "base_patch16_224", "num_frames": FRAMES,
                      "pretrained": True, "time_init": "rand", "allow_missing_vit": True},
        object_params={"model": "", "input_objects": True},
        text_params={"model": "distilbert-base-uncased", "pretrained": True, "random_init": True, "input": "text"},
        projection_dim=256)
    sd = m.state_dict()
    new = fill_seeded({k: v for k, v in sd.items()}, seed, 0.02)
    m.load_state_dict(new)
    return m.to(device)


def synth_batch(batch, rank, pinned):
    import torch
    from oa_transformer_b200.synth import synth_objects, synth_text
    g = torch.Generator().manual_seed(1234 + rank)
    video = torch.randn(batch, FRAMES, 3, IMG, IMG, generator=g)
    objects = synth_objects(batch, FRAMES, OBJECTS, g)
    text = synth_text(batch, TEXT_LEN, g)
    if pinned:
        video, objects = video.pin_memory(), objects.pin_memory()
        text = {k: v.pin_memory() for k, v in text.items()}
    return {"video": video, "object": objects, "text": text}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        with _StdoutToStderr():
            line = run_reference_arm(args)
        if line is not None:
            print(json.dumps(line))
        return
    with _StdoutToStderr():
        line = run_ours(args)
    if line is not None:
        print(json.dumps(line))


def run_multi_gpu_check(model, params, step, dev, rank, world, device):
    import torch
    import torch.distributed as dist
    torch.manual_seed(4321)                              # same DistilBERT dropout masks in both evaluations below
    loss = step(dev)                                     # with the all-reduce
    torch.cuda.synchronize()
    losses = [torch.zeros(1, device=device) for _ in range(world)]
    dist.all_gather(losses, loss.detach().reshape(1).float())
    loss_identical = all(bool(torch.equal(losses[0], x)) for x in losses)
    probe = [p for p in params if p.grad is not None]
    books = [getattr(e, "_gradbook", None) for e in (model.video_model._engine, model._text_engine)]
    ptrs = {v.data_ptr() for b in books if b is not None for v in b.views.values()}
    aliased = all(p.grad.data_ptr() in ptrs for p in probe)
    # checksum of every reduced gradient, compared bit for bit across ranks
    sums = torch.stack([p.grad.double().sum() for p in probe] + [p.grad.double().abs().sum() for p in probe])
    allsums = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(allsums, sums)
    grads_identical = all(bool(torch.equal(allsums[0], x)) for x in allsums)
    reduced = [p.grad.detach().clone() for p in probe]
    # local (unreduced) gradients of the same step, averaged with a separate all-reduce on copies
    os.environ["OAT_BENCH_NO_REDUCE"] = "1"
    torch.manual_seed(4321)
    try:
        step(dev)
    finally:
        os.environ.pop("OAT_BENCH_NO_REDUCE")
    torch.cuda.synchronize()
    # per tensor, relative to the tensor's share of the whole gradient (mathematically-zero gradients such as the key
    # biases - softmax is shift-invariant - are rounding noise in both evaluations), and over the whole gradient
    num2 = den2 = 0.0
    per = []
    for p, r in zip(probe, reduced):
        local = p.grad.detach().clone()
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        local /= world
        per.append((float((local - r).norm()), float(local.norm())))
        num2 += per[-1][0] ** 2
        den2 += per[-1][1] ** 2
    gnorm = den2 ** 0.5
    worst = max(e / d for e, d in per if d > 1e-4 * gnorm)
    return {"ranks": world, "loss_identical_across_ranks": loss_identical,
            "reduced_grads_identical_across_ranks": grads_identical, "p_grad_tensors_checked": len(probe),
            "reduced_vs_mean_of_local_grads_rel": (num2 / max(den2, 1e-300)) ** 0.5,
            "reduced_vs_mean_of_local_grads_max_rel_per_tensor": worst,
            "p_grad_aliases_the_reduced_book": bool(aliased)}


def run_ours(args):

    import torch
    import torch.distributed as dist
    from oa_transformer_b200 import ops
    from oa_transformer_b200._lib import lib, check
    from oa_transformer_b200.functional import AllGatherPairSlice
    from oa_transformer_b200.model import NormSoftmaxLoss, sim_matrix

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    check(lib().oat_device_check(), "oat_device_check")
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    W = max(args.warmup, 3)
    K = args.steps
    B = args.batch_per_gpu

    model = build_model(device)
    model.train()
    loss_fn = NormSoftmaxLoss(0.05)
    params = [p for p in model.parameters() if p.requires_grad]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # what DDP's reducer does (base/base_trainer.py:23): average parameter gradients across ranks. Each tower's flat
    # gradient book is all-reduced as soon as that tower's backward has been enqueued (engine.grad_ready_hook): the
    # text tower finishes first, so its all-reduce runs under the video tower's backward. The towers hand autograd
    # fresh views of the book, so p.grad ALIASES the book (no clone) and the in-place all-reduce is the all-reduce of
    # p.grad; step() waits for the handles before it returns (multi_gpu_check verifies p.grad across ranks).
    pending = []
    # OAT_LAYER_REDUCE=1: all-reduce block by block during the video backward (launched from the side stream) instead of
    # one all-reduce per tower. With the GEMM's dynamic tile scheduler the chain keeps its speed next to the NCCL CTAs
    # (the static tile stride lost 2 ms: 62.2 vs 59.8-60.5 ms at N = 2), but there is little left to hide: measured
    # 58.5-58.8 vs 58.8 ms (N = 2) and 60.9 vs 60.7 ms (N = 4) against 58.0 ms without any all-reduce - within the
    # box-to-box noise, so the simpler per-tower reduce stays the default.
    layer_reduce = os.environ.get("OAT_LAYER_REDUCE", "0") != "0"
    reduced = set()

    def book_range(book, prefix):
        """[lo, hi) of the flat book covered by the parameters whose name starts with `prefix` (contiguous by construction)."""
        spans = [(v.data_ptr(), v.numel()) for n, v in book.views.items() if n.startswith(prefix)]
        base = book.flat.data_ptr()
        lo = min(p for p, _ in spans)
        hi = max(p + 4 * n for p, n in spans)
        assert sum(n for _, n in spans) * 4 == hi - lo, "parameters of %s are not contiguous in the gradient book" % prefix
        return (lo - base) // 4, (hi - base) // 4

    def layer_hook(book, prefix):
        if world > 1 and layer_reduce and not os.environ.get("OAT_BENCH_NO_REDUCE"):
            lo, hi = book_range(book, prefix)
            reduced.add((lo, hi))
            pending.append(dist.all_reduce(book.flat[lo:hi], op=dist.ReduceOp.AVG, async_op=True))

    def start_reduce(book):
        if world > 1 and not os.environ.get("OAT_BENCH_NO_REDUCE"):
            done = sorted(r for r in reduced if True)
            if not (layer_reduce and book is getattr(model.video_model._engine, "_gradbook", None)):
                pending.append(dist.all_reduce(book.flat, op=dist.ReduceOp.AVG, async_op=True))
            else:                                   # the blocks went out one by one: reduce what lies around them
                cur = 0
                for lo, hi in done + [(book.flat.numel(), book.flat.numel())]:
                    if lo > cur:
                        pending.append(dist.all_reduce(book.flat[cur:lo], op=dist.ReduceOp.AVG, async_op=True))
                    cur = max(cur, hi)
            reduced.clear()

    def reduce_grads():
        while pending:
            pending.pop().wait()

    def step(data):
        for p in params:
            p.grad = None
        text_e, video_e = model(data, aug=True)
        for eng in (model.video_model._engine, model._text_engine):
            eng.grad_ready_hook = start_reduce
        model.video_model._engine.layer_grad_hook = layer_hook
        video_g, text_g = AllGatherPairSlice.apply(video_e, text_e, rank, world)      # ONE packed all-gather
        loss = loss_fn(sim_matrix(text_g, video_g))
        loss.backward()
        reduce_grads()
        return loss

    host = synth_batch(B, rank, pinned=True)
    dev = {"video": host["video"].to(device), "object": host["object"].to(device),
           "text": {k: v.to(device) for k, v in host["text"].items()}}

    # ---------------- multi-GPU self-check (N > 1): same loss on every rank, identical averaged p.grad on every rank, and
    # p.grad == average over ranks of the LOCAL gradients (what DDP delivers, base/base_trainer.py:23) where each local
    # gradient comes from the all-gather backward's unreduced slice (trainer_dist.py:40-45)
    multi_gpu_check = None
    if world > 1:
        multi_gpu_check = run_multi_gpu_check(model, params, step, dev, rank, world, device)

    # ---------------- device-resident timing
    for _ in range(W):
        step(dev)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        loss = step(dev)
    e1.record()
    barrier()
    launches = (ops.LAUNCHES - launches0) // K
    # diagnostic: host time to enqueue ONE step into an empty stream (no queue back-pressure), not part of any metric
    t_host = time.perf_counter()
    step(dev)
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3
    barrier()
    ms = e0.elapsed_time(e1) / K
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(loss.detach())
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = world * B / (ms / 1e3)

    # ---------------- end-to-end through the plugin API with host inputs
    e2e = None
    if not args.no_e2e:
        h2d = host["video"].numel() * 4 + host["object"].numel() * 4 + sum(v.numel() * 8 for v in host["text"].values())

        # the plugin's own input path: pinned host batches staged one step ahead on a copy stream
        # (data_loader.DevicePrefetcher); every step still moves one full batch host -> device inside the timed region
        from oa_transformer_b200.data_loader import DevicePrefetcher

        def host_batches(n):
            for _ in range(n):
                yield host

        host_ms = {"stage_next_batch": 0.0, "enqueue_step": 0.0, "wait_for_loss": 0.0}
        step_walls = []
        pool = {}             # the prefetcher's two device buffer sets, kept across the warm-up and the timed loop

        def run_e2e(n):
            it = iter(DevicePrefetcher(host_batches(n), device, pool=pool))
            while True:
                t0 = time.perf_counter()
                try:
                    data = next(it)                       # waits for nothing: enqueues the H2D copies of the NEXT batch
                except StopIteration:
                    break
                t1 = time.perf_counter()
                loss = step(data)
                t2 = time.perf_counter()
                float(loss.detach().item())               # D2H read of the loss closes the step
                t3 = time.perf_counter()
                host_ms["stage_next_batch"] += (t1 - t0) * 1e3
                host_ms["enqueue_step"] += (t2 - t1) * 1e3
                host_ms["wait_for_loss"] += (t3 - t2) * 1e3
                step_walls.append((t3 - t0) * 1e3)

        run_e2e(2)
        barrier()
        # diagnostic: the bare H2D copy of one batch (PCIe + host memory of this box), not part of any metric
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            host["video"].to(device, non_blocking=True)
            host["object"].to(device, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_ms = c0.elapsed_time(c1) / 3
        barrier()
        for k in host_ms:
            host_ms[k] = 0.0
        del step_walls[:]
        # every step ends with a host sync (the loss read-back), so a collector pause lands in the step time one to
        # one: collect now and keep the cyclic GC off inside the timed loop (what timeit does)
        gc.collect()
        gc.disable()
        try:
            t0 = time.perf_counter()
            e0.record()
            run_e2e(K)
            e1.record()
            barrier()
            wall = (time.perf_counter() - t0) / K * 1e3
        finally:
            gc.enable()
        ems = max(e0.elapsed_time(e1) / K, wall)
        t = torch.tensor([ems], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B / (float(t) / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": 4, "ms_per_step": float(t), "bare_h2d_copy_ms": h2d_ms,
               "host_ms_per_step": {k: round(v / K, 2) for k, v in host_ms.items()},
               "step_wall_ms_min_median_max": [round(sorted(step_walls)[0], 2), round(sorted(step_walls)[len(step_walls) // 2], 2),
                                               round(sorted(step_walls)[-1], 2)] if step_walls else None,
               "h2d": "pinned host batch -> device on a copy stream, one step ahead (double-buffered), every step"}

    # ---------------- per-launch roofline numbers from one extra instrumented step (rank 0 only)
    roofline = roofline_attn = None
    peaks = load_peaks()
    breakdown = None
    if not args.no_profile:
        # every rank runs the instrumented step (it contains collectives); only rank 0 keeps the numbers
        from oa_transformer_b200 import engine as _engine
        side_was, _engine.SIDE_STREAM = _engine.SIDE_STREAM, False   # per-launch durations need serial execution
        ops.PROFILE = []
        step(dev)
        barrier()
        prof, ops.PROFILE = ops.PROFILE, None
        _engine.SIDE_STREAM = side_was
    if rank == 0 and not args.no_profile:
        agg = {}
        for kind, work, a, b in prof:
            d = agg.setdefault(kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            d["ms"] += a.elapsed_time(b)
            d["n"] += 1
            if isinstance(work, tuple):
                d["bytes"] += work[0]
                d["flops"] += work[1]
            else:
                d["flops"] += work
        gm = agg.get("gemm")
        if gm and gm["ms"] > 0:
            ach = gm["flops"] / (gm["ms"] / 1e3) / 1e12
            traffic, traffic_note = load_gemm_traffic() if B == 32 else (None, None)
            roofline = {"kernel": "gemm_bf16_kernel (tcgen05)", "bound": "tensor", "achieved": ach,
                        "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["bf16_tflops_sustained"], "traffic": traffic,
                        "traffic_note": traffic_note, "launches": gm["n"],
                        "ms_in_step": gm["ms"], "peak_source": peaks["source"] + ", sustained bf16",
                        "note": "algorithmic 2*M*N*K summed over every GEMM launch of one step / summed CUDA-event "
                                "durations of those launches (instrumented extra step)"}
        # attention cores: HBM-bound. Algorithmic bytes per group = 128 B x (q + out rows, k + v rows) forward and
        # twice that backward (q, dO, O, dQ rows; k, v, dK, dV rows) - ops.attn_core_work / SURVEY.md 8(d).
        ktraffic = None
        tpath = traffic_table_path()
        if B == 32 and tpath is not None:
            with open(tpath) as f:
                ktraffic = json.load(f)
        roofline_attn = {}
        # (the profiler records the backward's own algorithmic bytes: without the O rows when delta comes from the GEMM)
        ext = "_delta" if (ktraffic and "attn_space_bwd_delta" in ktraffic and _engine.DELTA_EPI) else ""
        for key, label, fmult, tkey in (("attn_fwd_0", "space_fwd (tcgen05, CLS fused)", 1.0, "attn_space_fwd"),
                                        ("attn_bwd_0", "space_bwd (tcgen05)", 2.5, "attn_space_bwd" + ext),
                                        ("attn_fwd_1", "time_fwd (mma.sync, CLS fused)", 1.0, "attn_time_fwd"),
                                        ("attn_bwd_1", "time_bwd (mma.sync)", 2.5, "attn_time_bwd" + ext)):
            sp = agg.get(key)
            if not sp or sp["ms"] <= 0:
                continue
            gbs = sp["bytes"] / (sp["ms"] / 1e3) / 1e9
            tr = None
            if ktraffic and tkey in ktraffic:
                tr = ktraffic[tkey]["dram_read_bytes"] + ktraffic[tkey]["dram_write_bytes"]
            roofline_attn[label] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes": sp["bytes"] / sp["n"],
                                    "traffic": tr, "launches": sp["n"], "ms_in_step": sp["ms"],
                                    "tflops": fmult * sp["flops"] / (sp["ms"] / 1e3) / 1e12}
        breakdown = {k: {"ms": round(v["ms"], 3), "n": v["n"]} for k, v in agg.items()}

    # ---------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            cb = 4                                  # B = 4 at F = 8, 1 warm-up + 3 timed, median
            cstep = cpu_oracle_step_fn(cb, cores)
            cstep()
            times = []
            for _ in range(3):
                t0 = time.perf_counter()
                cstep()
                times.append(time.perf_counter() - t0)
            cdt = sorted(times)[1]
            cpu = {"value": cb / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "median of 3 fwd+bwd steps of %d pairs (same per-pair workload: 8x224^2 frames, 36 objects, "
                             "32 tokens), fp32 oracle port of the reference path, after 1 warm-up" % cb}
        except Exception as ex:  # the baseline is informative; never lose the GPU line over it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        flops = 3.0 * pair_flops()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "WebVid+CC3M-synth (BASELINE configs[3] per-GPU shard): 8-frame 224^2, 36 "
                                       "obj/frame, 32-token text, batch %d per GPU, fwd+bwd (+grad all-reduce when "
                                       "N>1), ViT-B/16 space-time + DistilBERT-base" % B,
                           "global_batch": world * B, "tokens_per_video": 1 + FRAMES * (196 + OBJECTS),
                           "parallelism": "dp%d" % world,
                           "l2": "per-step working set (activations ~%.0f GB) far exceeds the 126 MB L2; no flush "
                                 "needed" % (B * 0.85)},
                "loss": loss_val,
                "model_tflops": value * flops / 1e12,
                "mfu_vs_sustained_bf16": value * flops / 1e12 / world / peaks["bf16_tflops_sustained"],
                "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks, "e2e": e2e,
                "multi_gpu_check": multi_gpu_check, "roofline": roofline,
                "roofline_attention": roofline_attn, "kernel_ms_breakdown": breakdown, "cpu_baseline": cpu}
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


if __name__ == "__main__":
    main()
