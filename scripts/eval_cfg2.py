"""BASELINE configs[1] - "MSR-VTT 1k-A eval synth": 1000 synthetic video-text pairs (8-frame 224^2, 36 objects / frame,
32-token text) through the forward path only, as trainer_dist._valid_epoch does (trainer/trainer_dist.py:201-281): embed in
batches, gather, one 1000 x 1000 sim_matrix, t2v / v2t retrieval metrics. Reports forward latency per batch and the
metrics (random-init weights: recall is at chance level; the point is the path and the rank bookkeeping, which is checked
bit-exactly against the host numpy port of model/metric.py on the same matrix).

  python scripts/eval_cfg2.py [--pairs 1000] [--batch 50]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from oa_transformer_b200.model import metric as M, sim_matrix  # noqa: E402
from oa_transformer_b200.synth import synth_objects, synth_text  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=50)
    ap.add_argument("--oracle-pairs", type=int, default=24,
                    help="first N pairs also go through the fp32 CPU oracle (= what the reference computes): logits, "
                         "R@1 / R@5 and forward latency of that sub-problem, reference vs CUDA")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev).eval()
    nb = args.pairs // args.batch
    text_e, video_e, lat = [], [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for i in range(-2, nb):                                  # two warm-up batches
            g = torch.Generator().manual_seed(777 + max(i, 0))
            data = {"video": torch.randn(args.batch, bench.FRAMES, 3, bench.IMG, bench.IMG, generator=g).to(dev),
                    "object": synth_objects(args.batch, bench.FRAMES, bench.OBJECTS, g).to(dev),
                    "text": {k: v.to(dev) for k, v in synth_text(args.batch, bench.TEXT_LEN, g).items()}}
            torch.cuda.synchronize()
            e0.record()
            t, v = model(data, return_embeds=True)
            e1.record()
            torch.cuda.synchronize()
            if i >= 0:
                lat.append(e0.elapsed_time(e1))
                text_e.append(t.clone())
                video_e.append(v.clone())
        sims = sim_matrix(torch.cat(text_e), torch.cat(video_e))
        t2v, v2t = M.t2v_metrics(sims), M.v2t_metrics(sims)
        host = sims.cpu().numpy()
        same = M.t2v_metrics(host) == t2v and M.v2t_metrics(host) == v2t
    vs_ref = None
    if args.oracle_pairs > 0:
        import time
        from oracle import oracle as O
        n = min(args.oracle_pairs, args.batch)
        g = torch.Generator().manual_seed(777)
        video = torch.randn(args.batch, bench.FRAMES, 3, bench.IMG, bench.IMG, generator=g)[:n]
        objects = synth_objects(args.batch, bench.FRAMES, bench.OBJECTS, g)[:n]
        text = {k: v[:n] for k, v in synth_text(args.batch, bench.TEXT_LEN, g).items()}
        w = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            t0 = time.perf_counter()
            te, ve = O.dual_encoder({"video": video, "object": objects, "text": text}, w, O.OracleCfg())
            cpu_s = time.perf_counter() - t0
        ref = O.sim_matrix(te, ve).numpy()
        ours = host[:n, :n]
        vs_ref = {"pairs": n, "logit_max_abs_err": float(abs(ours - ref).max()),
                  "reference_t2v": M.t2v_metrics(ref), "ours_t2v": M.t2v_metrics(ours.copy()),
                  "reference_v2t": M.v2t_metrics(ref), "ours_v2t": M.v2t_metrics(ours.copy()),
                  "reference_cpu_fwd_pairs_per_s": n / cpu_s, "cpu_threads": os.cpu_count()}
    lat.sort()
    print(json.dumps({"vs_reference_fp32_oracle": vs_ref,"config": "cfg2: %d pairs, batch %d, 8x224^2 + 36 obj + 32 tok, forward only" % (args.pairs, args.batch),
                      "fwd_ms_per_batch_median": lat[len(lat) // 2], "fwd_pairs_per_s": args.batch / (lat[len(lat) // 2] / 1e3),
                      "t2v": t2v, "v2t": v2t, "device_metrics_equal_host_numpy": bool(same),
                      "note": "random-init weights: recall at chance level (0.1 % R@1 expected)"}))


if __name__ == "__main__":
    main()
