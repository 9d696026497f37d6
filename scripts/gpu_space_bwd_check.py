"""Space-attention backward on tcgen05: error vs the mma.sync kernel (OAT_SPACE_BWD_LEGACY=1) + timing.
OAT_SPACE_BWD_V1=1 selects the unpipelined tcgen05 kernel (read once per process: run twice for an A/B)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oa_transformer_b200 import ops
BF = torch.bfloat16


def run(B, F, n, H, time_it):
    T = 1 + F * n
    M = B * T
    torch.manual_seed(0)
    qkv = torch.randn(M, 3 * H * 64, device="cuda")
    qkv[:, :H * 64] *= 0.125
    qkv = qkv.to(BF)
    out = torch.zeros(M, H * 64, device="cuda", dtype=BF)
    lse = torch.zeros(B * H * T, device="cuda")
    ws = torch.zeros(ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F, n), device="cuda")
    ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws)
    dout = torch.randn(M, H * 64, device="cuda").to(BF)
    acc = torch.empty(B * H * 192, device="cuda")
    gr = {}
    for name in ("old", "tc"):
        if name == "old":
            os.environ["OAT_SPACE_BWD_LEGACY"] = "1"
        else:
            os.environ.pop("OAT_SPACE_BWD_LEGACY", None)
        dqkv = torch.zeros_like(qkv)
        ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc)
        torch.cuda.synchronize()
        gr[name] = dqkv.float().view(B, T, 3, H * 64)
    d = (gr["old"] - gr["tc"]).abs()
    rep = {"B": B, "F": F, "n": n, "H": H, "nan": bool(torch.isnan(gr["tc"]).any())}
    for i, nm in enumerate("qkv"):
        rep["d%s_max_err" % nm] = float(d[:, 1:, i].max())
        rep["d%s_cls_err" % nm] = float(d[:, 0, i].max())
        rep["d%s_rel" % nm] = float((gr["old"][:, :, i] - gr["tc"][:, :, i]).norm() / gr["old"][:, :, i].norm())
    if rep["dq_rel"] > 1e-2:
        dq = d[:, 1:, 0].view(B, F, n, H, 64)
        ref = gr["old"][:, 1:, 0].view(B, F, n, H, 64)
        rows = dq.amax(dim=(0, 1, 3, 4))
        bad = (rows > 0.05).nonzero().flatten().tolist()
        rep["bad_rows"] = "%d rows bad, first %s last %s" % (len(bad), bad[:6], bad[-6:])
        rep["bad_by_bfh"] = dq.amax(dim=(2, 4)).flatten().tolist()[:24]
        r0 = bad[0]
        rep["sample"] = [[round(float(x), 3) for x in gr["tc"][:, 1:, 0].view(B, F, n, H, 64)[0, 0, r0, 0, :6]],
                         [round(float(x), 3) for x in ref[0, 0, r0, 0, :6]]]
    if time_it:
        def timeit(fn, iters=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        dqkv = torch.zeros_like(qkv)
        rep["bwd_ms"] = timeit(lambda: ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc))
        by, fl = ops.attn_core_work(ops.MODE_SPACE, B, T, H, F, n)
        rep["bwd_GBps"] = 2 * by / rep["bwd_ms"] / 1e6
    print(json.dumps(rep), flush=True)


if __name__ == "__main__":
    print("kernel:", "v1" if os.environ.get("OAT_SPACE_BWD_V1") else "v2 (pipelined)")
    run(2, 2, 232, 3, False)
    run(1, 1, 196, 2, False)
    run(2, 3, 128, 2, False)
    run(3, 2, 250, 1, False)
    run(32, 8, 232, 12, True)
