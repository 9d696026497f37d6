"""First-light check of the tcgen05 space-attention forward: error vs the mma.sync kernel + timing."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oa_transformer_b200 import ops
BF = torch.bfloat16
B, F, n, H = int(os.environ.get("ATT_B", 2)), int(os.environ.get("ATT_F", 2)), int(os.environ.get("ATT_N", 232)), int(os.environ.get("ATT_H", 3))
T = 1 + F * n
M = B * T
torch.manual_seed(0)
qkv = (torch.randn(M, 3 * H * 64, device="cuda") * 1.0)
qkv[:, :H * 64] *= 0.125
qkv = qkv.to(BF)
res = {}
for name, use in (("old", False), ("tc", True)):
    out = torch.zeros(M, H * 64, device="cuda", dtype=BF)
    lse = torch.zeros(B * H * T, device="cuda")
    ws = torch.zeros(ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F, n), device="cuda") if use else None
    ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws)
    torch.cuda.synchronize()
    res[name] = (out.float(), lse)
o0, l0 = res["old"]; o1, l1 = res["tc"]
d = (o0 - o1).abs()
print(json.dumps({"max_abs_out": float(d.max()), "ref_max": float(o0.abs().max()), "max_abs_lse": float((l0 - l1).abs().max()),
                  "nan": bool(torch.isnan(o1).any()), "cls_err": float(d.view(B, T, -1)[:, 0].max()),
                  "tile0_err": float(d.view(B, T, -1)[:, 1:129].max()), "tile1_err": float(d.view(B, T, -1)[:, 129:1 + n].max())}))
if os.environ.get("ATT_TIME"):
    def timeit(fn, iters=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    out = torch.zeros(M, H * 64, device="cuda", dtype=BF); lse = torch.zeros(B * H * T, device="cuda")
    ws = torch.zeros(ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F, n), device="cuda")
    t_old = timeit(lambda: ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse))
    t_tc = timeit(lambda: ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws))
    by, fl = ops.attn_core_work(ops.MODE_SPACE, B, T, H, F, n)
    print(json.dumps({"old_ms": t_old, "tc_ms": t_tc, "tc_TFLOPs": fl / t_tc / 1e9, "tc_GBps": by / t_tc / 1e6}))

# ---- backward: tcgen05 kernel vs the mma.sync kernel (OAT_SPACE_BWD_LEGACY=1)
out = torch.zeros(M, H * 64, device="cuda", dtype=BF); lse = torch.zeros(B * H * T, device="cuda")
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
rep = {"nan": bool(torch.isnan(gr["tc"]).any())}
for i, nm in enumerate("qkv"):
    rep["d%s_max_err" % nm] = float(d[:, 1:, i].max()); rep["d%s_ref_max" % nm] = float(gr["old"][:, 1:, i].abs().max())
    rep["d%s_cls_err" % nm] = float(d[:, 0, i].max()); rep["d%s_cls_ref" % nm] = float(gr["old"][:, 0, i].abs().max())
    rep["d%s_rel" % nm] = float((gr["old"][:, :, i] - gr["tc"][:, :, i]).norm() / gr["old"][:, :, i].norm())
print(json.dumps(rep))
if os.environ.get("ATT_TIME"):
    dqkv = torch.zeros_like(qkv)
    os.environ["OAT_SPACE_BWD_LEGACY"] = "1"
    t_old = timeit(lambda: ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc))
    os.environ.pop("OAT_SPACE_BWD_LEGACY", None)
    t_tc = timeit(lambda: ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc))
    print(json.dumps({"bwd_old_ms": t_old, "bwd_tc_ms": t_tc, "bwd_tc_TFLOPs_5mm": 2.5 * fl / t_tc / 1e9}))
