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
    ws = torch.zeros(ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F), device="cuda") if use else None
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
    ws = torch.zeros(ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F), device="cuda")
    t_old = timeit(lambda: ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse))
    t_tc = timeit(lambda: ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws))
    by, fl = ops.attn_core_work(ops.MODE_SPACE, B, T, H, F, n)
    print(json.dumps({"old_ms": t_old, "tc_ms": t_tc, "tc_TFLOPs": fl / t_tc / 1e9, "tc_GBps": by / t_tc / 1e6}))
