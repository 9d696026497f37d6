"""Standalone timing of the attention kernels at the benchmark geometry (B=32, F=8, n=232, H=12)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oa_transformer_b200 import ops
BF = torch.bfloat16
B, F, n, H = int(os.environ.get("ATT_B", 32)), 8, 232, 12
T = 1 + F * n
M = B * T
qkv = (torch.randn(M, 3 * H * 64, device="cuda") * 0.5).to(BF)
out = torch.empty(M, H * 64, device="cuda", dtype=BF)
lse = torch.empty(B * H * T, device="cuda")
dout = torch.randn(M, H * 64, device="cuda").to(BF)
dqkv = torch.empty_like(qkv)
acc = torch.empty(B * H * 192, device="cuda")
ws = torch.zeros(max(ops.attn_fwd_workspace_floats(m, B, H, F, n) for m in (ops.MODE_SPACE, ops.MODE_TIME)), device="cuda")
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, mode in (("space", ops.MODE_SPACE), ("time", ops.MODE_TIME)):
    f = timeit(lambda: ops.attn_fwd(mode, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws))
    b = timeit(lambda: ops.attn_bwd(mode, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc))
    # with delta = rowsum(dO * O) handed in (engine: the projection dgrad's act-4 epilogue writes it)
    delta = (dout.float() * out.float()).view(M, H, 64).sum(-1).t().contiguous()
    bd = timeit(lambda: ops.attn_bwd(mode, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc, delta=delta))
    by, fl = ops.attn_core_work(mode, B, T, H, F, n)
    print(json.dumps({"mode": name, "fwd_ms": round(f, 4), "bwd_ms": round(b, 4), "bwd_ext_delta_ms": round(bd, 4), "fwd_GBps": round(by / f / 1e6, 1),
                      "fwd_TFLOPs": round(fl / f / 1e9, 1), "bwd_TFLOPs_5mm": round(2.5 * fl / b / 1e9, 1)}))
