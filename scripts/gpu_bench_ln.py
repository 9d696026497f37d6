"""Standalone timing of the LayerNorm kernels at the benchmark geometry (M = 59424 rows, D = 768)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oa_transformer_b200 import ops
M, D = 59424, 768
BF = torch.bfloat16
x = torch.randn(M, D, device="cuda"); gamma = torch.ones(D, device="cuda"); beta = torch.zeros(D, device="cuda")
mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda"); y16 = torch.empty(M, D, device="cuda", dtype=BF)
dy16 = torch.randn(M, D, device="cuda").to(BF); a1 = torch.randn(M, D, device="cuda"); a2 = torch.randn(M, D, device="cuda")
dx = torch.empty(M, D, device="cuda"); dx16 = torch.empty(M, D, device="cuda", dtype=BF)
dg, db, ds = (torch.zeros(D, device="cuda") for _ in range(3))
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
f = timeit(lambda: ops.layernorm_fwd(x, gamma, beta, 1e-6, y_bf16=y16, mean=mean, rstd=rstd))
b1 = timeit(lambda: ops.layernorm_bwd(x, mean, rstd, gamma, dy_bf16=dy16, add1=a1, dx=dx, dx_bf16=dx16, dgamma=dg, dbeta=db, dxsum=ds))
b2 = timeit(lambda: ops.layernorm_bwd(x, mean, rstd, gamma, dy_bf16=dy16, add1=a1, add2=a2, dx=dx, dx_bf16=dx16, dgamma=dg, dbeta=db, dxsum=ds))
b0 = timeit(lambda: ops.layernorm_bwd(x, mean, rstd, gamma, dy_bf16=dy16, dx=dx, dx_bf16=dx16, dgamma=dg, dbeta=db, dxsum=ds))
el = M * D
print(json.dumps({"fwd_ms": round(f, 4), "fwd_GBps": round(el * 6 / f / 1e6, 1),
                  "bwd_plain_ms": round(b0, 4), "bwd_plain_GBps": round(el * 12 / b0 / 1e6, 1),
                  "bwd_add1_ms": round(b1, 4), "bwd_add1_GBps": round(el * 16 / b1 / 1e6, 1),
                  "bwd_add2_ms": round(b2, 4), "bwd_add2_GBps": round(el * 20 / b2 / 1e6, 1)}))
x3 = torch.randn(M, 3 * D, device="cuda").to(BF); x4 = torch.randn(M, 4 * D, device="cuda").to(BF)
o3 = torch.zeros(3 * D, device="cuda"); o4 = torch.zeros(4 * D, device="cuda")
c3 = timeit(lambda: ops.colsum_bf16(x3, o3)); c4 = timeit(lambda: ops.colsum_bf16(x4, o4))
print(json.dumps({"colsum_2304_ms": round(c3, 4), "colsum_2304_GBps": round(M * 3 * D * 2 / c3 / 1e6, 1),
                  "colsum_3072_ms": round(c4, 4), "colsum_3072_GBps": round(M * 4 * D * 2 / c4 / 1e6, 1)}))
