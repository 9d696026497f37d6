"""Isolate the cost of MN-major operands / split-K in the weight-gradient GEMM shape [2304 x 768, K = 59424]."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oa_transformer_b200 import ops
BF = torch.bfloat16
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
M, N, K = 2304, 768, 59424
a_k = torch.randn(M, K, device="cuda").to(BF); a_mn = torch.randn(K, M, device="cuda").to(BF)
b_k = torch.randn(N, K, device="cuda").to(BF); b_mn = torch.randn(K, N, device="cuda").to(BF)
out = torch.zeros(M, N, device="cuda")
for am in (0, 1):
    for bm in (0, 1):
        for sk in (0, 4, 8, 16):
            A = a_mn if am else a_k; B = b_mn if bm else b_k
            ms = timeit(lambda: ops.gemm(A, B, a_major=am, b_major=bm, out_f32=out, accumulate=True, split_k=sk))
            print(json.dumps({"a_major": am, "b_major": bm, "split_k": sk, "ms": round(ms, 4), "tflops": round(2.0*M*N*K/ms/1e9, 1)}))
# swapped roles: compute dW^T [768 x 2304] instead (A = X mn-major with M=768, B = dY mn-major with N=2304)
out2 = torch.zeros(N, M, device="cuda")
for sk in (0, 8, 16):
    ms = timeit(lambda: ops.gemm(b_mn, a_mn, a_major=1, b_major=1, out_f32=out2, accumulate=True, split_k=sk))
    print(json.dumps({"swapped": True, "split_k": sk, "ms": round(ms, 4), "tflops": round(2.0*M*N*K/ms/1e9, 1)}))
