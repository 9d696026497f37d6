"""Per-shape timing of the tcgen05 GEMM on the shapes of one training step (B=32, T=1857 -> M=59424 token rows)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oa_transformer_b200 import ops  # noqa: E402

BF = torch.bfloat16


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


def main():
    M = int(os.environ.get("GEMM_M", 59424))
    only = os.environ.get("GEMM_ONLY")
    D = 768
    dev = "cuda"
    x = torch.randn(M, D, device=dev).to(BF)
    x4 = torch.randn(M, 4 * D, device=dev).to(BF)
    x3 = torch.randn(M, 3 * D, device=dev).to(BF)
    res = torch.randn(M, D, device=dev)
    o32 = torch.empty(M, D, device=dev)
    o16_3 = torch.empty(M, 3 * D, device=dev, dtype=BF)
    o16_4 = torch.empty(M, 4 * D, device=dev, dtype=BF)
    o16_4b = torch.empty(M, 4 * D, device=dev, dtype=BF)
    o16_1 = torch.empty(M, D, device=dev, dtype=BF)
    wqkv = torch.randn(3 * D, D, device=dev).to(BF)
    wproj = torch.randn(D, D, device=dev).to(BF)
    w1 = torch.randn(4 * D, D, device=dev).to(BF)
    w2 = torch.randn(D, 4 * D, device=dev).to(BF)
    bias3 = torch.randn(3 * D, device=dev)
    bias1 = torch.randn(D, device=dev)
    bias4 = torch.randn(4 * D, device=dev)
    aux1 = torch.randn(M, D, device=dev).to(BF)
    rowdot = torch.empty(D // 64, M, device=dev)
    dw_qkv = torch.zeros(3 * D, D, device=dev)
    dw_1 = torch.zeros(4 * D, D, device=dev)
    dw_2 = torch.zeros(D, 4 * D, device=dev)
    dw_p = torch.zeros(D, D, device=dev)
    cases = {
        "fwd_qkv   [M,2304,768]  bf16 out, bias+qscale": (lambda: ops.gemm(x, wqkv, bias=bias3, scale_cols=D, scale=0.125, out_bf16=o16_3), 2.0 * M * 3 * D * D),
        "fwd_qkv_plain [M,2304,768] bf16 out": (lambda: ops.gemm(x, wqkv, out_bf16=o16_3), 2.0 * M * 3 * D * D),
        "fwd_proj  [M,768,768]   f32 out + residual": (lambda: ops.gemm(x, wproj, bias=bias1, residual=res, out_f32=o32), 2.0 * M * D * D),
        "fwd_fc1   [M,3072,768]  gelu, 2x bf16 out": (lambda: ops.gemm(x, w1, bias=bias4, act=ops.ACT_GELU, out_bf16=o16_4, out2_bf16=o16_4b), 2.0 * M * 4 * D * D),
        "fwd_fc2   [M,768,3072]  f32 out + residual": (lambda: ops.gemm(x4, w2, bias=bias1, residual=res, out_f32=o32), 2.0 * M * 4 * D * D),
        "dgrad_fc2 [M,3072,768]  B mn-major, gelu' aux": (lambda: ops.gemm(x, w2, b_major=1, act=ops.ACT_GELU_BWD, aux=o16_4b, out_bf16=o16_4), 2.0 * M * 4 * D * D),
        "dgrad_fc1 [M,768,3072]  B mn-major bf16 out": (lambda: ops.gemm(x4, w1, b_major=1, out_bf16=o16_1), 2.0 * M * 4 * D * D),
        "dgrad_qkv [M,768,2304]  B mn-major bf16 out": (lambda: ops.gemm(x3, wqkv, b_major=1, out_bf16=o16_1), 2.0 * M * 3 * D * D),
        "dgrad_proj[M,768,768]   B mn-major bf16 out": (lambda: ops.gemm(x, wproj, b_major=1, out_bf16=o16_1), 2.0 * M * D * D),
        "dgrad_proj_rowdot [M,768,768] + delta = rowsum(dO*O)": (lambda: ops.gemm(x, wproj, b_major=1, act=ops.ACT_ROWDOT, aux=aux1, rowdot=rowdot, out_bf16=o16_1), 2.0 * M * D * D),
        "wgrad_qkv [2304,768,M]  both mn-major splitK": (lambda: ops.gemm(x3, x, a_major=1, b_major=1, out_f32=dw_qkv, accumulate=True), 2.0 * M * 3 * D * D),
        "wgrad_fc1 [3072,768,M]": (lambda: ops.gemm(x4, x, a_major=1, b_major=1, out_f32=dw_1, accumulate=True), 2.0 * M * 4 * D * D),
        "wgrad_fc2 [768,3072,M]": (lambda: ops.gemm(x, x4, a_major=1, b_major=1, out_f32=dw_2, accumulate=True), 2.0 * M * 4 * D * D),
        "wgrad_proj[768,768,M]": (lambda: ops.gemm(x, x, a_major=1, b_major=1, out_f32=dw_p, accumulate=True), 2.0 * M * D * D),
    }
    ref_ms = timeit(lambda: torch.matmul(x, wqkv.t()))
    print(json.dumps({"case": "torch.matmul (cuBLAS) [M,2304,768] bf16", "ms": ref_ms, "tflops": 2.0 * M * 3 * D * D / ref_ms / 1e9}))
    for name, (fn, flops) in cases.items():
        if only and only not in name:
            continue
        ms = timeit(fn)
        print(json.dumps({"case": name, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
