"""Launch every hot kernel of one training step ONCE at the benchmark geometry (B=32, F=8, n=232, H=12 -> M=59424 token
rows), for `ncu --set full` captures:

  ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attn_|layernorm_|colsum' \
      -o gpurun_out/kernels python scripts/gpu_kernels_once.py

KERNELS=name1,name2 restricts the list. Nothing here is timed: numbers printed under a profiler are never bench values.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oa_transformer_b200 import ops  # noqa: E402

BF = torch.bfloat16
B, F, n, H = int(os.environ.get("ATT_B", 32)), 8, 232, 12
T = 1 + F * n
M = B * T
D = 768
dev = "cuda"
only = set(filter(None, os.environ.get("KERNELS", "").split(",")))

x = (torch.randn(M, D, device=dev) * 0.5).to(BF)
x3 = (torch.randn(M, 3 * D, device=dev) * 0.5).to(BF)
x4 = (torch.randn(M, 4 * D, device=dev) * 0.5).to(BF)
res = torch.randn(M, D, device=dev)
o32 = torch.empty(M, D, device=dev)
o16_1 = torch.empty(M, D, device=dev, dtype=BF)
o16_3 = torch.empty(M, 3 * D, device=dev, dtype=BF)
o16_4 = torch.empty(M, 4 * D, device=dev, dtype=BF)
o16_4b = torch.empty(M, 4 * D, device=dev, dtype=BF)
wqkv = (torch.randn(3 * D, D, device=dev) * 0.02).to(BF)
wproj = (torch.randn(D, D, device=dev) * 0.02).to(BF)
w1 = (torch.randn(4 * D, D, device=dev) * 0.02).to(BF)
w2 = (torch.randn(D, 4 * D, device=dev) * 0.02).to(BF)
b1 = torch.randn(D, device=dev)
b3 = torch.randn(3 * D, device=dev)
b4 = torch.randn(4 * D, device=dev)
dw3 = torch.zeros(3 * D, D, device=dev)
dw4 = torch.zeros(4 * D, D, device=dev)
dw2 = torch.zeros(D, 4 * D, device=dev)
dwp = torch.zeros(D, D, device=dev)
lse = torch.empty(B * H * T, device=dev)
acc = torch.empty(B * H * 192, device=dev)
ws = torch.zeros(max(1, ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F, n),
                     ops.attn_fwd_workspace_floats(ops.MODE_TIME, B, H, F, n)), device=dev)
gamma = torch.ones(D, device=dev)
beta = torch.zeros(D, device=dev)
mean = torch.empty(M, device=dev)
rstd = torch.empty(M, device=dev)
dgam = torch.zeros(D, device=dev)
dbet = torch.zeros(D, device=dev)
dxs = torch.zeros(D, device=dev)
dx32 = torch.empty(M, D, device=dev)
colacc = torch.zeros(3 * D, device=dev)
o16_1b = (torch.randn(M, D, device=dev) * 0.5).to(BF)
delta = torch.zeros(H, M, device=dev)

cases = [
    ("gemm_fwd_qkv", lambda: ops.gemm(x, wqkv, bias=b3, scale_cols=D, scale=0.125, out_bf16=o16_3)),
    ("gemm_fwd_proj", lambda: ops.gemm(x, wproj, bias=b1, residual=res, out_f32=o32)),
    ("gemm_fwd_fc1", lambda: ops.gemm(x, w1, bias=b4, act=ops.ACT_GELU, out_bf16=o16_4, out2_bf16=o16_4b)),
    ("gemm_fwd_fc2", lambda: ops.gemm(x4, w2, bias=b1, residual=res, out_f32=o32)),
    ("gemm_dgrad_fc2", lambda: ops.gemm(x, w2, b_major=1, act=ops.ACT_GELU_BWD, aux=o16_4b, out_bf16=o16_4)),
    ("gemm_dgrad_fc1", lambda: ops.gemm(x4, w1, b_major=1, out_bf16=o16_1)),
    ("gemm_dgrad_qkv", lambda: ops.gemm(x3, wqkv, b_major=1, out_bf16=o16_1)),
    ("gemm_dgrad_proj", lambda: ops.gemm(x, wproj, b_major=1, out_bf16=o16_1)),
    ("gemm_wgrad_proj", lambda: ops.gemm(x, x, a_major=1, b_major=1, out_f32=dwp, accumulate=True)),
    ("gemm_wgrad_qkv", lambda: ops.gemm(x3, x, a_major=1, b_major=1, out_f32=dw3, accumulate=True)),
    ("gemm_wgrad_fc1", lambda: ops.gemm(x4, x, a_major=1, b_major=1, out_f32=dw4, accumulate=True)),
    ("gemm_wgrad_fc2", lambda: ops.gemm(x, x4, a_major=1, b_major=1, out_f32=dw2, accumulate=True)),
    ("attn_space_fwd", lambda: ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, x3, o16_1, lse, None, cls_ws=ws)),
    ("attn_space_bwd", lambda: ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, x3, o16_1, lse, x, o16_3, 0.125, acc)),
    ("attn_time_fwd", lambda: ops.attn_fwd(ops.MODE_TIME, B, T, H, F, n, x3, o16_1, lse, None, cls_ws=ws)),
    ("attn_time_bwd", lambda: ops.attn_bwd(ops.MODE_TIME, B, T, H, F, n, x3, o16_1, lse, x, o16_3, 0.125, acc)),
    ("layernorm_fwd", lambda: ops.layernorm_fwd(res, gamma, beta, 1e-6, y_bf16=o16_1, mean=mean, rstd=rstd)),
    ("layernorm_bwd", lambda: ops.layernorm_bwd(res, mean, rstd, gamma, dy_bf16=x, add1=o32, dx=dx32, dx_bf16=o16_1,
                                                dgamma=dgam, dbeta=dbet, dxsum=dxs)),
    ("colsum", lambda: ops.colsum_bf16(x3, colacc)),
    # delta = rowsum(dO * O) from the projection dgrad's epilogue, and the attention backward kernels fed with it
    ("gemm_dgrad_proj_rowdot", lambda: ops.gemm(x, wproj, b_major=1, act=ops.ACT_ROWDOT, aux=o16_1b, rowdot=delta, out_bf16=o16_1)),
    ("attn_space_bwd_delta", lambda: ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, x3, o16_1b, lse, x, o16_3, 0.125, acc, delta=delta)),
    ("attn_time_bwd_delta", lambda: ops.attn_bwd(ops.MODE_TIME, B, T, H, F, n, x3, o16_1b, lse, x, o16_3, 0.125, acc, delta=delta)),
]
torch.cuda.synchronize()
for name, fn in cases:
    if only and name not in only:
        continue
    fn()
    torch.cuda.synchronize()
    print("ran", name, flush=True)
