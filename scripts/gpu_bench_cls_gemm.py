"""CUDA-event timing of the few-row (CLS / projection) products at the benchmark geometry (B = 32 rows, T = 1857), through
oat_gemm_bf16 with the skinny kernel on and off (OAT_GEMM_SKINNY is read once per process: run twice)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oa_transformer_b200 import ops
BF = torch.bfloat16
B, T, D = 32, 1857, 768
dev = "cuda"
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # a large write between launches keeps the weights out of L2, as in the real step
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3
a3 = torch.randn(B, 3 * D, device=dev).to(BF)
w_qkv3 = torch.randn(3 * D, 3 * D, device=dev).to(BF)
w_fc13 = torch.randn(4 * D, 3 * D, device=dev).to(BF)
w_fc23 = torch.randn(D, 12 * D, device=dev).to(BF)
w_proj3 = torch.randn(D, 3 * D, device=dev).to(BF)
g3 = torch.randn(B, 12 * D, device=dev).to(BF)
qkv = torch.zeros(B * T, 3 * D, device=dev, dtype=BF)
x = torch.zeros(B * T, D, device=dev)
abuf = torch.randn(B * T, D, device=dev).to(BF)
g = torch.zeros(B * T, 4 * D, device=dev, dtype=BF)
u = torch.zeros(B * T, 4 * D, device=dev, dtype=BF)
g32 = torch.zeros(B, 4 * D, device=dev)
bias3 = torch.zeros(3 * D, device=dev); bias4 = torch.zeros(4 * D, device=dev)
cls = lambda tt: tt.view(B, T * tt.shape[1])[:, :tt.shape[1]]
res = {"skinny": os.environ.get("OAT_GEMM_SKINNY", "1")}
res["qkv_cls_us"] = t(lambda: ops.gemm(a3, w_qkv3, bias=bias3, scale_cols=D, scale=0.125, out_bf16=cls(qkv)))
res["proj_lo_us"] = t(lambda: ops.gemm(cls(abuf), w_proj3[:, D:2 * D], out_f32=cls(x), accumulate=True))
res["fc1_cls_us"] = t(lambda: ops.gemm(a3, w_fc13, bias=bias4, act=ops.ACT_GELU, out_f32=g32, out_bf16=cls(g), out2_bf16=cls(u)))
res["fc2_corr_us"] = t(lambda: ops.gemm(g3[:, 4 * D:], w_fc23[:, 4 * D:], out_f32=cls(x), accumulate=True))
res["per_block_us"] = 2 * res["qkv_cls_us"] + 2 * res["proj_lo_us"] + res["fc1_cls_us"] + res["fc2_corr_us"]
print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}))
