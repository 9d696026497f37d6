"""clock64 timeline of the pipelined space-attention backward (build with OAT_SPACE_DBG=1): CTA 0, groups 1..4."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oa_transformer_b200 import ops
from oa_transformer_b200._lib import lib
BF = torch.bfloat16
B, F, n, H = 32, 8, 232, 12
T = 1 + F * n; M = B * T
qkv = (torch.randn(M, 3 * H * 64, device="cuda") * 0.5).to(BF)
out = torch.empty(M, H * 64, device="cuda", dtype=BF)
lse = torch.empty(B * H * T, device="cuda")
dout = torch.randn(M, H * 64, device="cuda").to(BF)
dqkv = torch.empty_like(qkv)
acc = torch.empty(B * H * 192, device="cuda")
ws = torch.zeros(ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F, n), device="cuda")
ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws)
for _ in range(3):
    ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
lib().oat_debug_timeline(buf, 1024)
print("group start-to-start cycles (CTA 0):", [buf[(i + 1) * 128] - buf[i * 128] for i in range(0, 7)])
if os.environ.get("LIGHT"):
    sys.exit(0)
for i in range(1, 5):
    t0 = buf[i * 128 + 0]
    b = lambda k, j=i: buf[j * 128 + k] - t0
    print("group %d (t=0: MMA warp starts S/dP of sub-unit 0; previous group's t0 at %d)" % (i, buf[(i - 1) * 128] - t0))
    for v in range(8):
        print("  v%d: MMA sdp start=%6d ops ready=%6d | grads: wait math from=%6d seen=%6d || math: wait from=%6d st_full seen=%6d done=%6d (math %d)" % (
            v, b(v * 4), b(v * 4 + 1), b(v * 4 + 2), b(v * 4 + 3), b(32 + v * 3), b(33 + v * 3), b(34 + v * 3), b(34 + v * 3) - b(33 + v * 3)))
    for v in range(8):
        print("  v%d MMA thread: math_done seen %6d | free-waits done %6d | dV/dK issued %6d | dQ+commits issued %6d || sdp(v): start %6d ops ready %6d issued %6d" % (
            v, b(v * 4 + 3), b(104 + v), b(72 + v), b(80 + v), b(v * 4), b(v * 4 + 1), b(88 + v)))
    print("  producer: empty A/B/C seen at %d %d %d | delta warps: start %d done %d" % (b(100), b(101), b(102), b(60), b(61)))
    print("  epilogue: kt0 acc_full %d read %d stored %d | kt1 %d %d %d | dq_full %d stored %d" % (
        b(64), b(65), b(66), b(67), b(68), b(69), b(70), b(71)))
