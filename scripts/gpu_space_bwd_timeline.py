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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc)
e1.record(); torch.cuda.synchronize()
print("event time per call: %.1f us" % (e0.elapsed_time(e1) * 100))
buf = (ctypes.c_longlong * 4400)()
lib().oat_debug_timeline(buf, 4400)
cyc = sorted(buf[4096 + k] for k in range(148))
print('kernel cycles per CTA: min %d median %d max %d' % (cyc[0], cyc[74], cyc[-1]))
print('per CTA (k):', [int(buf[4096 + k] // 1000) for k in range(148)])
print("group start-to-start cycles (CTA 0):", [buf[(i + 1) * 128] - buf[i * 128] for i in range(0, 7)])
print("CTA 0 milestones (cycles from CTA start): zero-fill %d | barriers+TMEM ready %d | pdl_wait %d | first group fetched %d | MMA warp has group %d | first S/dP issued %d | last MMA issued %d | epilogue done %d | stores drained %d | exit %d" % tuple([buf[3900 + k] for k in range(9)] + [buf[4096]]))
if os.environ.get("LIGHT"):
    print("all groups:", [int(buf[(i + 1) * 128] - buf[i * 128]) for i in range(0, 24)])
if os.environ.get("LIGHT"):
    sys.exit(0)
for i in range(1, 5):
    t0 = buf[i * 128 + 0]
    b = lambda k, j=i: buf[j * 128 + k] - t0
    print("group %d (t=0: MMA warp starts S/dP of sub-unit 0; previous group's t0 at %d)" % (i, buf[(i - 1) * 128] - t0))
    for v in range(8):
        print("  v%d: MMA thread: waits for math from=%6d seen=%6d all waits done=%6d MMAs issued=%6d || math warp 4 done=%6d" % (
            v, b(v * 4 + 2), b(v * 4 + 3), b(104 + v), b(80 + v), b(34 + v * 3)))
    print("  producer: empty A/B/C seen at %d %d %d | delta warps: start %d done %d" % (b(100), b(101), b(102), b(60), b(61)))
    print("  epilogue: kt0 acc_full %d read %d stored %d | kt1 %d %d %d | dq_full %d stored %d" % (
        b(64), b(65), b(66), b(67), b(68), b(69), b(70), b(71)))
