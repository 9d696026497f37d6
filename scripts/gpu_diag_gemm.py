"""GPU diagnostic for the tcgen05 GEMM: each case runs in its own subprocess (a trap must not poison the others)
and prints error statistics that localise layout mistakes (per 8-row / 64-col block error maps)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # (M, N, K, a_major, b_major)
    (128, 256, 64, 0, 0),
    (128, 128, 64, 0, 0),
    (128, 256, 256, 0, 0),
    (300, 768, 768, 0, 0),
    (128, 256, 64, 0, 1),
    (128, 256, 64, 1, 0),
    (128, 256, 64, 1, 1),
    (128, 128, 64, 1, 1),
    (300, 768, 768, 0, 1),
    (300, 768, 768, 1, 1),
    (1857, 2304, 768, 0, 0),
    (59424, 2304, 768, 0, 0),
]


def run_case(i):
    import torch
    from oa_transformer_b200 import ops
    M, N, K, am, bm = CASES[i]
    g = torch.Generator().manual_seed(i)
    Mp = M + (-M) % 8
    A = torch.randn((M, K) if am == 0 else (K, Mp), generator=g).to(torch.bfloat16).cuda()
    if am == 1:
        A = A[:, :M]
    B = torch.randn((N, K) if bm == 0 else (K, N), generator=g).to(torch.bfloat16).cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(A, B, a_major=am, b_major=bm, out_f32=out)
    torch.cuda.synchronize()
    a = A.float() if am == 0 else A.float().t()
    b = B.float() if bm == 0 else B.float().t()
    ref = a @ b.t()
    diff = (out - ref).abs()
    nan = torch.isnan(out).sum().item()
    res = {"case": CASES[i], "max_err": diff[~torch.isnan(diff)].max().item() if nan < out.numel() else None,
           "ref_max": ref.abs().max().item(), "nan": nan}
    if M <= 512 and (res["max_err"] is None or res["max_err"] > 1e-2 * res["ref_max"]):
        d = torch.nan_to_num(diff, nan=1e9)
        rows = ((M + 7) // 8) * 8
        dp = torch.zeros(rows, ((N + 63) // 64) * 64, device="cuda")
        dp[:M, :N] = d
        blk = dp.view(rows // 8, 8, dp.shape[1] // 64, 64).amax(dim=(1, 3))
        res["bad_blocks_8x64"] = (blk > 1e-2 * res["ref_max"]).int().cpu().tolist()[:20]
    if M >= 4096:
        import time
        for _ in range(3):
            ops.gemm(A, B, a_major=am, b_major=bm, out_f32=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(A, B, a_major=am, b_major=bm, out_f32=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
    else:
        for i in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=180)
                print("case %d rc=%d %s %s" % (i, r.returncode, r.stdout.strip()[-2000:], r.stderr.strip()[-600:]))
            except subprocess.TimeoutExpired:
                print("case %d TIMEOUT %s" % (i, CASES[i]))
            sys.stdout.flush()
