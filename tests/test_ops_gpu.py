"""Per-kernel parity on the GPU: each liboat op (through the C ABI) vs the CPU oracle / golden fixtures on the same
seeded inputs. bf16 tensor-core ops are compared with the oracle in its bf16-operand mode (gate ~1e-3 relative);
integer / index bookkeeping must be bit-exact."""
import os

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BF = torch.bfloat16


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,D,eps", [(1000, 768, 1e-6), (37, 768, 1e-12), (64, 128, 1e-6)])
def test_layernorm_fwd_bwd(rows, D, eps):
    from oa_transformer_b200 import ops
    g = gen(1)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.5).cuda()
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).cuda()
    beta = (0.1 * torch.randn(D, generator=g)).cuda()
    y16 = torch.empty(rows, D, device="cuda", dtype=BF)
    y32 = torch.empty(rows, D, device="cuda")
    mean = torch.empty(rows, device="cuda")
    rstd = torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, eps, y_bf16=y16, y_f32=y32, mean=mean, rstd=rstd)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, eps)
    assert torch.allclose(y32, ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(y16.float(), ref, rtol=1e-2, atol=1e-2)
    dy16 = torch.randn(rows, D, generator=g).to(BF).cuda()
    dy32 = torch.randn(rows, D, generator=g).cuda()
    add1 = torch.randn(rows, D, generator=g).cuda()
    add2 = torch.randn(rows, D, generator=g).cuda()
    dx = torch.empty(rows, D, device="cuda")
    dx16 = torch.empty(rows, D, device="cuda", dtype=BF)
    dgamma = torch.zeros(D, device="cuda")
    dbeta = torch.zeros(D, device="cuda")
    dxsum = torch.zeros(D, device="cuda")
    ops.layernorm_bwd(x, mean, rstd, gamma, dy_bf16=dy16, dy_f32=dy32, add1=add1, add2=add2, dx=dx, dx_bf16=dx16,
                      dgamma=dgamma, dbeta=dbeta, dxsum=dxsum)
    ref.backward(dy16.float() + dy32)
    assert rel(dx, xr.grad + add1 + add2) < 1e-5
    assert rel(dx16.float(), xr.grad + add1 + add2) < 5e-3
    assert rel(dgamma, gr.grad) < 1e-4 and rel(dbeta, br.grad) < 1e-4
    assert rel(dxsum, dx.sum(0)) < 1e-4


def _split_ref(v):
    hi = v.to(BF)
    return hi, (v - hi.float()).to(BF)


def test_split_bf16_operands_bit_exact_and_three_term_gemm():
    """Split-bf16 operands (include/oat.h, oat_split3_bf16 / oat_cast_multi kind 2 / oat_layernorm_fwd y_split): the
    packing is pure rounding bookkeeping -> bit-exact against torch; the K-concatenated GEMM [hi|hi|lo] x [hi|lo|hi]
    must agree with the fp64 product far below one bf16 ulp (2^-9), which a single bf16 GEMM cannot."""
    from oa_transformer_b200 import ops
    g = gen(11)
    rows, K, N, T = 24, 768, 256, 5
    x = torch.randn(rows * T, K, generator=g).cuda()
    w = (0.05 * torch.randn(N, K, generator=g)).cuda()
    # activation layout from the strided rows 0, T, 2T, ... (the CLS rows), with ReLU
    a3 = torch.empty(rows, 3 * K, device="cuda", dtype=BF)
    ops.split3_bf16(x, a3, rows=rows, cols=K, lds=T * K, relu=True)
    hi, lo = _split_ref(x[::T].clamp_min(0))
    assert torch.equal(a3[:, :K], hi) and torch.equal(a3[:, K:2 * K], hi) and torch.equal(a3[:, 2 * K:], lo)
    # weight layout through the one-launch cast plan
    w3 = torch.zeros(N, 3 * K, device="cuda", dtype=BF)
    plan = ops.CastPlan()
    plan.add(w, w3, split=True)
    plan.run()
    whi, wlo = _split_ref(w)
    assert torch.equal(w3[:, :K], whi) and torch.equal(w3[:, K:2 * K], wlo) and torch.equal(w3[:, 2 * K:], whi)
    # LayerNorm forward emitting the split copy of every T-th row
    gamma = (1 + 0.1 * torch.randn(K, generator=g)).cuda()
    beta = (0.1 * torch.randn(K, generator=g)).cuda()
    y32 = torch.empty(rows * T, K, device="cuda")
    y16 = torch.empty(rows * T, K, device="cuda", dtype=BF)
    c3 = torch.empty(rows, 3 * K, device="cuda", dtype=BF)
    ops.layernorm_fwd(x, gamma, beta, 1e-6, y_bf16=y16, y_f32=y32, y_split=c3, split_period=T)
    hi, lo = _split_ref(y32[::T])
    assert torch.equal(c3[:, :K], hi) and torch.equal(c3[:, K:2 * K], hi) and torch.equal(c3[:, 2 * K:], lo)
    assert torch.equal(y16[::T], hi)
    # three-term product vs fp64, and the correction form  hi.hi (plain GEMM) + [hi|lo] x [lo|hi] (accumulated)
    xr = x[::T].clamp_min(0)
    exact = (xr.double() @ w.double().t()).float()
    out3 = torch.empty(rows, N, device="cuda")
    ops.gemm(a3, w3, out_f32=out3)
    out1 = torch.empty(rows, N, device="cuda")
    ops.gemm(a3[:, :K], w3[:, :K], out_f32=out1)
    assert rel(out3, exact) < 3e-5 and rel(out1, exact) > 1e-3
    ops.gemm(a3[:, K:], w3[:, K:], out_f32=out1, accumulate=True)
    assert rel(out1, exact) < 3e-5
    # strided output rows (the CLS rows of a token buffer) for both the overwrite and the accumulate form
    big = torch.zeros(rows * T, N, device="cuda")
    cls_rows = big.view(rows, T * N)[:, :N]
    ops.gemm(a3, w3, out_f32=cls_rows)
    assert rel(big[::T], exact) < 3e-5 and float(big[1::T].abs().max()) == 0.0
    big16 = torch.zeros(rows * T, N, device="cuda", dtype=BF)
    ops.gemm(a3, w3, out_bf16=big16.view(rows, T * N)[:, :N])
    assert rel(big16[::T].float(), exact) < 4e-3 and float(big16[1::T].float().abs().max()) == 0.0
    ops.gemm(a3[:, K:], w3[:, K:], out_f32=cls_rows, accumulate=True)
    torch.cuda.synchronize()
    assert float(big[1::T].abs().max()) == 0.0


@pytest.mark.parametrize("variant", ["adds", "add1", "plain", "text"])
@pytest.mark.parametrize("rows,D", [(5000, 768), (8 * 148 + 3, 768), (100, 256)])
def test_layernorm_bwd_staged_variants(rows, D, variant):
    """The cp.async double-buffered backward (the operand combinations the two towers use) vs torch autograd."""
    from oa_transformer_b200 import ops
    g = gen(2)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.5).cuda()
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).cuda()
    beta = (0.1 * torch.randn(D, generator=g)).cuda()
    mean = torch.empty(rows, device="cuda")
    rstd = torch.empty(rows, device="cuda")
    y16 = torch.empty(rows, D, device="cuda", dtype=BF)
    ops.layernorm_fwd(x, gamma, beta, 1e-6, y_bf16=y16, mean=mean, rstd=rstd)
    dy16 = torch.randn(rows, D, generator=g).to(BF).cuda()
    dy32 = torch.randn(rows, D, generator=g).cuda() if variant == "text" else None
    add1 = torch.randn(rows, D, generator=g).cuda() if variant in ("adds", "add1") else None
    add2 = torch.randn(rows, D, generator=g).cuda() if variant == "adds" else None
    dx = torch.empty(rows, D, device="cuda")
    dx16 = torch.empty(rows, D, device="cuda", dtype=BF)
    dgamma, dbeta, dxsum = (torch.zeros(D, device="cuda") for _ in range(3))
    ops.layernorm_bwd(x, mean, rstd, gamma, dy_bf16=dy16, dy_f32=dy32, add1=add1, add2=add2, dx=dx, dx_bf16=dx16,
                      dgamma=dgamma, dbeta=dbeta, dxsum=dxsum)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-6)
    ref.backward(dy16.float() + (dy32 if dy32 is not None else 0))
    want = xr.grad + (add1 if add1 is not None else 0) + (add2 if add2 is not None else 0)
    assert rel(dx, want) < 1e-5
    assert rel(dx16.float(), want) < 5e-3
    assert rel(dgamma, gr.grad) < 1e-4 and rel(dbeta, br.grad) < 1e-4
    assert rel(dxsum, dx.sum(0)) < 1e-4


def test_layernorm_strided_cls_rows():
    from oa_transformer_b200 import ops
    B, T, D = 5, 33, 768
    x = torch.randn(B * T, D, generator=gen(2)).cuda()
    gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    y32 = torch.empty(B, D, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, 1e-6, rows=B, ldx=T * D, y_f32=y32)
    ref = torch.nn.functional.layer_norm(x.view(B, T, D)[:, 0], (D,), gamma, beta, 1e-6)
    assert torch.allclose(y32, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B", [5, 32, 40])
def test_layernorm_bwd_strided_cls_rows(B):
    """Final-LayerNorm backward on the CLS rows only (row pitch T*D), as the video engine calls it: B = 32 / 40 take the
    staged kernel, B = 5 the register kernel; rows between the CLS rows must stay untouched."""
    from oa_transformer_b200 import ops
    T, D = 7, 768
    g = gen(6)
    x = torch.randn(B * T, D, generator=g).cuda()
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).cuda()
    beta = torch.zeros(D, device="cuda")
    mean = torch.empty(B, device="cuda")
    rstd = torch.empty(B, device="cuda")
    y16 = torch.empty(B, D, device="cuda", dtype=BF)
    ops.layernorm_fwd(x, gamma, beta, 1e-6, rows=B, ldx=T * D, y_bf16=y16, mean=mean, rstd=rstd)
    dy16 = torch.randn(B, D, generator=g).to(BF).cuda()
    dx = torch.full((B * T, D), 3.0, device="cuda")
    dx16 = torch.full((B * T, D), 3.0, device="cuda", dtype=BF)
    dgamma, dbeta, dxsum = (torch.zeros(D, device="cuda") for _ in range(3))
    ops.layernorm_bwd(x, mean, rstd, gamma, dy_bf16=dy16, rows=B, ldx=T * D, dx=dx, dx_bf16=dx16, lddx=T * D, lddxb=T * D,
                      dgamma=dgamma, dbeta=dbeta, dxsum=dxsum)
    xr = x.view(B, T, D)[:, 0].clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), gr, beta, 1e-6).backward(dy16.float())
    got = dx.view(B, T, D)
    assert rel(got[:, 0], xr.grad) < 1e-5
    assert bool((got[:, 1:] == 3.0).all()) and bool((dx16.view(B, T, D)[:, 1:].float() == 3.0).all())
    assert rel(dx16.view(B, T, D)[:, 0].float(), xr.grad) < 5e-3
    assert rel(dgamma, gr.grad) < 1e-4 and rel(dxsum, xr.grad.sum(0)) < 1e-4


# ------------------------------------------------------------------------------------------------ attention
def _qkv(B, T, H, seed, scale=1.0):
    q = torch.randn(B, T, 3, H, 64, generator=gen(seed)) * scale
    q[:, :, 0] *= 0.125      # the q slot holds the pre-scaled query (video_transformer.py:105)
    return q.to(BF)


def _run_attn(mode, B, F, n, H, qkv16, dout16, key_mask=None, fwd_ws=True):
    from oa_transformer_b200 import ops
    T = qkv16.shape[1]
    qkv = qkv16.reshape(B * T, 3 * H * 64).cuda()
    out = torch.zeros(B * T, H * 64, device="cuda", dtype=BF)
    lse = torch.zeros(B * H * T, device="cuda")
    km = key_mask.to(torch.int32).cuda().contiguous() if key_mask is not None else None
    ws = torch.empty(max(1, ops.attn_fwd_workspace_floats(mode, B, H, F, n)), device="cuda") if fwd_ws else None
    ops.attn_fwd(mode, B, T, H, F, n, qkv, out, lse, km, cls_ws=ws)
    dqkv = torch.zeros_like(qkv)
    acc = torch.empty(B * H * 3 * 64, device="cuda") if mode != ops.MODE_PLAIN else None
    ops.attn_bwd(mode, B, T, H, F, n, qkv, out, lse, dout16.reshape(B * T, H * 64).cuda(), dqkv, 0.125, acc, km)
    torch.cuda.synchronize()
    return out.cpu().float().view(B, T, H * 64), dqkv.cpu().float().view(B, T, 3, H, 64)


@pytest.mark.parametrize("mode,B,F,n,H", [("space", 2, 2, 16, 2), ("time", 2, 2, 16, 2), ("space", 2, 3, 232, 3),
                                          ("time", 1, 8, 40, 2), ("time", 1, 16, 9, 1), ("space", 1, 2, 196, 12),
                                          ("time", 1, 4, 232, 12), ("space", 4, 4, 232, 12), ("space", 1, 2, 128, 2),
                                          ("space", 1, 3, 255, 2), ("time", 2, 8, 232, 12), ("time", 1, 6, 19, 2),
                                          ("time", 3, 4, 7, 1), ("time", 1, 16, 232, 2), ("time", 2, 3, 33, 3),
                                          ("time", 2, 1, 4, 2), ("space", 2, 1, 4, 2), ("time", 2, 1, 196, 12),
                                          ("space", 2, 1, 196, 12)])
def test_divided_attention_fwd_bwd(mode, B, F, n, H):
    from oa_transformer_b200 import ops
    T = 1 + F * n
    qkv16 = _qkv(B, T, H, 3)
    dout16 = torch.randn(B, T, H * 64, generator=gen(4)).to(BF)
    out, dqkv = _run_attn(ops.MODE_SPACE if mode == "space" else ops.MODE_TIME, B, F, n, H, qkv16, dout16)
    # oracle: q in the buffer is the already-scaled query; the kernel returns d/d(unscaled q) = scale * d/d(q)
    cfg = O.OracleCfg(heads=H, bf16=True)
    x = qkv16.float().requires_grad_(True)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = O.divided_attention_core(q, k, v, mode, F, n, cfg)
    ref.backward(dout16.float())
    gref = x.grad.clone()
    gref[:, :, 0] *= 0.125
    assert rel(out, ref.detach()) < 4e-3, rel(out, ref.detach())
    for i, name in enumerate("qkv"):
        assert rel(dqkv[:, :, i], gref[:, :, i]) < 6e-3, (name, rel(dqkv[:, :, i], gref[:, :, i]))
    # the CLS row on its own (cross-group atomics path)
    assert rel(dqkv[:, 0], gref[:, 0]) < 6e-3, rel(dqkv[:, 0], gref[:, 0])
    assert rel(out[:, 0], ref.detach()[:, 0]) < 4e-3


@pytest.mark.parametrize("mode,B,F,n,H", [("space", 2, 3, 232, 3), ("space", 4, 4, 232, 12), ("space", 1, 3, 255, 2),
                                          ("time", 2, 8, 232, 12), ("time", 1, 6, 19, 2), ("time", 2, 1, 196, 12),
                                          ("space", 2, 2, 16, 2)])
def test_attention_bwd_with_delta_from_the_gemm_epilogue(mode, B, F, n, H):
    """oat_attn_args.delta: rowsum(dO * O) per head handed in (as the dO-producing GEMM's act-4 epilogue writes it, layout
    [H, ld >= B*T]) must give the gradients the kernels get when they compute delta from O themselves."""
    from oa_transformer_b200 import ops
    m = ops.MODE_SPACE if mode == "space" else ops.MODE_TIME
    T = 1 + F * n
    qkv = _qkv(B, T, H, 31).reshape(B * T, 3 * H * 64).cuda()
    dout = torch.randn(B * T, H * 64, generator=gen(32)).to(BF).cuda()
    out = torch.zeros(B * T, H * 64, device="cuda", dtype=BF)
    lse = torch.zeros(B * H * T, device="cuda")
    ws = torch.empty(max(1, ops.attn_fwd_workspace_floats(m, B, H, F, n)), device="cuda")
    ops.attn_fwd(m, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws)
    g0, g1 = torch.zeros_like(qkv), torch.zeros_like(qkv)
    acc = torch.empty(B * H * 3 * 64, device="cuda")
    ops.attn_bwd(m, B, T, H, F, n, qkv, out, lse, dout, g0, 0.125, acc)
    ld = B * T + 3
    delta = torch.full((H, ld), float("nan"), device="cuda")
    delta[:, :B * T] = (dout.float() * out.float()).view(B * T, H, 64).sum(-1).t()
    # O poisoned: with delta given the patch rows must not depend on it any more (the CLS-row finalize still reads row 0)
    out_p = out.clone()
    out_p.view(B, T, H * 64)[:, 1:] = 1e4
    ops.attn_bwd(m, B, T, H, F, n, qkv, out_p, lse, dout, g1, 0.125, acc, delta=delta)
    torch.cuda.synchronize()
    uses_delta = mode == "time" or n >= 128          # small space groups run the generic kernel, which ignores delta
    if uses_delta:
        assert rel(g1.float(), g0.float()) < 1e-3, rel(g1.float(), g0.float())
    else:
        ops.attn_bwd(m, B, T, H, F, n, qkv, out, lse, dout, g1, 0.125, acc, delta=delta)
        torch.cuda.synchronize()
        assert torch.equal(g1, g0)


@pytest.mark.parametrize("B,F,n,H", [(2, 8, 232, 12), (1, 5, 21, 2), (2, 16, 40, 1)])
def test_time_attention_fused_cls_matches_separate_pass(B, F, n, H):
    """The CLS query fused into the time kernel (per-warp partials + combine) vs the separate all-keys pass."""
    from oa_transformer_b200 import ops
    T = 1 + F * n
    qkv16 = _qkv(B, T, H, 12, scale=1.5)
    dout16 = torch.randn(B, T, H * 64, generator=gen(5)).to(BF)
    o1, g1 = _run_attn(ops.MODE_TIME, B, F, n, H, qkv16, dout16, fwd_ws=True)
    o2, g2 = _run_attn(ops.MODE_TIME, B, F, n, H, qkv16, dout16, fwd_ws=False)
    assert torch.equal(o1[:, 1:], o2[:, 1:])
    assert rel(o1[:, 0], o2[:, 0]) < 4e-3, rel(o1[:, 0], o2[:, 0])
    assert rel(g1, g2) < 2e-3, rel(g1, g2)


@pytest.mark.parametrize("B,F,n,H", [(1, 1, 232, 1), (4, 4, 232, 12), (2, 3, 196, 5), (1, 2, 128, 2), (1, 2, 255, 3),
                                     (1, 3, 160, 2)])
def test_space_attention_tcgen05_fwd(B, F, n, H):
    """tcgen05/TMEM space kernel (CLS query fused) vs the oracle and vs the mma.sync kernel it replaces; more groups than
    SMs in one case so the persistent 2-stage ring wraps."""
    from oa_transformer_b200 import ops
    T = 1 + F * n
    qkv16 = _qkv(B, T, H, 11, scale=1.5)
    qkv = qkv16.reshape(B * T, 3 * H * 64).cuda()
    outs = []
    for use_ws in (True, False):
        out = torch.zeros(B * T, H * 64, device="cuda", dtype=BF)
        lse = torch.zeros(B * H * T, device="cuda")
        ws = torch.empty(ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F, n), device="cuda") if use_ws else None
        ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws)
        torch.cuda.synchronize()
        outs.append((out.cpu().float().view(B, T, H * 64), lse.cpu().view(B, H, T)))
    cfg = O.OracleCfg(heads=H, bf16=True)
    q, k, v = (qkv16.float()[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = O.divided_attention_core(q, k, v, "space", F, n, cfg)
    (o_tc, l_tc), (o_old, l_old) = outs
    assert rel(o_tc, ref) < 4e-3, rel(o_tc, ref)
    assert rel(o_tc[:, 0], ref[:, 0]) < 4e-3, rel(o_tc[:, 0], ref[:, 0])
    assert rel(o_tc, o_old) < 4e-3
    assert (l_tc - l_old).abs().max() < 2e-3, (l_tc - l_old).abs().max()


@pytest.mark.parametrize("B,F,n,H", [(6, 8, 232, 12), (5, 5, 196, 7), (5, 7, 128, 9), (4, 6, 250, 8), (9, 4, 191, 5),
                                     (7, 4, 193, 6)])
def test_space_attention_tcgen05_bwd_pipelined_many_groups(B, F, n, H, monkeypatch):
    """The pipelined tcgen05 backward (dynamic group scheduler, operands released tile by tile, query halves visited in
    alternating order, dQ accumulators handed over between groups) with several groups per SM, launched three times in a
    row (the scheduler's counters re-arm themselves), against the mma.sync kernel on the same inputs. Geometries: the
    bench one, 64-query blocks cut by the CLS row at every position (n = 128, 191, 193, 196, 250)."""
    from oa_transformer_b200 import ops
    assert B * F * H > 148
    T = 1 + F * n
    qkv16 = _qkv(B, T, H, 31, scale=1.2)
    dout16 = torch.randn(B, T, H * 64, generator=gen(32)).to(BF)
    qkv = qkv16.reshape(B * T, 3 * H * 64).cuda()
    dout = dout16.reshape(B * T, H * 64).cuda()
    out = torch.zeros(B * T, H * 64, device="cuda", dtype=BF)
    lse = torch.zeros(B * H * T, device="cuda")
    ws = torch.empty(ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, F, n), device="cuda")
    ops.attn_fwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, None, cls_ws=ws)
    acc = torch.empty(B * H * 3 * 64, device="cuda")
    res = []
    for rep in range(3):
        dqkv = torch.full_like(qkv, float("nan"))
        ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, dout, dqkv, 0.125, acc)
        torch.cuda.synchronize()
        res.append(dqkv.float().view(B, T, 3, H * 64))
    monkeypatch.setenv("OAT_SPACE_BWD_LEGACY", "1")
    ref = torch.full_like(qkv, float("nan"))
    ops.attn_bwd(ops.MODE_SPACE, B, T, H, F, n, qkv, out, lse, dout, ref, 0.125, acc)
    torch.cuda.synchronize()
    ref = ref.float().view(B, T, 3, H * 64)
    assert torch.isfinite(ref).all()
    for r in res:
        assert torch.isfinite(r).all()          # every row of every group was written
        for i, name in enumerate("qkv"):
            assert rel(r[:, :, i], ref[:, :, i]) < 2e-3, (name, rel(r[:, :, i], ref[:, :, i]))
            assert rel(r[:, 0, i], ref[:, 0, i]) < 2e-3, (name, "cls", rel(r[:, 0, i], ref[:, 0, i]))
    # Launch-to-launch: the dynamic scheduler decides which SM iteration a group lands in, and with it the order in which
    # the group's query halves are accumulated into dV / dK (fp32, in TMEM): equal to fp32 rounding, not bit-identical.
    assert rel(res[0], res[1]) < 1e-4 and rel(res[0], res[2]) < 1e-4


@pytest.mark.parametrize("late", ["ramp", "spike", "cls"])
def test_space_attention_single_pass_softmax_reference_moves(late):
    """The forward softmax runs ONE pass over each score row with a running reference (first-chunk maximum, moved only
    when a later 32-key chunk exceeds it by > 2^8, re-scaling the P chunks already in TMEM). Rows whose large scores come
    late - a ramp over the keys, a single spike in the last chunk, a dominant CLS key (the LAST key column) - force that
    path; the result must still be the exact softmax, and backward (which recomputes P from the saved log-sum-exp) must
    agree with it."""
    from oa_transformer_b200 import ops
    B, F, n, H = 2, 2, 232, 2
    T = 1 + F * n
    g = gen(21)
    x = torch.randn(B, T, 3, H, 64, generator=g)
    x[:, :, 0] *= 0.125
    k = x[:, :, 1]
    if late == "ramp":            # key norms grow along the frame: every chunk raises the maximum by far more than 2^8
        ramp = torch.linspace(0.2, 6.0, n).repeat(F)
        k[:, 1:] *= ramp.view(1, -1, 1, 1)
        x[:, :, 0] *= 4.0
    elif late == "spike":         # one key near the end of every frame aligned with every query direction
        x[:, :, 0] = x[:, :, 0].abs() * 2.0
        k[:, 1:].view(B, F, n, H, 64)[:, :, n - 3] = 1.5
    else:                         # the CLS key (read as the last key column by the kernel) dominates
        x[:, :, 0] = x[:, :, 0].abs() * 2.0
        k[:, 0] = 1.5
    qkv16 = x.to(BF)
    dout16 = torch.randn(B, T, H * 64, generator=g).to(BF)
    out, dqkv = _run_attn(ops.MODE_SPACE, B, F, n, H, qkv16, dout16)
    cfg = O.OracleCfg(heads=H, bf16=True)
    xr = qkv16.float().requires_grad_(True)
    q, kk, v = (xr[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    s_max = float((q[:, :, 1:2] @ kk.transpose(-1, -2)).abs().max())
    assert s_max > 12.0, s_max                       # > 2^8 / log2(e) = 5.5: the late scores move the reference
    ref = O.divided_attention_core(q, kk, v, "space", F, n, cfg)
    ref.backward(dout16.float())
    gref = xr.grad.clone()
    gref[:, :, 0] *= 0.125
    assert torch.isfinite(out).all() and rel(out, ref.detach()) < 5e-3, rel(out, ref.detach())
    # backward recomputes P from the saved log-sum-exp of this forward. With a peaked softmax dS = P (dP - delta) is a
    # difference of nearly equal numbers (delta comes from the bf16-rounded output), so dq / dk are compared only in the
    # ramp case; dv = P^T dO is well conditioned in all three.
    for i, name in enumerate("qkv"):
        if late == "ramp" or name == "v":
            assert rel(dqkv[:, :, i], gref[:, :, i]) < 1.5e-2, (name, rel(dqkv[:, :, i], gref[:, :, i]))


@pytest.mark.parametrize("B,L,H", [(3, 32, 12), (2, 8, 2), (2, 50, 2)])
def test_text_attention_with_padding_mask(B, L, H):
    from oa_transformer_b200 import ops
    qkv16 = _qkv(B, L, H, 5)
    dout16 = torch.randn(B, L, H * 64, generator=gen(6)).to(BF)
    mask = torch.ones(B, L, dtype=torch.long)
    mask[1, L // 2:] = 0
    out, dqkv = _run_attn(ops.MODE_PLAIN, B, 0, 0, H, qkv16, dout16, key_mask=mask)
    cfg = O.OracleCfg(heads=H, bf16=True)
    x = qkv16.float().requires_grad_(True)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    add = torch.zeros(B, 1, 1, L).masked_fill(mask.view(B, 1, 1, L) == 0, torch.finfo(torch.float32).min)
    ref = O._softmax_attention(q, k, v, cfg, add).permute(0, 2, 1, 3).reshape(B, L, H * 64)
    ref.backward(dout16.float())
    gref = x.grad.clone()
    gref[:, :, 0] *= 0.125
    assert rel(out, ref.detach()) < 4e-3
    for i in range(3):
        assert rel(dqkv[:, :, i], gref[:, :, i]) < 6e-3, i
    # masked keys receive exactly zero gradient
    assert float(dqkv[1, L // 2:, 1:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ tokens
def test_patch_embed_and_token_assembly_with_objects():
    from oa_transformer_b200 import ops
    from oracle.weights import fill_seeded, video_tower_spec
    B, F, O_, D = 2, 3, 5, 768
    spec = {k: v for k, v in video_tower_spec(depth=0, frames=4, objects=True).items()}
    p = fill_seeded(spec, 7, 0.05)
    g = gen(8)
    video = torch.randn(B, F, 3, 224, 224, generator=g)
    objects = O.synth_objects(B, F, O_, g)
    cfg = O.OracleCfg(bf16=True)
    ref, n = O.video_tokens(video, p, cfg, objects)
    N = 196
    cols = torch.empty(B * F * N, 768, device="cuda", dtype=BF)
    ops.im2col_patches(video.cuda(), cols)
    # bit-exact bookkeeping: im2col must equal the reference unfold order
    ref_cols = video.reshape(B * F, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * F * N, 768).to(BF)
    assert torch.equal(cols.cpu(), ref_cols)
    w16 = torch.empty(768, 768, device="cuda", dtype=BF)
    ops.cast_bf16(p["video_model.patch_embed.proj.weight"].reshape(768, 768).cuda(), w16)
    patch = torch.empty(B * F * N, D, device="cuda")
    ops.gemm(cols, w16, bias=p["video_model.patch_embed.proj.bias"].cuda(), out_f32=patch)
    obj16 = torch.empty(B * F * O_, 2112, device="cuda", dtype=BF)
    ops.cast_bf16(objects.reshape(-1, 2054).cuda(), obj16)
    assert torch.equal(obj16[:, :2054].cpu(), objects.reshape(-1, 2054).to(BF)) and float(obj16[:, 2054:].abs().max()) == 0
    wo16 = torch.empty(768, 2112, device="cuda", dtype=BF)
    ops.cast_bf16(p["video_model.object_embed.weight"].cuda(), wo16)
    objemb = torch.empty(B * F * O_, D, device="cuda")
    ops.gemm(obj16, wo16, bias=p["video_model.object_embed.bias"].cuda(), out_f32=objemb)
    T = 1 + F * (N + O_)
    x = torch.empty(B * T, D, device="cuda")
    ops.assemble_tokens(patch, objemb, p["video_model.cls_token"].cuda(), p["video_model.pos_embed"].cuda(),
                        p["video_model.temporal_embed"].cuda(), None, x, B, F, N, O_, D)
    assert n == N + O_
    assert rel(x.cpu().view(B, T, D), ref) < 1e-3, rel(x.cpu().view(B, T, D), ref)
    # backward of the assembly
    dx = torch.randn(B * T, D, generator=g).cuda()
    dpatch = torch.empty(B * F * N, D, device="cuda", dtype=BF)
    dobj = torch.empty(B * F * O_, D, device="cuda", dtype=BF)
    dcls, dpos = torch.zeros(D, device="cuda"), torch.zeros(197 * D, device="cuda")
    dtem = torch.zeros(4 * D, device="cuda")
    ops.assemble_tokens_bwd(dx, dpatch, dobj, dcls, dpos, dtem, None, B, F, N, O_, D)
    d = dx.cpu().view(B, T, D)
    body = d[:, 1:].reshape(B, F, N + O_, D)
    assert torch.equal(dpatch.cpu(), body[:, :, :N].reshape(-1, D).to(BF))
    assert torch.equal(dobj.cpu(), body[:, :, N:].reshape(-1, D).to(BF))
    assert rel(dcls.cpu(), d[:, 0].sum(0)) < 1e-5
    ref_pos = torch.cat([d[:, 0].sum(0, keepdim=True), body[:, :, :N].sum((0, 1))], 0)
    assert rel(dpos.cpu().view(197, D), ref_pos) < 1e-5
    ref_tem = torch.zeros(4, D)
    ref_tem[:F] = body.sum((0, 2))
    assert rel(dtem.cpu().view(4, D), ref_tem) < 1e-5


def test_vecmat_and_strided_colsum_for_the_qkv_bias_identity():
    """oat_vecmat_f32 (db_v = db_proj . W_proj) and a column sum over the q slice of a [M, 3D] buffer (row pitch 3D)."""
    from oa_transformer_b200 import ops
    D, M = 768, 5000
    g = gen(41)
    v = torch.randn(D, generator=g).cuda()
    W = torch.randn(D, D, generator=g).cuda()
    out = torch.full((3 * D,), 2.0, device="cuda")
    ops.vecmat_f32(v, W, out[2 * D:])
    dqkv = torch.randn(M, 3 * D, generator=g).to(BF).cuda()
    ops.colsum_bf16(dqkv[:, :D], out[:D])
    torch.cuda.synchronize()
    assert rel(out[2 * D:] - 2.0, v.double() @ W.double()) < 1e-5
    assert rel(out[:D] - 2.0, dqkv[:, :D].double().sum(0)) < 1e-5
    assert bool((out[D:2 * D] == 2.0).all())


def test_colsum_and_text_embed():
    from oa_transformer_b200 import ops
    g = gen(9)
    x = torch.randn(1000, 2304, generator=g).to(BF).cuda()
    out = torch.ones(2304, device="cuda")
    ops.colsum_bf16(x, out)
    assert rel(out, x.float().sum(0) + 1) < 1e-5
    V, L, D, B = 500, 8, 768, 4
    word, pos = torch.randn(V, D, generator=g).cuda(), torch.randn(16, D, generator=g).cuda()
    ids = torch.randint(0, V, (B, L), generator=g).cuda()
    emb = torch.empty(B * L, D, device="cuda")
    ops.text_embed(ids, word, pos, emb, L)
    assert torch.equal(emb.view(B, L, D), word[ids] + pos[:L].unsqueeze(0))
    dsum = torch.randn(B * L, D, generator=g).cuda()
    dword, dpos = torch.zeros(V, D, device="cuda"), torch.zeros(16, D, device="cuda")
    ops.text_embed_bwd(ids, dsum, dword, dpos, L)
    rw = torch.zeros(V, D, device="cuda").index_add_(0, ids.view(-1), dsum)
    assert rel(dword, rw) < 1e-5
    assert rel(dpos[:L], dsum.view(B, L, D).sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ loss
def test_infonce_against_reference_golden():
    """sim_matrix + NormSoftmaxLoss vs the REFERENCE outputs in tests/golden/loss.pt (incl. the eps-clamped zero row)."""
    from oa_transformer_b200 import ops
    gold = torch.load(os.path.join(GOLD, "loss.pt"), map_location="cpu", weights_only=False)
    for name, c in gold.items():
        loss, sims, dt, dv = ops.infonce_fwd_bwd(c["a"].cuda().contiguous(), c["b"].cuda().contiguous(), want_sims=True)
        assert float((sims.cpu() - c["sims"]).abs().max()) < 1e-3, name     # north-star gate on logits
        assert float((sims.cpu() - c["sims"]).abs().max()) < 2e-5, name     # what the split-bf16 GEMM actually gives
        assert abs(float(loss) - float(c["loss"])) < 1e-4 * max(1.0, abs(float(c["loss"]))), name
        assert rel(dt.cpu(), c["ga"]) < 1e-3 and rel(dv.cpu(), c["gb"]) < 1e-3, (name, rel(dt.cpu(), c["ga"]))


# ------------------------------------------------------------------------------------------------ object -> patch (X4)
def test_patch_masks_bit_exact_vs_reference_fixture():
    from oa_transformer_b200 import ops
    g = torch.load(os.path.join(GOLD, "patch_masks.pt"), map_location="cpu", weights_only=False)
    masks = ops.patch_masks_from_bbox(g["boxes"].cuda().contiguous())
    assert torch.equal(masks.cpu().double(), g["masks"])


@pytest.mark.parametrize("mode,C,Ob", [("softmax", 256, 5), ("sigmoid", 256, 5), ("softmax", 768, 36), ("mask", 768, 20)])
def test_object_patch_attention_modes(mode, C, Ob):
    from oa_transformer_b200 import ops
    g = gen(11)
    B, L = 3, 196
    q = torch.randn(B, Ob, C, generator=g) * 0.3
    k = torch.randn(B, L, C, generator=g) * 0.3
    v = torch.randn(B, L, C, generator=g)
    masks = None
    if mode == "mask":
        boxes = O.synth_objects(B, 1, Ob, g)[:, 0, :, 2048:2052].reshape(-1, 4).double()
        masks = torch.from_numpy(O.patch_masks_from_bbox(boxes.numpy())).float().view(B, Ob, L)
    w_ref, o_ref = O.object_patch_attention(q, k, v, mode, masks)
    w, o = ops.object_patch_attention(None if mode == "mask" else q.cuda(), None if mode == "mask" else k.cuda(),
                                      v.cuda(), mode, None if masks is None else masks.cuda().contiguous())
    assert rel(w.cpu(), w_ref) < 1e-5
    assert rel(o.cpu(), o_ref) < 1e-5


def test_bookkeeping_kernels_bit_exact_vs_reference_fixture():
    """SURVEY.md 8f-2 on the device, against OUTPUTS OF THE REFERENCE (tests/golden/bookkeeping.pt): same-class union
    patch masks (base_dataset_region_mem.py:233-247), tag-token end offsets (base_dataset_global_local.py:395-405), region
    features in confidence order with class de-duplication and numpy 'edge' padding (base_dataset.py:593-650)."""
    from oa_transformer_b200 import ops
    g = torch.load(os.path.join(GOLD, "bookkeeping.pt"), map_location="cpu", weights_only=False)
    for c in g["region_mem_masks"]:
        masks = ops.patch_masks_same_class(c["boxes"].cuda(), c["classes"].cuda(), c["indexs"].cuda())
        assert torch.equal(masks.cpu().double(), c["masks"])
    lens = g["object_tags"]["lens"].cuda()
    for c in g["object_tags"]["cases"]:
        ends, total = ops.object_tags_masks(lens, c["indices"].cuda())
        assert torch.equal(ends.cpu(), c["mask"]) and total == c["total"]
    for c in g["region_features"]:
        feat = ops.region_features_topk(c["x"].cuda(), c["bbox"].cuda(), c["conf"].cuda(), c["ids"].cuda(),
                                        c["image_w"], c["image_h"], c["top_k"], c["v"])
        assert feat.shape == c["feat"].shape
        assert torch.equal(feat.cpu(), c["feat"]), (c["top_k"], c["v"], float((feat.cpu() - c["feat"]).abs().max()))


@pytest.mark.parametrize("mode,C,Ob,L,with_v", [("sigmoid", 256, 5, 196, False), ("sigmoid", 256, 5, 196, True),
                                                   ("softmax", 768, 36, 196, True), ("softmax", 256, 20, 50, False),
                                                   ("mask", 0, 20, 196, True)])
def test_object_patch_attention_backward(mode, C, Ob, L, with_v):
    """X4 on the fwd/bwd path: gradients of the three score -> weight modes w.r.t. q, k, v through the autograd wrapper
    (functional.object_patch_attention) vs torch autograd of the oracle (fp32 both sides)."""
    from oa_transformer_b200.functional import object_patch_attention
    g = gen(31)
    B, Cv = 3, 128
    q = (0.2 * torch.randn(B, Ob, C, generator=g)) if C else None
    k = (0.2 * torch.randn(B, L, C, generator=g)) if C else None
    v = torch.randn(B, L, Cv, generator=g) if with_v else None
    masks = (torch.rand(B, Ob, L, generator=g) > 0.7).float() if mode == "mask" else None
    pw, po = torch.randn(B, Ob, L, generator=g), torch.randn(B, Ob, Cv, generator=g)

    def run(fn, dev):
        t = [None if x is None else x.to(dev).clone().requires_grad_(True) for x in (q, k, v)]
        w, o = fn(t[0], t[1], t[2], mode=mode, masks=None if masks is None else masks.to(dev))
        loss = (w * pw.to(dev)).sum() if mode != "mask" else 0.0
        if o is not None:
            loss = loss + (o * po.to(dev)).sum()
        loss.backward()
        return w.detach().cpu(), None if o is None else o.detach().cpu(), [None if x is None else x.grad.cpu() for x in t]

    w_ref, o_ref, g_ref = run(O.object_patch_attention, "cpu")
    w_out, o_out, g_out = run(object_patch_attention, "cuda")
    assert rel(w_out, w_ref) < 1e-5 and (o_ref is None or rel(o_out, o_ref) < 1e-5)
    for a, b, name in zip(g_out, g_ref, "qkv"):
        assert (a is None) == (b is None), name
        if b is not None:
            assert rel(a, b) < 2e-5, (name, rel(a, b))


def test_token_pool_bce_and_linear_heads_vs_torch():
    """The small head ops of the region / global-local variants vs torch: (cls + mean(tokens)) / 2 on a [:, 1:] slice
    of strided clips, scale * BCELoss(sum) with its gradient, and the differentiable bf16 linear (ReLU -> Linear)."""
    from oa_transformer_b200 import functional as OF
    g = gen(33)
    B2, T, P = 6, 9, 256
    tok = torch.randn(B2, T, P, generator=g)
    cls = torch.randn(B2, P, generator=g)
    probe = torch.randn(B2 // 2, P, generator=g)
    tr, cr = tok.clone().requires_grad_(True), cls.clone().requires_grad_(True)
    ref = (cr[1::2] + tr[1::2, 1:].mean(dim=1)) / 2
    (ref * probe).sum().backward()
    tc, cc = tok.cuda().requires_grad_(True), cls.cuda().requires_grad_(True)
    out = OF.token_pool(cc[1::2], tc[1::2, 1:], 0.5, 0.5)
    (out * probe.cuda()).sum().backward()
    assert rel(out.detach().cpu(), ref.detach()) < 1e-6
    assert rel(tc.grad.cpu(), tr.grad) < 1e-6 and rel(cc.grad.cpu(), cr.grad) < 1e-6
    # BCE(sum): probabilities incl. saturated ones (torch clamps the logs at -100 and the gradient denominator at 1e-12)
    p = torch.rand(40, 196, generator=g)
    p[0, :4] = torch.tensor([0.0, 1.0, 1e-30, 1 - 1e-7])
    t = (torch.rand(40, 196, generator=g) > 0.5).float()
    pr = p.clone().requires_grad_(True)
    lref = 0.1 * torch.nn.BCELoss(reduction="sum")(pr, t) / 40
    lref.backward()
    pc = p.cuda().requires_grad_(True)
    lout = OF.bce_sum(pc, t.cuda(), 0.1 / 40)
    lout.backward()
    assert abs(float(lout) - float(lref)) < 1e-5 * abs(float(lref))
    assert torch.allclose(pc.grad.cpu(), pr.grad, rtol=1e-5, atol=1e-9)
    # ReLU -> Linear(512, 256) (txt_proj_2) vs the bf16-operand oracle
    x = torch.randn(20, 512, generator=g)
    w = 0.05 * torch.randn(256, 512, generator=g)
    b = 0.1 * torch.randn(256, generator=g)
    probe = torch.randn(20, 256, generator=g)
    xr, wr, br = (z.clone().requires_grad_(True) for z in (x, w, b))
    yref = O.linear(torch.relu(xr), wr, br, O.OracleCfg(bf16=True))
    (yref * probe).sum().backward()
    xc, wc, bc = (z.cuda().requires_grad_(True) for z in (x, w, b))
    y = OF.linear(xc, wc, bc, relu=True)
    (y * probe.cuda()).sum().backward()
    assert rel(y.detach().cpu(), yref.detach()) < 1e-5
    for a, r_, name in ((xc, xr, "x"), (wc, wr, "w"), (bc, br, "b")):
        assert rel(a.grad.cpu(), r_.grad) < 6e-3, (name, rel(a.grad.cpu(), r_.grad))


def test_dropout_draws_statistics_and_elementwise_kernels():
    """Philox dropout (include/oat.h): keep rate and independence of the draws over sites / seeds, and the forward /
    backward element kernels against the exported mask (exact)."""
    from oa_transformer_b200 import ops
    n, p = 1 << 20, 0.1
    keeps = {}
    for seed in (1, 2, 987654321012345):
        for site in (0, 1, 7):
            k = ops.dropout_mask(n, p, seed, site, "cuda").float()
            keeps[(seed, site)] = k
            rate = float(k.mean())
            assert abs(rate - (1 - p)) < 5 * (p * (1 - p) / n) ** 0.5, (seed, site, rate)       # 5 sigma
            # no structure along the index: lag-1 / lag-4 / lag-768 autocorrelation of the keep flags ~ 0
            c = k - k.mean()
            for lag in (1, 4, 768):
                assert abs(float((c[:-lag] * c[lag:]).mean()) / float(c.var())) < 6e-3, (seed, site, lag)
    a, b, c = keeps[(1, 0)], keeps[(2, 0)], keeps[(1, 1)]
    for x, y in ((a, b), (a, c)):       # different seed / different site: independent masks
        agree = float((x == y).float().mean())
        assert abs(agree - ((1 - p) ** 2 + p ** 2)) < 3e-3, agree
    assert torch.equal(ops.dropout_mask(n, p, 1, 0, "cuda").float(), a)         # same triple, same draws
    rows, cols, seed, site = 37, 768, 42, 5
    g = gen(41)
    x = torch.randn(rows, cols, generator=g).cuda()
    res = torch.randn(rows, cols, generator=g).cuda()
    m = ops.dropout_mask(rows * cols, p, seed, site, "cuda").view(rows, cols).float() / (1 - p)
    out = torch.empty_like(x)
    out16 = torch.empty(rows, cols, device="cuda", dtype=BF)
    out3 = torch.empty(rows, 3 * cols, device="cuda", dtype=BF)
    ops.dropout_fwd(x, p, seed, site, residual=res, out=out, out_bf16=out16, out_split3=out3)
    ref = x * m + res
    assert torch.equal(out, ref) and torch.equal(out16, ref.to(BF))
    hi, lo = _split_ref(ref)
    assert torch.equal(out3[:, :cols], hi) and torch.equal(out3[:, cols:2 * cols], hi) and torch.equal(out3[:, 2 * cols:], lo)
    dy32 = torch.randn(rows, cols, generator=g).cuda()
    dy16 = torch.randn(rows, cols, generator=g).to(BF).cuda()
    dx32 = torch.empty_like(x)
    dx16 = torch.empty(rows, cols, device="cuda", dtype=BF)
    ops.dropout_bwd(p, seed, site, dy_f32=dy32, dy_bf16=dy16, dx_f32=dx32, dx_bf16=dx16)
    assert torch.equal(dx32, (dy32 + dy16.float()) * m) and torch.equal(dx16, ((dy32 + dy16.float()) * m).to(BF))


def test_text_attention_weight_dropout_vs_oracle_with_same_mask():
    """mode 2 attention with dropout on the softmax weights (HF DistilBERT attention_dropout): forward and backward against
    the oracle given the multiplier built from the exported draws of the same (seed, site)."""
    from oa_transformer_b200 import ops
    B, L, H, p, seed, site = 3, 32, 4, 0.1, 77, 4
    T = L
    qkv16 = _qkv(B, T, H, 5)
    dout16 = torch.randn(B, T, H * 64, generator=gen(6)).to(BF)
    key_mask = torch.ones(B, L, dtype=torch.long)
    key_mask[1, 20:] = 0
    qkv = qkv16.reshape(B * T, 3 * H * 64).cuda()
    out = torch.zeros(B * T, H * 64, device="cuda", dtype=BF)
    lse = torch.zeros(B * H * T, device="cuda")
    km = key_mask.to(torch.int32).cuda().contiguous()
    ops.attn_fwd(ops.MODE_PLAIN, B, T, H, 0, 0, qkv, out, lse, km, dropout=(p, seed, site))
    dqkv = torch.zeros_like(qkv)
    ops.attn_bwd(ops.MODE_PLAIN, B, T, H, 0, 0, qkv, out, lse, dout16.reshape(B * T, H * 64).cuda(), dqkv, 0.125, None, km,
                 dropout=(p, seed, site))
    mult = ops.dropout_mask(B * H * T * T, p, seed, site, "cuda").view(B, H, T, T).float().cpu() / (1 - p)
    x = qkv16.float().view(B, T, 3, H, 64).clone().requires_grad_(True)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    add = torch.zeros(B, 1, 1, T).masked_fill(key_mask.view(B, 1, 1, T) == 0, float("-inf"))
    ref = O._softmax_attention(q, k, v, O.OracleCfg(heads=H, bf16=True), add, mult).permute(0, 2, 1, 3).reshape(B, T, H * 64)
    ref.backward(dout16.float())
    gref = x.grad.clone()
    gref[:, :, 0] *= 0.125
    o = out.cpu().float().view(B, T, H * 64)
    valid = key_mask.bool()
    assert rel(o[valid], ref.detach()[valid]) < 6e-3
    dq = dqkv.cpu().float().view(B, T, 3, H, 64)
    assert rel(dq[valid], gref[valid]) < 1.2e-2, rel(dq[valid], gref[valid])
    # and it differs from the no-dropout result by the expected amount (sqrt(p / (1 - p)) relative noise on the weights)
    out0 = torch.zeros_like(out)
    ops.attn_fwd(ops.MODE_PLAIN, B, T, H, 0, 0, qkv, out0, lse, km)
    assert rel(o[valid], out0.cpu().float().view(B, T, H * 64)[valid]) > 2e-2


def test_device_prefetcher_stages_batches_in_order():
    from oa_transformer_b200.data_loader import DevicePrefetcher
    host = [{"video": torch.full((2, 3, 8), float(i)).pin_memory(),
             "text": {"input_ids": torch.full((2, 4), i, dtype=torch.int64).pin_memory()}, "meta": {"i": i}}
            for i in range(5)]
    seen = []
    for data in DevicePrefetcher(host, torch.device("cuda", 0)):
        assert data["video"].is_cuda and data["text"]["input_ids"].is_cuda
        seen.append((float(data["video"].sum()) / 48.0, int(data["text"]["input_ids"][0, 0]), data["meta"]["i"]))
    assert seen == [(float(i), i, i) for i in range(5)]


@pytest.mark.parametrize("n,ties", [(1000, False), (257, True), (7, True)])
def test_retrieval_metrics_device_matches_host(n, ties):
    """oat_retrieval_ranks + cols2metrics vs the numpy port (itself pinned to the reference fixture), incl. tied scores."""
    from oa_transformer_b200.model import metric as M
    g = gen(31)
    sims = torch.randn(n, n, generator=g)
    sims += 2.0 * torch.eye(n) * (torch.rand(n, generator=g) > 0.3).float().diag()   # some pairs retrieved, some not
    if ties:
        sims = (sims * 2).round() / 2            # coarse grid: many exact ties, also with the diagonal
        sims[0] = 0.0                            # a constant row: the optimistic rule gives rank 0, averaging (n-1)/2
    dev = sims.cuda()
    for fn in (M.t2v_metrics, M.v2t_metrics):
        host = fn(sims.numpy())
        devm = fn(dev)
        assert host.keys() == devm.keys()
        for k in host:
            assert host[k] == devm[k], (fn.__name__, k, host[k], devm[k])


@pytest.mark.parametrize("correct_bias,wd", [(True, 0.0), (False, 0.01), (True, 0.05)])
def test_fused_adamw_matches_transformers_semantics(correct_bias, wd):
    """oat_adamw_multi vs a plain-torch restatement of transformers.AdamW.step (optimization.py, 4.6), several steps,
    odd sizes and an unaligned view (scalar tail path)."""
    from oa_transformer_b200.optim import AdamW
    g = gen(41)
    flat = torch.randn(5000, generator=g).cuda()
    shapes = [(768, 33), (1,), (1027,), (256, 768)]
    params = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    params.append(torch.nn.Parameter(flat[1:1 + 999]))            # 4-byte aligned only
    ref = [p.detach().clone() for p in params]
    m = [torch.zeros_like(p) for p in ref]
    v = [torch.zeros_like(p) for p in ref]
    lr, b1, b2, eps = 3e-3, 0.9, 0.999, 1e-6
    opt = AdamW(params, lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd, correct_bias=correct_bias)
    for step in range(1, 4):
        grads = [torch.randn(p.shape, generator=g).cuda() * 0.1 for p in params]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        opt.step()
        for i, gr in enumerate(grads):
            m[i].mul_(b1).add_(gr, alpha=1 - b1)
            v[i].mul_(b2).addcmul_(gr, gr, value=1 - b2)
            step_size = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step) if correct_bias else lr
            ref[i].addcdiv_(m[i], v[i].sqrt().add_(eps), value=-step_size)
            if wd > 0:
                ref[i].add_(ref[i], alpha=-lr * wd)
    torch.cuda.synchronize()
    for p, r in zip(params, ref):
        assert torch.allclose(p.detach(), r, rtol=2e-6, atol=1e-6), (p.shape, (p.detach() - r).abs().max())
    sd = opt.state_dict()
    assert sd["state"][0]["step"] == 3 and sd["state"][0]["exp_avg"].shape == (768, 33)
