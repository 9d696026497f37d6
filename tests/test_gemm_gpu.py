"""tcgen05 GEMM vs torch fp32 matmul on the same bf16-rounded operands (tolerance: fp32 accumulation order only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).to(torch.bfloat16).cuda()


def _ref(A, B, a_major, b_major):
    a = A.float() if a_major == 0 else A.float().t()
    b = B.float() if b_major == 0 else B.float().t()
    return a @ b.t()


@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 768, 768), (1857, 2304, 768), (256, 128, 200), (64, 256, 2112)])
def test_gemm_plain(M, N, K, a_major, b_major):
    from oa_transformer_b200 import ops
    A = _mk((M, K) if a_major == 0 else (K, M + (-M) % 8), 1)
    B = _mk((N, K) if b_major == 0 else (K, N), 2)
    if a_major == 1:
        A = A[:, :M]
    if K % 8 != 0 and (a_major == 0 or b_major == 0):
        pytest.skip("K-major operands need a 16-byte row pitch")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(A, B, a_major=a_major, b_major=b_major, out_f32=out)
    torch.cuda.synchronize()
    ref = _ref(A, B, a_major, b_major)
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, "max abs err %g (scale %g)" % (err, scale)


def test_gemm_epilogue_bias_scale_residual_bf16():
    from oa_transformer_b200 import ops
    M, N, K = 500, 2304, 768
    A, B = _mk((M, K), 3), _mk((N, K), 4)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, scale_cols=768, scale=0.125, residual=res, out_f32=o32, out_bf16=o16)
    ref = A.float() @ B.float().t() + bias
    ref[:, :768] *= 0.125
    ref = ref + res
    assert torch.allclose(o32, ref, rtol=1e-3, atol=2e-3)
    assert torch.allclose(o16.float(), ref, rtol=1e-2, atol=2e-2)


def test_gemm_gelu_forward_backward():
    from oa_transformer_b200 import ops
    M, N, K = 300, 3072, 768
    A, B = _mk((M, K), 5), _mk((N, K), 6)
    B = (B.float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    dact = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    act = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, act=ops.ACT_GELU, out_bf16=act, out2_bf16=dact)
    u = (A.float() @ B.float().t() + bias).requires_grad_(True)
    g = torch.nn.functional.gelu(u)
    g.sum().backward()
    assert torch.allclose(act.float(), g.detach(), rtol=1e-2, atol=1e-2)
    assert torch.allclose(dact.float(), u.grad, rtol=1e-2, atol=1e-2)      # out2 holds GELU'(u)
    # backward: dU = (dG @ W) * GELU'(u)   with W = [N_out, K_in] stored row-major -> MN-major B operand
    dG = _mk((M, 768), 7)
    W2 = (_mk((768, N), 8).float() * 0.05).to(torch.bfloat16)  # fc2.weight [768, 3072]
    dU = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dG, W2, b_major=1, act=ops.ACT_GELU_BWD, aux=dact, out_bf16=dU)
    ref = (dG.float() @ W2.float()) * dact.float()
    assert torch.allclose(dU.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("M", [100, 300, 1857 * 2])
def test_gemm_rowdot_epilogue(M):
    """act 4: the bf16 output is the plain product and rowdot[h, r] = sum over head h's 64 columns of bf16(out) * aux -
    delta = rowsum(dO * O) of the attention backward from the epilogue of the dO-producing GEMM."""
    from oa_transformer_b200 import ops
    N, K = 768, 768
    dY = _mk((M, K), 21)
    W = (_mk((K, N), 22).float() * 0.05).to(torch.bfloat16)          # proj.weight [out = K, in = N]: MN-major B operand
    O_ = _mk((M, N), 23)
    plain = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dY, W, b_major=1, out_bf16=plain)
    out = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    ld = M + 5
    rowdot = torch.full((N // 64, ld), -3.0, device="cuda")
    ops.gemm(dY, W, b_major=1, act=ops.ACT_ROWDOT, aux=O_, rowdot=rowdot, out_bf16=out)
    torch.cuda.synchronize()
    assert torch.equal(out, plain)                                   # the output itself is untouched by the extra work
    ref = (out.float() * O_.float()).view(M, N // 64, 64).sum(-1).t()
    assert torch.allclose(rowdot[:, :M], ref, rtol=1e-5, atol=1e-4), (rowdot[:, :M] - ref).abs().max().item()
    assert bool((rowdot[:, M:] == -3.0).all())                       # nothing written past row M
    with pytest.raises(Exception):                                   # needs 256-column granularity
        ops.gemm(dY, W[:, :192].contiguous(), b_major=1, act=ops.ACT_ROWDOT, aux=O_[:, :192].contiguous(),
                 rowdot=rowdot[:3], out_bf16=out[:, :192].contiguous())


def test_gemm_wgrad_splitk_accumulate():
    from oa_transformer_b200 import ops
    T, Nout, Kin = 5000, 768, 768
    dY, X = _mk((T, Nout), 9), _mk((T, Kin), 10)
    dW = torch.ones(Nout, Kin, device="cuda")
    ops.gemm(dY, X, a_major=1, b_major=1, out_f32=dW, accumulate=True)
    ref = dY.float().t() @ X.float() + 1.0
    err = (dW - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-2, err
    # forced split counts
    for s in (1, 3, 7):
        dW.zero_()
        ops.gemm(dY, X, a_major=1, b_major=1, out_f32=dW, accumulate=True, split_k=s)
        err = (dW - (ref - 1.0)).abs().max().item()
        assert err <= 2e-3 * ref.abs().max().item() + 1e-2, (s, err)


def test_gemm_many_tiles_persistent():
    from oa_transformer_b200 import ops
    M, N, K = 128 * 150 + 17, 768, 256  # more tiles than SMs, TMEM double-buffer phases wrap several times
    A, B = _mk((M, K), 11), _mk((N, K), 12)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, out_bf16=out)
    ref = A.float() @ B.float().t()
    assert torch.allclose(out.float(), ref, rtol=1e-2, atol=0.15)


@pytest.mark.parametrize("M", [500, 128 * 149 + 33])
def test_gemm_tma_epilogue_f32_bias_residual(M):
    """fp32 output = acc + bias + fp32 residual through the TMA-box epilogue (ragged last row block, > 148 tiles)."""
    from oa_transformer_b200 import ops
    N, K = 768, 768
    A, B = _mk((M, K), 21), _mk((N, K), 22)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    out = torch.full((M + 3, N), 7.0, device="cuda")            # rows beyond M must stay untouched
    ops.gemm(A, B, bias=bias, residual=res, out_f32=out[:M])
    ref = A.float() @ B.float().t() + bias + res
    assert torch.allclose(out[:M], ref, rtol=1e-3, atol=2e-3)
    assert bool((out[M:] == 7.0).all())
    out2 = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, bias=bias, out_f32=out2)                      # no residual
    assert torch.allclose(out2, ref - res, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("scale_cols", [768, 68, 0])
def test_gemm_tma_epilogue_bf16_bias_qscale(scale_cols):
    from oa_transformer_b200 import ops
    M, N, K = 1857, 2304, 768
    A, B = _mk((M, K), 23), _mk((N, K), 24)
    B = (B.float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.full((M + 2, N), 3.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, scale_cols=scale_cols, scale=0.125, out_bf16=out[:M])
    ref = A.float() @ B.float().t() + bias
    ref[:, :scale_cols] *= 0.125
    assert torch.allclose(out[:M].float(), ref, rtol=1e-2, atol=2e-2)
    assert bool((out[M:].float() == 3.0).all())


@pytest.mark.parametrize("pairs", ["1", "0"])
def test_gemm_cta_pairs_and_single_cta_agree(pairs, monkeypatch):
    """Every specialised epilogue through CTA pairs (cta_group::2, default) and through single CTAs (OAT_GEMM_2CTA=0)."""
    from oa_transformer_b200 import ops
    monkeypatch.setenv("OAT_GEMM_2CTA", pairs)
    M, N, K = 128 * 75 + 40, 768, 768            # odd number of 128-row blocks: the last pair has an empty second CTA
    A, B = _mk((M, K), 31), _mk((N, K), 32)
    B = (B.float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    res = torch.randn(M, N, device="cuda")
    ref = A.float() @ B.float().t()
    o32 = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, bias=bias, residual=res, out_f32=o32)
    assert torch.allclose(o32, ref + bias + res, rtol=1e-3, atol=2e-3)
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, scale_cols=256, scale=0.125, out_bf16=o16)
    r2 = ref + bias
    r2[:, :256] *= 0.125
    assert torch.allclose(o16.float(), r2, rtol=1e-2, atol=2e-2)
    g16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    d16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, act=ops.ACT_GELU, out_bf16=g16, out2_bf16=d16)
    assert torch.allclose(g16.float(), torch.nn.functional.gelu(ref + bias), rtol=1e-2, atol=1e-2)
    m16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, b_major=0, act=ops.ACT_GELU_BWD, aux=d16, out_bf16=m16)
    assert torch.allclose(m16.float(), ref * d16.float(), rtol=2e-2, atol=2e-2)
    # MN-major operands + split-K reduce-add (weight-gradient shape): dW[N, K] += dY[M, N]^T X[M, K]
    dW = torch.ones(N, K, device="cuda")
    ops.gemm(o16, A, a_major=1, b_major=1, out_f32=dW, accumulate=True)
    refw = o16.float().t() @ A.float() + 1.0
    assert (dW - refw).abs().max().item() <= 2e-3 * refw.abs().max().item() + 2e-2
    # dgrad shape: B MN-major
    dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    ops.gemm(o16, B, b_major=1, out_bf16=dx)
    assert torch.allclose(dx.float(), o16.float() @ B.float(), rtol=2e-2, atol=5e-2)


@pytest.mark.parametrize("M,N,K", [(32, 2304, 2304), (32, 768, 6144), (4, 256, 768), (64, 3072, 2304), (1, 768, 768),
                                   (17, 260, 136)])
def test_skinny_gemm_epilogues_strided_and_reproducible(M, N, K, monkeypatch):
    """gemm_skinny.cu (<= 64 rows, K-major x K-major: the split-bf16 CLS rows and the projections): every epilogue of
    oat_gemm_bf16 on strided rows (the CLS rows of a token buffer), the accumulate form, and bit-for-bit reproducibility
    (one owner per output element, fixed k order: no atomics)."""
    from oa_transformer_b200 import ops
    T = 5
    A_all = _mk((M * T, K), 7)
    A = A_all.view(M, T * K)[:, :K]                       # rows 0, T, 2T, ... of a token buffer
    B = (_mk((N, K), 8).float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    exact = A.float() @ B.float().t()
    scale = exact.abs().max().item()
    # plain fp32 + bf16 outputs on strided rows, with bias and the q column scale
    big32 = torch.zeros(M * T, N, device="cuda")
    big16 = torch.zeros(M * T, N, device="cuda", dtype=torch.bfloat16)
    sc = (N // 3) // 4 * 4
    ops.gemm(A, B, bias=bias, scale_cols=sc, scale=0.125, out_f32=big32.view(M, T * N)[:, :N],
             out_bf16=big16.view(M, T * N)[:, :N])
    ref = exact + bias
    ref[:, :sc] *= 0.125
    assert (big32[::T] - ref).abs().max().item() <= 2e-3 * scale + 1e-3
    assert torch.allclose(big16[::T].float(), ref, rtol=1e-2, atol=1e-2 * scale)
    assert float(big32[1::T].abs().max()) == 0.0 and float(big16[1::T].float().abs().max()) == 0.0
    # reproducible bit for bit
    again = torch.zeros_like(big32)
    ops.gemm(A, B, bias=bias, scale_cols=sc, scale=0.125, out_f32=again.view(M, T * N)[:, :N])
    assert torch.equal(again, big32)
    # residual, then the accumulate form on top (the accumulate form itself runs split-K on the tcgen05 kernel)
    res = torch.randn(M, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, residual=res, out_f32=out)
    ops.gemm(A, B, out_f32=out, accumulate=True)
    assert (out - (2 * exact + res)).abs().max().item() <= 4e-3 * scale + 2e-3
    # GELU with its stored derivative, fp32 + bf16 outputs at once (the fc1 CLS rows)
    g32 = torch.empty(M, N, device="cuda")
    g16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    d16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, act=ops.ACT_GELU, out_f32=g32, out_bf16=g16, out2_bf16=d16)
    u = (exact + bias).double()
    cdf = 0.5 * (1 + torch.erf(u / 2 ** 0.5))
    gref = (u * cdf).float()
    dref = (cdf + u * torch.exp(-0.5 * u * u) / (2 * 3.141592653589793) ** 0.5).float()
    assert (g32 - gref).abs().max().item() <= 3e-3 * scale + 1e-3
    assert torch.allclose(d16.float(), dref, rtol=2e-2, atol=2e-2)
    # x aux (GELU backward) and ReLU
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, act=ops.ACT_GELU_BWD, aux=d16, out_bf16=o)
    assert torch.allclose(o.float(), exact * d16.float(), rtol=2e-2, atol=2e-2 * scale)
    r = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, bias=bias, act=ops.ACT_RELU, out_f32=r)
    assert (r - (exact + bias).clamp_min(0)).abs().max().item() <= 2e-3 * scale + 1e-3
    # and it agrees with the tcgen05 kernel on the same problem
    monkeypatch.setenv("OAT_GEMM_SKINNY", "0")


@pytest.mark.parametrize("pairs", ["1", "0"])
@pytest.mark.parametrize("M,N,K", [(2304, 776, 5000), (3072, 776, 1857 * 4), (768, 264, 3000), (2304, 784, 4000)])
def test_gemm_wgrad_with_ones_column_gives_bias_gradient(M, N, K, pairs, monkeypatch):
    """Weight + bias gradient in one GEMM: dY^T . [X | 1 0 .. 0] with a ragged, 16-wide last column block (multiplied by
    an N = 16 instruction, gemm_tcgen05.cu) accumulated into an fp32 scratch, then oat_unpack_wgrad. vs torch."""
    from oa_transformer_b200 import ops
    monkeypatch.setenv("OAT_GEMM_2CTA", pairs)
    kin = (N // 256) * 256
    dy = _mk((K, M), 11)                                   # [tokens, N_out]
    xe = torch.zeros(K, N, device="cuda", dtype=torch.bfloat16)
    xe[:, :kin] = _mk((K, kin), 12)
    xe[:, kin] = 1.0
    scratch = torch.zeros(M, N, device="cuda")
    ops.gemm(dy, xe, a_major=1, b_major=1, out_f32=scratch, accumulate=True)
    ref = dy.float().t() @ xe.float()
    scale = ref[:, :kin].abs().max().item()
    assert (scratch[:, :kin] - ref[:, :kin]).abs().max().item() <= 2e-3 * scale + 1e-2
    bsum = dy.float().sum(0)
    assert (scratch[:, kin] - bsum).abs().max().item() <= 2e-3 * bsum.abs().max().item() + 1e-2
    dw = torch.ones(M, kin, device="cuda")
    db = torch.ones(M, device="cuda")
    ops.unpack_wgrad(scratch, kin, dw, db)
    assert (dw - 1 - ref[:, :kin]).abs().max().item() <= 2e-3 * scale + 1e-2
    assert (db - 1 - bsum).abs().max().item() <= 2e-3 * bsum.abs().max().item() + 1e-2
    assert float(scratch.abs().max()) == 0.0                # handed back zeroed for the next accumulate
