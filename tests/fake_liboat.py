"""TEST INFRASTRUCTURE ONLY - a plain-torch (CPU) stand-in for `oa_transformer_b200.ops` so that the HOST-SIDE schedules
(engine.VideoEngine / engine.TextEngine: buffer management, residual wiring, gradient bookkeeping, operand packing) can
be exercised by the `-m "not gpu"` suite against the oracle without a GPU. It follows the operator contracts of
include/oat.h (epilogue order of oat_gemm_bf16, accumulate semantics of the LayerNorm / column-sum reductions, token
indexing of oat_assemble_tokens, ...). Nothing in the product imports this file; the product path has no CPU
implementation and fails loudly without liboat.so / a B200 (tests/test_cabi_cpu.py::test_no_cpu_fallback)."""
import torch

from oracle import oracle as O

ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_RELU, ACT_ROWDOT = 0, 1, 2, 3, 4
MODE_SPACE, MODE_TIME, MODE_PLAIN = 0, 1, 2
LAUNCHES = 0
GRAPH_REPLAYS = 0
PROFILE = None
BF = torch.bfloat16


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _rows(t, rows, pitch, cols):
    """rows x cols view of a contiguous buffer read with row pitch `pitch` (elements)."""
    flat = t.reshape(-1)
    return torch.as_strided(flat, (rows, cols), (pitch, 1))


class CastPlan:
    def __init__(self):
        self.items = []

    def add(self, src, dst, rows=None, cols=None, split=False):
        src2 = src.detach()
        src2 = src2.reshape(src2.shape[0], -1) if src2.dim() != 2 else src2
        self.items.append((src2, dst, src2.shape[0] if rows is None else rows, src2.shape[1] if cols is None else cols,
                           split))

    def run(self):
        for s, d, r, c, split in self.items:
            d[:r].zero_()
            if split:                       # weight layout [hi | lo | hi], each segment dst.shape[1] // 3 wide
                seg = d.shape[1] // 3
                hi = s[:r, :c].to(BF)
                lo = (s[:r, :c] - hi.float()).to(BF)
                d[:r, :c], d[:r, seg:seg + c], d[:r, 2 * seg:2 * seg + c] = hi, lo, hi
            else:
                d[:r, :c] = s[:r, :c].to(d.dtype)
        self.items = []
        _count(1)


def split3_bf16(src, dst, *, rows=None, cols=None, lds=None, relu=False):
    """activation layout [hi | hi | lo]"""
    rows = dst.shape[0] if rows is None else rows
    cols = src.shape[-1] if cols is None else cols
    lds = (src.stride(0) if src.dim() == 2 else cols) if lds is None else lds
    v = _rows(src.detach(), rows, lds, cols).float()
    if relu:
        v = v.clamp_min(0)
    hi = v.to(BF)
    dst[:rows, :cols], dst[:rows, cols:2 * cols], dst[:rows, 2 * cols:3 * cols] = hi, hi, (v - hi.float()).to(BF)
    _count(1)


def cast_bf16(src, dst, *, rows=None, cols=None, lds=None, relu=False):
    rows = dst.shape[0] if rows is None else rows
    cols = src.shape[-1] if cols is None else cols
    lds = (src.stride(0) if src.dim() == 2 else cols) if lds is None else lds
    v = _rows(src.detach(), rows, lds, cols).float()
    if relu:
        v = v.clamp_min(0)
    dst[:rows].zero_()
    dst[:rows, :cols] = v.to(BF)
    _count(1)


def relu_bwd(x, dy_bf16, dx, *, rows, cols, ldx, lddx=None):
    lddx = dx.stride(0) if lddx is None else lddx
    mask = _rows(x, rows, ldx, cols) > 0
    _rows(dx, rows, lddx, cols).copy_(torch.where(mask, dy_bf16[:rows, :cols].float(), torch.zeros(())))
    _count(1)


def gemm(A, B, *, a_major=0, b_major=0, alpha=1.0, bias=None, scale_cols=0, scale=1.0, act=ACT_NONE, aux=None,
         residual=None, out_f32=None, out_bf16=None, out2_bf16=None, accumulate=False, split_k=0, rowdot=None):
    assert A.dtype == BF and B.dtype == BF
    a = A.float() if a_major == 0 else A.float().t()
    b = B.float() if b_major == 0 else B.float().t()
    c = alpha * (a @ b.t())
    if bias is not None:
        c = c + bias
    if scale_cols:
        c[:, :scale_cols] = c[:, :scale_cols] * scale
    if act == ACT_GELU:
        cdf = 0.5 * (1.0 + torch.erf(c * 0.7071067811865476))
        pdf = 0.3989422804014327 * torch.exp(-0.5 * c * c)
        out2_bf16.copy_((cdf + c * pdf).to(BF))
        c = c * cdf
    elif act == ACT_GELU_BWD:
        c = c * aux.float()
    elif act == ACT_RELU:
        c = c.clamp_min(0)
    if residual is not None:
        c = c + residual
    if out_f32 is not None:
        if accumulate:
            out_f32 += c
        else:
            out_f32.copy_(c)
    if out_bf16 is not None:
        out_bf16.copy_(c.to(BF))
    if act == ACT_ROWDOT:          # per-64-column-block dots of the stored (bf16) output rows with aux
        M_, N_ = c.shape
        rowdot[:, :M_] = (c.to(BF).float() * aux.float()).view(M_, N_ // 64, 64).sum(-1).t()
    _count(1)


def layernorm_fwd(x, gamma, beta, eps, *, rows=None, ldx=None, y_bf16=None, y_f32=None, mean=None, rstd=None,
                  y_split=None, split_period=1):
    D = gamma.numel()
    rows = x.shape[0] if rows is None else rows
    ldx = x.stride(0) if ldx is None else ldx
    xv = _rows(x, rows, ldx, D)
    mu = xv.mean(1)
    var = ((xv - mu[:, None]) ** 2).mean(1)
    rs = torch.rsqrt(var + eps)
    y = (xv - mu[:, None]) * rs[:, None] * gamma + beta
    if y_f32 is not None:
        y_f32[:rows].copy_(y)
    if y_bf16 is not None:
        y_bf16[:rows].copy_(y.to(BF))
    if mean is not None:
        mean[:rows].copy_(mu)
    if rstd is not None:
        rstd[:rows].copy_(rs)
    if y_split is not None:                 # rows r with r % split_period == 0 -> y_split[r // split_period] = [hi | hi | lo]
        ys = y[::split_period]
        hi = ys.to(BF)
        k = ys.shape[0]
        y_split[:k, :D], y_split[:k, D:2 * D], y_split[:k, 2 * D:] = hi, hi, (ys - hi.float()).to(BF)
    _count(1)


def layernorm_bwd(x, mean, rstd, gamma, *, dy_bf16=None, dy_f32=None, rows=None, ldx=None, lddyf=None, add1=None,
                  add2=None, dx=None, dx_bf16=None, lddx=None, lddxb=None, dgamma=None, dbeta=None, dxsum=None):
    D = gamma.numel()
    rows = x.shape[0] if rows is None else rows
    ldx = x.stride(0) if ldx is None else ldx
    xv = _rows(x, rows, ldx, D)
    dy = torch.zeros(rows, D)
    if dy_bf16 is not None:
        dy = dy + dy_bf16[:rows].float()
    if dy_f32 is not None:
        dy = dy + _rows(dy_f32, rows, dy_f32.stride(0) if lddyf is None else lddyf, D)
    xh = (xv - mean[:rows, None]) * rstd[:rows, None]
    g = dy * gamma
    o = rstd[:rows, None] * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True))
    for add in (add1, add2):
        if add is not None:
            o = o + add[:rows]
    if dx is not None:
        _rows(dx, rows, dx.stride(0) if lddx is None else lddx, D).copy_(o)
    if dx_bf16 is not None:
        _rows(dx_bf16, rows, dx_bf16.stride(0) if lddxb is None else lddxb, D).copy_(o.to(BF))
    if dgamma is not None:
        dgamma += (dy * xh).sum(0)
    if dbeta is not None:
        dbeta += dy.sum(0)
    if dxsum is not None:
        dxsum += o.sum(0)
    _count(1)


def dropout_mask(n, p, seed, site, device=None):
    """Stand-in for the Philox draws: a torch generator seeded by (seed, site) - consistent inside the fake world."""
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(site)) % (2 ** 63))
    return (torch.rand(n, generator=g) >= p).to(torch.uint8)


def _mult(shape, p, seed, site):
    n = 1
    for d in shape:
        n *= d
    return dropout_mask(n, p, seed, site).view(shape).float() / (1.0 - p)


def dropout_fwd(x, p, seed, site, *, residual=None, out=None, out_bf16=None, out_split3=None, rows=None):
    v = x.float() * _mult(tuple(x.shape), p, seed, site)
    if residual is not None:
        v = v + residual
    cols = x.shape[1]
    if out is not None:
        out.copy_(v)
    if out_bf16 is not None:
        out_bf16.copy_(v.to(BF))
    if out_split3 is not None:
        hi = v.to(BF)
        out_split3[:, :cols], out_split3[:, cols:2 * cols], out_split3[:, 2 * cols:] = hi, hi, (v - hi.float()).to(BF)
    _count(1)


def dropout_bwd(p, seed, site, *, dy_f32=None, dy_bf16=None, dx_f32=None, dx_bf16=None):
    ref = dy_f32 if dy_f32 is not None else dy_bf16
    v = torch.zeros(ref.shape)
    if dy_f32 is not None:
        v = v + dy_f32
    if dy_bf16 is not None:
        v = v + dy_bf16.float()
    v = v * _mult(tuple(ref.shape), p, seed, site)
    if dx_f32 is not None:
        dx_f32.copy_(v)
    if dx_bf16 is not None:
        dx_bf16.copy_(v.to(BF))
    _count(1)


def attn_core_work(mode, B, T, H, F, n):
    return (0.0, 0.0)


def attn_fwd_workspace_floats(mode, B, H, F, n=1):
    return 1


def _attn(mode, B, T, H, F, n, q, k, v, key_mask, dropout=None):
    cfg = O.OracleCfg(heads=H, bf16=True)
    if mode == MODE_PLAIN:
        add = None
        if key_mask is not None:
            add = torch.zeros(B, 1, 1, T)
            add.masked_fill_(key_mask.view(B, 1, 1, T) == 0, float("-inf"))
        dm = _mult((B, H, T, T), dropout[0], dropout[1], dropout[2]) if dropout is not None else None
        out = O._softmax_attention(q, k, v, cfg, add, dm)                   # (B, H, T, d)
        return out.permute(0, 2, 1, 3).reshape(B, T, H * 64)
    return O.divided_attention_core(q, k, v, "space" if mode == MODE_SPACE else "time", F, n, cfg)


def attn_fwd(mode, B, T, H, F, n, qkv, out, lse, key_mask=None, cls_ws=None, dropout=None):
    x = qkv.float().view(B, T, 3, H, 64)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    out.copy_(_attn(mode, B, T, H, F, n, q, k, v, key_mask, dropout).reshape(B * T, H * 64).to(BF))
    _count(1)


def attn_bwd(mode, B, T, H, F, n, qkv, out, lse, dout, dqkv, scale, cls_acc=None, key_mask=None, dropout=None, delta=None):
    if delta is not None:          # the schedule must hand over delta = rowsum(dO * O) per head of THIS attention
        ref = (dout.float() * out.float()).view(B * T, H, 64).sum(-1).t()
        assert torch.allclose(delta[:, :B * T], ref, rtol=1e-4, atol=1e-5), "attn_bwd: stale or wrong delta"
    with torch.enable_grad():         # may run inside an autograd.Function.backward (grad mode off)
        x = qkv.float().view(B, T, 3, H, 64).clone().requires_grad_(True)
        q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        _attn(mode, B, T, H, F, n, q, k, v, key_mask, dropout).backward(dout.float().view(B, T, H * 64))
    g = x.grad.clone()
    g[:, :, 0] *= scale           # the q slot holds the scaled query: d/d(unscaled q) = scale * d/d(q)
    dqkv.copy_(g.reshape(B * T, 3 * H * 64).to(BF))
    _count(1)


def im2col_patches(video, out, P=16):
    B, Fr, C, H, W = video.shape
    gh, gw = H // P, W // P
    out.copy_(video.reshape(B * Fr, C, gh, P, gw, P).permute(0, 2, 4, 1, 3, 5).reshape(B * Fr * gh * gw, C * P * P).to(BF))
    _count(1)


def assemble_tokens(patch, obj, cls_token, pos_embed, temporal, type_embed, x, B, Fr, N, O, D):
    n = N + O
    xv = x.view(B, 1 + Fr * n, D)
    pos, tem = pos_embed.detach().view(-1, D), temporal.detach().view(-1, D)
    xv[:, 0] = cls_token.detach().view(D) + pos[0]
    body = xv[:, 1:].view(B, Fr, n, D)
    body[:, :, :N] = patch.view(B, Fr, N, D) + pos[1:1 + N].view(1, 1, N, D) + tem[:Fr].view(1, Fr, 1, D)
    if type_embed is not None:
        body[:, :, :N] += type_embed.detach()[0]
    if O > 0:
        body[:, :, N:] = obj.view(B, Fr, O, D) + tem[:Fr].view(1, Fr, 1, D)
        if type_embed is not None:
            body[:, :, N:] += type_embed.detach()[1]
    _count(1)


def assemble_tokens_bwd(dx, dpatch, dobj, dcls, dpos, dtemporal, dtype_embed, B, Fr, N, O, D):
    n = N + O
    dv = dx.view(B, 1 + Fr * n, D)
    body = dv[:, 1:].view(B, Fr, n, D)
    dpatch.copy_(body[:, :, :N].reshape(B * Fr * N, D).to(BF))
    if O > 0:
        dobj.copy_(body[:, :, N:].reshape(B * Fr * O, D).to(BF))
    dcls.view(D).add_(dv[:, 0].sum(0))
    dp = dpos.view(-1, D)
    dp[0] += dv[:, 0].sum(0)
    dp[1:1 + N] += body[:, :, :N].sum((0, 1))
    dtemporal.view(-1, D)[:Fr] += body.sum((0, 2))
    if dtype_embed is not None:
        dtype_embed[0] += body[:, :, :N].sum((0, 1, 2))
        if O > 0:
            dtype_embed[1] += body[:, :, N:].sum((0, 1, 2))
    _count(1)


def colsum_bf16(x, out):
    out += x.float().sum(0)
    _count(1)


def vecmat_f32(v, W, out):
    out += v.detach() @ W.detach()
    _count(1)


def unpack_wgrad(scratch, cols, dw, db):
    dw += scratch[:, :cols]
    db += scratch[:, cols]
    scratch.zero_()
    _count(1)


def text_embed(ids, word, pos, out, L):
    r = torch.arange(ids.numel())
    out.copy_(word.detach()[ids.view(-1)] + pos.detach()[r % L])
    _count(1)


def text_embed_bwd(ids, dsum, dword, dpos, L):
    dword.index_add_(0, ids.view(-1), dsum)
    dpos[:L] += dsum.view(-1, L, dsum.shape[1]).sum(0)
    _count(1)
