"""Python-side operator wrappers: torch tensors in, liboat C-ABI calls out (on the current CUDA stream).

Each wrapper only validates dtype/contiguity, fills the POD argument struct and launches; no arithmetic is done in
PyTorch here.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, check, lib, ptr, stream_ptr

ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_RELU, ACT_ROWDOT = 0, 1, 2, 3, 4

# Kernel-launch accounting (bench.py reports it) and an optional per-launch CUDA-event profiler for the roofline lines.
LAUNCHES = 0
GRAPH_REPLAYS = 0   # CUDA-graph replays of whole schedules (engine._GraphedSchedule); their kernels are counted in LAUNCHES
PROFILE = None      # when a list: (kind, work, start_event, end_event) is appended around every profiled launch


def _count(n):
    global LAUNCHES
    LAUNCHES += n


class _Prof:
    def __init__(self, kind, work):
        self.kind, self.work = kind, work

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.kind, self.work, self.e0, self.e1))
        return False


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a 2-D tensor with unit inner stride"
    return t.stride(0)


def gemm(A, B, *, a_major=0, b_major=0, alpha=1.0, bias=None, scale_cols=0, scale=1.0, act=ACT_NONE, aux=None,
         residual=None, out_f32=None, out_bf16=None, out2_bf16=None, accumulate=False, split_k=0, rowdot=None):
    """C[M,N] = alpha * A(M,K) . B(N,K)^T with the fused epilogue of oat_gemm_bf16 (include/oat.h).

    a_major=0: A is [M,K]; a_major=1: A is stored [K,M] (M contiguous). Same for B with N.
    act=ACT_ROWDOT: rowdot (fp32 [N // 64, >= M]) receives the per-64-column-block dots of the bf16 output rows with aux.
    """
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    if a_major == 0:
        M, K = A.shape
    else:
        K, M = A.shape
    if b_major == 0:
        N, Kb = B.shape
    else:
        Kb, N = B.shape
    assert K == Kb, "contraction mismatch: %d vs %d" % (K, Kb)
    a = GemmArgs()
    a.A, a.lda, a.a_major = ptr(A), _ld(A), a_major
    a.B, a.ldb, a.b_major = ptr(B), _ld(B), b_major
    a.M, a.N, a.K = M, N, K
    a.alpha = alpha
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    a.bias = ptr(bias)
    a.scale_cols, a.scale = scale_cols, scale
    a.act = act
    if aux is not None:
        assert aux.dtype == torch.bfloat16 and aux.shape == (M, N)
        a.aux_bf16, a.ld_aux = ptr(aux), _ld(aux)
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (M, N)
        a.residual, a.ldr = ptr(residual), _ld(residual)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.shape == (M, N)
        a.out_f32, a.ld_f32 = ptr(out_f32), _ld(out_f32)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.shape == (M, N)
        a.out_bf16, a.ld_bf16 = ptr(out_bf16), _ld(out_bf16)
    if out2_bf16 is not None:
        assert out2_bf16.dtype == torch.bfloat16 and out2_bf16.shape == (M, N)
        a.out2_bf16, a.ld2 = ptr(out2_bf16), _ld(out2_bf16)
    a.accumulate = 1 if accumulate else 0
    a.split_k = split_k
    if rowdot is not None:
        assert act == ACT_ROWDOT and rowdot.dtype == torch.float32 and rowdot.shape[0] == N // 64 and rowdot.shape[1] >= M
        a.rowdot, a.ld_rowdot = ptr(rowdot), _ld(rowdot)
    _count(1)
    # products with <= 64 rows take the mma.sync weight-stream kernel (gemm_skinny.cu) unless they accumulate: a different
    # kernel, profiled under its own name so that "gemm" is the tcgen05 kernel only
    skinny = M <= 64 and a_major == 0 and b_major == 0 and not accumulate and K % 8 == 0
    with _Prof("gemm_skinny" if skinny else "gemm", 2.0 * M * N * K):
        check(lib().oat_gemm_bf16(ctypes.byref(a), stream_ptr()), "oat_gemm_bf16")


# ------------------------------------------------------------------------------------------------ LayerNorm
_i64, _i32, _f32, _vp, _sz = ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t


def layernorm_fwd(x, gamma, beta, eps, *, rows=None, ldx=None, y_bf16=None, y_f32=None, mean=None, rstd=None,
                  y_split=None, split_period=1):
    """Row LayerNorm of fp32 x (rows x D, row pitch ldx) -> bf16 and/or fp32 outputs, plus mean / rstd.
    y_split (bf16 [ceil(rows / split_period), 3*D]): split-bf16 copy [hi | hi | lo] of every split_period-th row."""
    D = gamma.numel()
    rows = x.shape[0] if rows is None else rows
    ldx = x.stride(0) if ldx is None else ldx
    if y_split is not None:
        assert y_split.dtype == torch.bfloat16 and y_split.shape[1] == 3 * D
        assert y_split.shape[0] * split_period >= rows
    _count(1)
    check(lib().oat_layernorm_fwd(
        ptr(x), _i64(ldx), ptr(gamma), ptr(beta), _f32(eps), _i64(rows), _i32(D),
        ptr(y_bf16), _i64(y_bf16.stride(0) if y_bf16 is not None else 0),
        ptr(y_f32), _i64(y_f32.stride(0) if y_f32 is not None else 0),
        ptr(mean), ptr(rstd), ptr(y_split), _i64(y_split.stride(0) if y_split is not None else 0),
        _i64(split_period), stream_ptr()), "oat_layernorm_fwd")


def layernorm_bwd(x, mean, rstd, gamma, *, dy_bf16=None, dy_f32=None, rows=None, ldx=None, lddyf=None, add1=None,
                  add2=None, dx=None, dx_bf16=None, lddx=None, lddxb=None, dgamma=None, dbeta=None, dxsum=None):
    D = gamma.numel()
    rows = x.shape[0] if rows is None else rows
    ldx = x.stride(0) if ldx is None else ldx
    ldadd = add1.stride(0) if add1 is not None else (add2.stride(0) if add2 is not None else 0)
    if add1 is not None and add2 is not None:
        assert add1.stride(0) == add2.stride(0)
    _count(1)
    check(lib().oat_layernorm_bwd(
        ptr(dy_bf16), _i64(dy_bf16.stride(0) if dy_bf16 is not None else 0),
        ptr(dy_f32), _i64((dy_f32.stride(0) if lddyf is None else lddyf) if dy_f32 is not None else 0),
        ptr(x), _i64(ldx), ptr(mean), ptr(rstd), ptr(gamma), _i64(rows), _i32(D),
        ptr(add1), ptr(add2), _i64(ldadd),
        ptr(dx), _i64((dx.stride(0) if lddx is None else lddx) if dx is not None else 0),
        ptr(dx_bf16), _i64((dx_bf16.stride(0) if lddxb is None else lddxb) if dx_bf16 is not None else 0),
        ptr(dgamma), ptr(dbeta), ptr(dxsum), stream_ptr()), "oat_layernorm_bwd")


# ------------------------------------------------------------------------------------------------ attention
class AttnArgs(ctypes.Structure):
    _fields_ = [
        ("mode", _i32), ("B", _i32), ("T", _i32), ("H", _i32), ("F", _i32), ("n", _i32),
        ("qkv", _vp), ("ld_qkv", _i64),
        ("out", _vp), ("ld_out", _i64),
        ("lse", _vp),
        ("key_mask", _vp),
        ("dout", _vp), ("ld_dout", _i64),
        ("dqkv", _vp), ("ld_dqkv", _i64),
        ("scale", _f32),
        ("cls_acc", _vp),
        ("dropout_p", _f32), ("dropout_site", ctypes.c_uint32), ("dropout_seed", ctypes.c_uint64),
        ("delta", _vp), ("ld_delta", _i64),
    ]


MODE_SPACE, MODE_TIME, MODE_PLAIN = 0, 1, 2


def _attn_args(mode, B, T, H, F, n, qkv, out, lse, key_mask):
    a = AttnArgs()
    a.mode, a.B, a.T, a.H, a.F, a.n = mode, B, T, H, F, n
    assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and lse.dtype == torch.float32
    a.qkv, a.ld_qkv = ptr(qkv), qkv.stride(0)
    a.out, a.ld_out = ptr(out), out.stride(0)
    a.lse = ptr(lse)
    if key_mask is not None:
        assert key_mask.dtype == torch.int32 and key_mask.numel() == B * T
    a.key_mask = ptr(key_mask)
    return a


def attn_core_work(mode, B, T, H, F, n):
    """(algorithmic bf16 bytes, FLOPs) of one forward attention core over all groups:
    per group bytes = 2 B * 64 * (q rows + k rows + v rows + out rows), flops = 4 * nq * nk * 64."""
    if mode == MODE_SPACE:
        groups, nq, nk = B * H * F, n, n + 1
    elif mode == MODE_TIME:
        groups, nq, nk = B * H * n, F, F + 1
    else:
        groups, nq, nk = B * H, T, T
    return (groups * 128.0 * (2 * nq + 2 * nk), groups * 4.0 * nq * nk * 64)


def attn_bwd_core_work(mode, B, T, H, F, n, ext_delta=False):
    """(algorithmic bytes, forward FLOPs) of one attention backward: q, dO, O, dQ rows + k, v, dK, dV rows = twice the
    forward's bytes; with delta handed in (oat_attn_args.delta) the O rows are not read - 4 B of delta per query row are."""
    by, fl = attn_core_work(mode, B, T, H, F, n)
    if mode == MODE_SPACE:
        groups, nq = B * H * F, n
    elif mode == MODE_TIME:
        groups, nq = B * H * n, F
    else:
        groups, nq = B * H, T
    by = 2.0 * by
    if ext_delta:
        by += groups * nq * (4.0 - 128.0)
    return (by, fl)


def attn_fwd_workspace_floats(mode, B, H, F, n=1):
    """fp32 words of workspace the space / time forward needs for the partials of the fused CLS query."""
    f = lib().oat_attn_fwd_workspace_floats
    f.restype = _sz
    return int(f(_i32(mode), _i32(B), _i32(H), _i32(F), _i32(n)))


def _set_dropout(a, mode, dropout):
    if dropout is not None and dropout[0] > 0.0:
        assert mode == MODE_PLAIN, "attention-weight dropout exists in the text tower only"
        a.dropout_p, a.dropout_seed, a.dropout_site = float(dropout[0]), int(dropout[1]), int(dropout[2])


def attn_fwd(mode, B, T, H, F, n, qkv, out, lse, key_mask=None, cls_ws=None, dropout=None):
    """qkv bf16 [B*T, 3*H*64] (q pre-scaled) -> out bf16 [B*T, H*64], lse fp32 [B*H*T].
    cls_ws (fp32, attn_fwd_workspace_floats) selects the tcgen05 space kernel with the CLS query fused in.
    dropout = (p, seed, site): mode 2 only, Philox dropout on the softmax weights (same triple in attn_bwd)."""
    a = _attn_args(mode, B, T, H, F, n, qkv, out, lse, key_mask)
    _set_dropout(a, mode, dropout)
    if cls_ws is not None:
        assert cls_ws.dtype == torch.float32 and cls_ws.numel() >= attn_fwd_workspace_floats(mode, B, H, F, n)
        a.cls_acc = ptr(cls_ws)
    _count(1 if mode == MODE_PLAIN else 2)
    with _Prof("attn_fwd_%d" % mode, attn_core_work(mode, B, T, H, F, n)):
        check(lib().oat_attn_fwd(ctypes.byref(a), stream_ptr()), "oat_attn_fwd")


def attn_bwd(mode, B, T, H, F, n, qkv, out, lse, dout, dqkv, scale, cls_acc=None, key_mask=None, dropout=None, delta=None):
    """delta (fp32 [H, >= B*T], modes 0/1): rowsum(dout * out) per head from the GEMM that produced dout (gemm act=ACT_ROWDOT)."""
    a = _attn_args(mode, B, T, H, F, n, qkv, out, lse, key_mask)
    if delta is not None:
        assert mode != MODE_PLAIN and delta.dtype == torch.float32 and delta.shape[0] == H and delta.shape[1] >= B * T
        a.delta, a.ld_delta = ptr(delta), _ld(delta)
    _set_dropout(a, mode, dropout)
    assert dout.dtype == torch.bfloat16 and dqkv.dtype == torch.bfloat16
    a.dout, a.ld_dout = ptr(dout), dout.stride(0)
    a.dqkv, a.ld_dqkv = ptr(dqkv), dqkv.stride(0)
    a.scale = scale
    a.cls_acc = ptr(cls_acc)
    _count(1 if mode == MODE_PLAIN else 2)
    with _Prof("attn_bwd_%d" % mode, attn_bwd_core_work(mode, B, T, H, F, n, delta is not None)):
        check(lib().oat_attn_bwd(ctypes.byref(a), stream_ptr()), "oat_attn_bwd")


# ------------------------------------------------------------------------------------------------ packing / tokens
def cast_bf16(src, dst, *, rows=None, cols=None, lds=None, relu=False):
    """dst (bf16, 2-D) <- src (fp32). Columns [cols, dst.shape[1]) are zero-filled."""
    rows = dst.shape[0] if rows is None else rows
    cols = src.shape[-1] if cols is None else cols
    lds = (src.stride(0) if src.dim() == 2 else cols) if lds is None else lds
    _count(1)
    check(lib().oat_cast_bf16(ptr(src), _i64(lds), ptr(dst), _i64(dst.stride(0)), _i64(rows), _i32(cols),
                              _i32(dst.shape[1]), _i32(1 if relu else 0), stream_ptr()), "oat_cast_bf16")


def split3_bf16(src, dst, *, rows=None, cols=None, lds=None, relu=False):
    """dst (bf16 [rows, 3*cols]) <- split-bf16 activation operand [hi | hi | lo] of src (fp32), optional ReLU first."""
    rows = dst.shape[0] if rows is None else rows
    cols = src.shape[-1] if cols is None else cols
    lds = (src.stride(0) if src.dim() == 2 else cols) if lds is None else lds
    assert dst.dtype == torch.bfloat16 and dst.shape[1] == 3 * cols and src.dtype == torch.float32
    _count(1)
    check(lib().oat_split3_bf16(ptr(src), _i64(lds), ptr(dst), _i64(dst.stride(0)), _i64(rows), _i32(cols),
                                _i32(1 if relu else 0), stream_ptr()), "oat_split3_bf16")


_u64, _u32 = ctypes.c_uint64, ctypes.c_uint32


def dropout_fwd(x, p, seed, site, *, residual=None, out=None, out_bf16=None, out_split3=None, rows=None):
    """out = keep(x) / (1 - p) + residual, optionally also as bf16 / split-bf16 [hi | hi | lo] (include/oat.h)."""
    rows = x.shape[0] if rows is None else rows
    cols = x.shape[1]
    assert x.dtype == torch.float32 and x.stride(1) == 1
    _count(1)
    check(lib().oat_dropout_fwd(ptr(x), _i64(x.stride(0)), ptr(residual), _i64(residual.stride(0) if residual is not None else 0),
                                ptr(out), _i64(out.stride(0) if out is not None else 0),
                                ptr(out_bf16), _i64(out_bf16.stride(0) if out_bf16 is not None else 0),
                                ptr(out_split3), _i64(out_split3.stride(0) if out_split3 is not None else 0),
                                _i64(rows), _i32(cols), _f32(p), _u64(seed), _u32(site), stream_ptr()), "oat_dropout_fwd")


def dropout_bwd(p, seed, site, *, dy_f32=None, dy_bf16=None, dx_f32=None, dx_bf16=None):
    """dx = keep(dy_f32 + dy_bf16) / (1 - p) as fp32 and / or bf16."""
    ref = dy_f32 if dy_f32 is not None else dy_bf16
    rows, cols = ref.shape
    _count(1)
    check(lib().oat_dropout_bwd(ptr(dy_f32), _i64(dy_f32.stride(0) if dy_f32 is not None else 0),
                                ptr(dy_bf16), _i64(dy_bf16.stride(0) if dy_bf16 is not None else 0),
                                ptr(dx_f32), _i64(dx_f32.stride(0) if dx_f32 is not None else 0),
                                ptr(dx_bf16), _i64(dx_bf16.stride(0) if dx_bf16 is not None else 0),
                                _i64(rows), _i32(cols), _f32(p), _u64(seed), _u32(site), stream_ptr()), "oat_dropout_bwd")


def dropout_mask(n, p, seed, site, device):
    """uint8 keep flags of elements 0..n-1 of a dropout site: the draws every liboat kernel makes for (seed, site)."""
    keep = torch.empty(n, dtype=torch.uint8, device=device)
    _count(1)
    check(lib().oat_dropout_mask(ptr(keep), _i64(n), _f32(p), _u64(seed), _u32(site), stream_ptr()), "oat_dropout_mask")
    return keep


class CastPlan:
    """A set of fp32 -> bf16 (or fp32 -> fp32) 2-D copies executed by ONE launch (oat_cast_multi). The device table is
    built once and re-used for as long as the source / destination pointers stay the same."""

    def __init__(self):
        self.items = []
        self._key = None
        self._dev = None

    def add(self, src, dst, rows=None, cols=None, split=False):
        """dst[:rows, :] <- src (2-D view, last dim contiguous); columns [cols, dst.shape[1]) are zero-filled.
        split=True: dst is [rows, 3*w] and receives the split-bf16 weight copy [hi | lo | hi] (segments w wide)."""
        src2 = src.detach()
        src2 = src2.reshape(src2.shape[0], -1) if src2.dim() != 2 else src2
        rows = src2.shape[0] if rows is None else rows
        cols = src2.shape[1] if cols is None else cols
        assert src2.dtype == torch.float32 and src2.stride(1) == 1 and dst.dim() == 2 and dst.stride(1) == 1
        width = dst.shape[1] // 3 if split else dst.shape[1]
        assert (not split) or (dst.shape[1] % 3 == 0 and dst.dtype == torch.bfloat16)
        assert width % 4 == 0 and dst.stride(0) % 4 == 0 and width >= cols
        self.items.append((src2, dst, rows, cols, split))

    def run(self):
        if not self.items:
            return
        key = tuple((s.data_ptr(), d.data_ptr(), r, c, sp) for s, d, r, c, sp in self.items)
        if key != self._key:
            rows_, prefix = [], [0]
            for s, d, r, c, sp in self.items:
                colsp = d.shape[1] // 3 if sp else d.shape[1]
                rows_.append([s.data_ptr(), d.data_ptr(), r, c, colsp, s.stride(0), d.stride(0),
                              2 if sp else (1 if d.dtype == torch.float32 else 0)])
                prefix.append(prefix[-1] + (r * colsp + 1023) // 1024)
            dev = self.items[0][1].device
            self._dev = (torch.tensor(rows_, dtype=torch.int64, device=dev),
                         torch.tensor(prefix, dtype=torch.int64, device=dev), len(rows_), prefix[-1])
            self._key = key
        table, prefix, n, total = self._dev
        _count(1)
        check(lib().oat_cast_multi(ptr(table), ptr(prefix), _i32(n), _i64(total), stream_ptr()), "oat_cast_multi")
        self.items = []


def relu_bwd(x, dy_bf16, dx, *, rows, cols, ldx, lddx=None):
    _count(1)
    check(lib().oat_relu_bwd(ptr(x), _i64(ldx), ptr(dy_bf16), _i64(dy_bf16.stride(0)), ptr(dx),
                             _i64(dx.stride(0) if lddx is None else lddx),
                             _i64(rows), _i32(cols), stream_ptr()), "oat_relu_bwd")


def im2col_patches(video, out, P=16):
    B, Fr, C, H, W = video.shape
    assert video.is_contiguous() and video.dtype == torch.float32
    _count(1)
    check(lib().oat_im2col_patches(ptr(video), ptr(out), _i64(B * Fr), _i32(C), _i32(H), _i32(W), _i32(P),
                                   stream_ptr()), "oat_im2col_patches")


def assemble_tokens(patch, obj, cls_token, pos_embed, temporal, type_embed, x, B, Fr, N, O, D):
    _count(1)
    check(lib().oat_assemble_tokens(ptr(patch), ptr(obj), ptr(cls_token), ptr(pos_embed), ptr(temporal),
                                    ptr(type_embed), ptr(x), _i32(B), _i32(Fr), _i32(N), _i32(O), _i32(D),
                                    stream_ptr()), "oat_assemble_tokens")


def assemble_tokens_bwd(dx, dpatch, dobj, dcls, dpos, dtemporal, dtype_embed, B, Fr, N, O, D):
    _count(1)
    check(lib().oat_assemble_tokens_bwd(ptr(dx), ptr(dpatch), ptr(dobj), ptr(dcls), ptr(dpos), ptr(dtemporal),
                                        ptr(dtype_embed), _i32(B), _i32(Fr), _i32(N), _i32(O), _i32(D),
                                        stream_ptr()), "oat_assemble_tokens_bwd")


def colsum_bf16(x, out):
    """out[c] += sum_r x[r, c] (x bf16 2-D, out fp32)."""
    _count(1)
    check(lib().oat_colsum_bf16(ptr(x), _i64(x.stride(0)), _i64(x.shape[0]), _i32(x.shape[1]), ptr(out),
                                stream_ptr()), "oat_colsum_bf16")


def vecmat_f32(v, W, out):
    """out[c] += sum_k v[k] * W[k, c] (fp32; W 2-D with unit inner stride)."""
    assert v.dtype == torch.float32 and W.dtype == torch.float32 and out.dtype == torch.float32
    K, N = W.shape
    assert v.numel() == K and out.numel() == N and W.stride(1) == 1 and v.is_contiguous() and out.is_contiguous()
    _count(1)
    check(lib().oat_vecmat_f32(ptr(v), ptr(W), _i64(W.stride(0)), _i32(K), _i32(N), ptr(out), stream_ptr()), "oat_vecmat_f32")


def unpack_wgrad(scratch, cols, dw, db):
    """dw (+)= scratch[:, :cols]; db (+)= scratch[:, cols]; scratch <- 0 (see oat_unpack_wgrad)."""
    rows = scratch.shape[0]
    assert scratch.dtype == torch.float32 and dw.dtype == torch.float32 and db.dtype == torch.float32
    assert dw.shape == (rows, cols) and db.numel() == rows and scratch.shape[1] >= cols + 4
    _count(1)
    check(lib().oat_unpack_wgrad(ptr(scratch), _i64(scratch.stride(0)), _i32(cols), ptr(dw), _i64(dw.stride(0)), ptr(db),
                                 _i64(rows), stream_ptr()), "oat_unpack_wgrad")


def text_embed(ids, word, pos, out, L):
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    _count(1)
    check(lib().oat_text_embed(ptr(ids), ptr(word), ptr(pos), ptr(out), _i64(ids.numel()), _i32(L),
                               _i32(word.shape[1]), stream_ptr()), "oat_text_embed")


def text_embed_bwd(ids, dsum, dword, dpos, L):
    _count(1)
    check(lib().oat_text_embed_bwd(ptr(ids), ptr(dsum), ptr(dword), ptr(dpos), _i64(ids.numel()), _i32(L),
                                   _i32(dsum.shape[1]), stream_ptr()), "oat_text_embed_bwd")


# ------------------------------------------------------------------------------------------------ loss
def infonce_workspace_bytes(n, P):
    f = lib().oat_infonce_workspace_bytes
    f.restype = _sz
    return int(f(_i32(n), _i32(P)))


def infonce_fwd_bwd(text, video, temperature=0.05, eps=1e-8, want_sims=False, want_grad=True, workspace=None):
    """text, video: fp32 [n, P] gathered embeddings -> (loss[1], sims or None, dtext or None, dvideo or None)."""
    n, P = text.shape
    assert text.dtype == torch.float32 and video.shape == (n, P) and text.is_contiguous() and video.is_contiguous()
    nbytes = infonce_workspace_bytes(n, P)
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=text.device)
    loss = torch.empty(1, dtype=torch.float32, device=text.device)
    sims = torch.empty(n, n, dtype=torch.float32, device=text.device) if want_sims else None
    dt = torch.empty_like(text) if want_grad else None
    dv = torch.empty_like(video) if want_grad else None
    _count(9)
    check(lib().oat_infonce_fwd_bwd(ptr(text), ptr(video), _i32(n), _i32(P), _f32(temperature), _f32(eps), ptr(sims),
                                    ptr(loss), ptr(dt), ptr(dv), ptr(workspace), _sz(nbytes), stream_ptr()),
          "oat_infonce_fwd_bwd")
    return loss, sims, dt, dv


def sim_workspace_bytes(n, m, P):
    f = lib().oat_sim_workspace_bytes
    f.restype = _sz
    return int(f(_i32(n), _i32(m), _i32(P)))


def sim_matrix_fwd(a, b, eps, sims, ws):
    n, P = a.shape
    m = b.shape[0]
    _count(3)
    check(lib().oat_sim_matrix_fwd(ptr(a), ptr(b), _i32(n), _i32(m), _i32(P), _f32(eps), ptr(sims), ptr(ws),
                                   _sz(ws.numel()), stream_ptr()), "oat_sim_matrix_fwd")


def sim_matrix_bwd(dsims, eps, da, db, ws, n, m, P):
    _count(2)
    check(lib().oat_sim_matrix_bwd(ptr(dsims), _i32(n), _i32(m), _i32(P), _f32(eps), ptr(da), ptr(db), ptr(ws),
                                   _sz(ws.numel()), stream_ptr()), "oat_sim_matrix_bwd")


def norm_softmax_loss(sims, temperature, loss, dsims):
    n = sims.shape[0]
    scratch = torch.empty(2 * n, dtype=torch.float32, device=sims.device)
    _count(3)
    check(lib().oat_norm_softmax_loss(ptr(sims), _i32(n), _i64(sims.stride(0)), _f32(temperature), ptr(loss),
                                      ptr(dsims), ptr(scratch), stream_ptr()), "oat_norm_softmax_loss")


def retrieval_ranks(sims):
    """sims: fp32 CUDA [n, n] (rows = text queries, columns = videos) -> (t2v_rank[n], v2t_rank[n]) fp32, the column of
    the ground-truth pair in the sorted distance row with the reference tie rules (model/metric.py:62-69, 153, 183)."""
    assert sims.dtype == torch.float32 and sims.dim() == 2 and sims.shape[0] == sims.shape[1] and sims.stride(1) == 1
    n = sims.shape[0]
    t2v = torch.empty(n, dtype=torch.float32, device=sims.device)
    v2t = torch.empty(n, dtype=torch.float32, device=sims.device)
    _count(1)
    check(lib().oat_retrieval_ranks(ptr(sims), _i32(n), _i64(sims.stride(0)), ptr(t2v), ptr(v2t), stream_ptr()),
          "oat_retrieval_ranks")
    return t2v, v2t


# ------------------------------------------------------------------------------------------------ object -> patch
XATTN_MASK, XATTN_SIGMOID, XATTN_SOFTMAX = 0, 1, 2


def object_patch_attention(q, k, v=None, mode="softmax", masks=None, want_weights=True):
    """One op, three score->weight modes. q (B,O,C), k (B,L,C), v (B,L,Cv), masks (B,O,L): fp32 CUDA.
    Returns (weights (B,O,L) or None, out (B,O,Cv) or None)."""
    m = {"mask": XATTN_MASK, "sigmoid": XATTN_SIGMOID, "softmax": XATTN_SOFTMAX}[mode]
    ref = masks if m == XATTN_MASK else q
    B, O = ref.shape[0], ref.shape[1]
    L = masks.shape[2] if m == XATTN_MASK else k.shape[1]
    C = 0 if m == XATTN_MASK else q.shape[2]
    dev = ref.device
    for t in (q, k, v, masks):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    weights = torch.empty(B, O, L, dtype=torch.float32, device=dev) if want_weights else None
    out = torch.empty(B, O, v.shape[2], dtype=torch.float32, device=dev) if v is not None else None
    _count(1)
    check(lib().oat_object_patch_attn(ptr(q), ptr(k), ptr(v), ptr(masks), ptr(weights), ptr(out), _i32(B), _i32(O),
                                      _i32(L), _i32(C), _i32(v.shape[2] if v is not None else 0), _i32(m),
                                      stream_ptr()), "oat_object_patch_attn")
    return weights, out


def object_patch_attention_bwd(q, k, v, weights, mode, dweights=None, dout=None, need_q=True, need_k=True,
                               need_v=True):
    """Gradients of object_patch_attention: returns (dq, dk, dv), each None when not applicable / not requested."""
    m = {"mask": XATTN_MASK, "sigmoid": XATTN_SIGMOID, "softmax": XATTN_SOFTMAX}[mode]
    B, O, L = weights.shape
    dev = weights.device
    for t in (q, k, v, weights, dweights, dout):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    C = 0 if m == XATTN_MASK else q.shape[2]
    Cv = v.shape[2] if v is not None else 0
    dq = torch.empty_like(q) if (m != XATTN_MASK and need_q) else None
    dk = torch.empty_like(k) if (m != XATTN_MASK and need_k) else None
    dv = torch.empty_like(v) if (v is not None and dout is not None and need_v) else None
    ds = torch.empty(B, O, L, dtype=torch.float32, device=dev) if m != XATTN_MASK else None
    _count(2 if m != XATTN_MASK else 1)
    check(lib().oat_object_patch_attn_bwd(ptr(q), ptr(k), ptr(v), ptr(weights), ptr(dweights), ptr(dout), ptr(dq),
                                          ptr(dk), ptr(dv), ptr(ds), _i32(B), _i32(O), _i32(L), _i32(C), _i32(Cv),
                                          _i32(m), stream_ptr()), "oat_object_patch_attn_bwd")
    return dq, dk, dv


def token_pool(cls, tok, a, b):
    """out (B, P) = a * cls + b * mean over tokens of tok (B, L, P); tok may be a [:, 1:] style slice (any batch / token
    pitch, unit inner stride); cls (B, P) or None."""
    B, L, P = tok.shape
    assert tok.dtype == torch.float32 and tok.stride(2) == 1
    assert cls is None or (cls.dtype == torch.float32 and cls.shape == (B, P) and cls.stride(1) == 1)
    out = torch.empty(B, P, dtype=torch.float32, device=tok.device)
    _count(1)
    check(lib().oat_token_pool(ptr(cls), _i64(cls.stride(0) if cls is not None else 0), ptr(tok), _i64(tok.stride(0)),
                               _i64(tok.stride(1)), ptr(out), _i32(B), _i32(L), _i32(P), _f32(a), _f32(b),
                               stream_ptr()), "oat_token_pool")
    return out


def token_pool_bwd(dout, B, L, P, a, b, need_cls=True):
    assert dout.dtype == torch.float32 and dout.is_contiguous() and dout.shape == (B, P)
    dcls = torch.empty(B, P, dtype=torch.float32, device=dout.device) if need_cls else None
    dtok = torch.empty(B, L, P, dtype=torch.float32, device=dout.device)
    _count(1)
    check(lib().oat_token_pool_bwd(ptr(dout), ptr(dcls), ptr(dtok), _i64(L * P), _i64(P), _i32(B), _i32(L), _i32(P),
                                   _f32(a), _f32(b), stream_ptr()), "oat_token_pool_bwd")
    return dcls, dtok


def bce_sum(p, target, scale, want_grad=True):
    """scale * BCELoss(reduction='sum')(p, target) -> (loss[1], dp or None); torch's clamps."""
    assert p.dtype == torch.float32 and target.dtype == torch.float32 and p.is_contiguous() and target.is_contiguous()
    assert p.numel() == target.numel()
    loss = torch.empty(1, dtype=torch.float32, device=p.device)
    dp = torch.empty_like(p) if want_grad else None
    _count(1)
    check(lib().oat_bce_sum(ptr(p), ptr(target), _i64(p.numel()), _f32(scale), ptr(loss), ptr(dp), stream_ptr()),
          "oat_bce_sum")
    return loss, dp


def patch_masks_from_bbox(boxes, patch_rows=14):
    """boxes: fp64 CUDA tensor [n, >=4] with (x1, y1, x2, y2) in [0,1] -> fp32 masks [n, patch_rows^2] (bit-exact
    with base/base_dataset_global_local.py:348-356)."""
    assert boxes.dtype == torch.float64 and boxes.dim() == 2 and boxes.is_contiguous()
    n = boxes.shape[0]
    masks = torch.empty(n, patch_rows * patch_rows, dtype=torch.float32, device=boxes.device)
    _count(1)
    check(lib().oat_patch_masks_from_bbox(ptr(boxes), _i32(boxes.shape[1]), ptr(masks), _i32(n), _i32(patch_rows),
                                          stream_ptr()), "oat_patch_masks_from_bbox")
    return masks


def patch_masks_same_class(boxes, classes, sel, patch_rows=14):
    """base/base_dataset_region_mem.py:233-247 on the device: boxes fp64 [n, >=4] in [0,1], classes int32 [n], sel int32
    [para] (the indices random.sample drew) -> fp32 masks [para, patch_rows^2], each the union over the boxes of the
    selected box's class. Bit-exact."""
    assert boxes.dtype == torch.float64 and boxes.dim() == 2 and boxes.is_contiguous()
    assert classes.dtype == torch.int32 and sel.dtype == torch.int32 and classes.numel() == boxes.shape[0]
    para = sel.numel()
    masks = torch.empty(para, patch_rows * patch_rows, dtype=torch.float32, device=boxes.device)
    _count(1)
    check(lib().oat_patch_masks_same_class(ptr(boxes), _i32(boxes.shape[1]), ptr(classes), ptr(sel), ptr(masks),
                                           _i32(boxes.shape[0]), _i32(para), _i32(patch_rows), stream_ptr()),
          "oat_patch_masks_same_class")
    return masks


def object_tags_masks(token_lens, indices):
    """base/base_dataset_global_local.py:395-405: (ends fp32 [k] = running end offset of each tag's tokens, total)."""
    assert token_lens.dtype == torch.float64 and indices.dtype == torch.int64
    k = indices.numel()
    ends = torch.empty(k, dtype=torch.float32, device=indices.device)
    total = torch.empty(1, dtype=torch.int32, device=indices.device)
    _count(1)
    check(lib().oat_object_tags_masks(ptr(token_lens), ptr(indices), ptr(ends), ptr(total), _i32(k), stream_ptr()),
          "oat_object_tags_masks")
    return ends, int(total.item())


def region_features_topk(x, bbox, conf, ids, image_w, image_h, top_k=10, v=1):
    """base/base_dataset.py:593-650 (read_object_from_disk after np.load) on the device -> fp32 [top_k, 2054 (+ pad)]."""
    n, fdim = x.shape
    for t in (x, bbox, conf):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert ids is None or (ids.dtype == torch.int64 and ids.numel() == n)
    ld = fdim + top_k + 6
    out = torch.zeros(top_k, ld, dtype=torch.float32, device=x.device)
    m = torch.empty(1, dtype=torch.int32, device=x.device)
    _count(1)
    check(lib().oat_region_features_topk(ptr(x), ptr(bbox), ptr(conf), ptr(ids), _i32(n), _i32(fdim), _i32(top_k),
                                         _i32(v), _i32(int(image_w)), _i32(int(image_h)), ptr(out), _i64(ld), ptr(m),
                                         stream_ptr()), "oat_region_features_topk")
    res = max(0, top_k - int(m.item()))
    return out[:, :fdim + res + 6].contiguous()


# ------------------------------------------------------------------------------------------------ profiling of the rest
def _profiled(kind, fn):
    """Per-launch CUDA-event timing (only while PROFILE is a list) for the ops that do not time themselves."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        if PROFILE is None:
            return fn(*a, **k)
        with _Prof(kind, 0.0):
            return fn(*a, **k)
    return wrapper


for _kind, _names in (("ln_fwd", ("layernorm_fwd",)), ("ln_bwd", ("layernorm_bwd",)), ("colsum", ("colsum_bf16", "unpack_wgrad")),
                      ("cast", ("cast_bf16", "split3_bf16", "relu_bwd")),
                      ("embed", ("im2col_patches", "assemble_tokens", "assemble_tokens_bwd", "text_embed", "text_embed_bwd")),
                      ("loss", ("infonce_fwd_bwd", "sim_matrix_fwd", "sim_matrix_bwd", "norm_softmax_loss"))):
    for _n in _names:
        globals()[_n] = _profiled(_kind, globals()[_n])
CastPlan.run = _profiled("cast_multi", CastPlan.run)
