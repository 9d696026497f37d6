"""autograd glue: each tower, the similarity matrix, the InfoNCE loss and the embedding all-gather are ONE
torch.autograd.Function whose forward/backward are liboat launches (engine.py / ops.py). Nothing here computes.

Reference semantics kept:
  * AllGather_multi (OATrans/trainer/trainer_dist.py:29-45): forward all-gathers and concatenates in rank order;
    backward returns the local slice of the incoming gradient with NO reduction.
  * sim_matrix (OATrans/model/model.py:164-172) and NormSoftmaxLoss (OATrans/model/loss.py:7-25) stay separate
    callables, as the trainer invokes them separately (trainer_dist.py:161-162).
"""
import ctypes

import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr
from .engine import GradBook

_i32, _f32, _sz = ctypes.c_int32, ctypes.c_float, ctypes.c_size_t


class _TowerFn(torch.autograd.Function):
    """forward(runner, *params): runner.fwd() launches the tower forward; backward returns one fp32 gradient per
    parameter, in order (views of the runner's flat gradient book)."""

    @staticmethod
    def forward(ctx, runner, *params):
        ctx.runner = runner
        ctx.set_materialize_grads(False)
        return runner.fwd()          # embeddings, or (embeddings, token features) when the engine was asked for tokens

    @staticmethod
    def backward(ctx, dout, dtokens=None):
        grads = ctx.runner.bwd(dout, dtokens)
        return (None,) + tuple(grads)


class TowerRunner:
    def __init__(self, engine, named_params, fwd_kwargs):
        self.engine = engine
        self.named = named_params                      # list of (name, Parameter)
        self.pdict = {n: p for n, p in named_params}
        self.kw = fwd_kwargs
        self._book = None

    def fwd(self):
        out = self.engine.forward(self.pdict, **self.kw)
        self.out_shape = tuple((out[0] if isinstance(out, tuple) else out).shape)
        return out

    def bwd(self, dout, dtokens=None):
        eng = self.engine
        key = tuple((n, tuple(p.shape)) for n, p in self.named)
        book = getattr(eng, "_gradbook", None)
        if book is None or getattr(eng, "_gradbook_key", None) != key:
            book = GradBook(self.named, (dout if dout is not None else dtokens).device)
            eng._gradbook, eng._gradbook_key = book, key
        # Gradients are returned as FRESH views of the flat book, so autograd's AccumulateGrad adopts them as p.grad
        # without a copy (it clones a tensor somebody else still references): p.grad then aliases the book, and an
        # in-place all-reduce of book.flat (bench.py, grad_ready_hook) lands in p.grad itself. A p.grad that still
        # aliases the book from the previous step (no zero_grad(set_to_none=True) in between: gradient accumulation)
        # is detached from it first, otherwise zeroing the book would wipe the accumulated value.
        for n, p in self.named:
            if p.grad is not None and p.grad.data_ptr() == book.views[n].data_ptr():
                p.grad = p.grad.clone()
        book.zero()
        if dout is None:                # only the token features were used downstream
            dout = torch.zeros(self.out_shape, dtype=torch.float32, device=dtokens.device)
        if dtokens is not None:
            eng.backward(self.pdict, book, dout.contiguous().float(), dtokens.contiguous().float())
        else:
            eng.backward(self.pdict, book, dout.contiguous().float())
        hook = getattr(eng, "grad_ready_hook", None)
        if hook is not None:        # e.g. start this tower's gradient all-reduce while the other tower still runs backward
            hook(book)
        return [book.fresh_view(n) if p.requires_grad else None for n, p in self.named]


def run_tower(engine, named_params, **fwd_kwargs):
    """Differentiable tower call. named_params: list of (name, Parameter) that the engine reads by name."""
    runner = TowerRunner(engine, named_params, fwd_kwargs)
    if not torch.is_grad_enabled() or not any(p.requires_grad for _, p in named_params):
        return engine.forward(runner.pdict, save=False, **fwd_kwargs)
    return _TowerFn.apply(runner, *[p for _, p in named_params])


# ---------------------------------------------------------------------------------------------- similarity + loss
class _SimMatrixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, eps):
        a = a.contiguous().float()
        b = b.contiguous().float()
        n, P = a.shape
        m = b.shape[0]
        ws = torch.empty(ops.sim_workspace_bytes(n, m, P), dtype=torch.uint8, device=a.device)
        sims = torch.empty(n, m, dtype=torch.float32, device=a.device)
        ops.sim_matrix_fwd(a, b, eps, sims, ws)
        ctx.ws, ctx.shape, ctx.eps = ws, (n, m, P), eps
        return sims

    @staticmethod
    def backward(ctx, dsims):
        n, m, P = ctx.shape
        da = torch.empty(n, P, dtype=torch.float32, device=dsims.device)
        db = torch.empty(m, P, dtype=torch.float32, device=dsims.device)
        ops.sim_matrix_bwd(dsims.contiguous().float(), ctx.eps, da, db, ctx.ws, n, m, P)
        return da, db, None


def sim_matrix(a, b, eps=1e-8):
    """Cosine-similarity matrix with eps-clamped norms (model/model.py:164-172). a: (n,P) text, b: (m,P) video."""
    return _SimMatrixFn.apply(a, b, eps)


class _NormSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, temperature):
        x = x.contiguous().float()
        assert x.shape[0] == x.shape[1], "NormSoftmaxLoss expects a square similarity matrix"
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        ops.norm_softmax_loss(x, temperature, loss, dx)
        ctx.save_for_backward(dx)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None


def norm_softmax_loss(x, temperature=0.05):
    return _NormSoftmaxFn.apply(x, temperature)


# ---------------------------------------------------------------------------------------------- all-gather
class AllGatherPairSlice(torch.autograd.Function):
    """The two embedding all-gathers of trainer_dist.py:159-160 as ONE collective: text and video rows are packed into a
    (B_loc, 2P) buffer, gathered once into (W * B_loc, 2P) and split again (the payload is 64 KB per rank,
    pure launch latency). Backward of each half = the local row slice, unreduced (trainer_dist.py:40-45)."""

    @staticmethod
    def forward(ctx, a, b, rank, world_size):
        import torch.distributed as dist
        ctx.rank, ctx.bs, ctx.pa = rank, a.shape[0], a.shape[1]
        if world_size == 1 or not (dist.is_available() and dist.is_initialized()):
            return a.contiguous().clone(), b.contiguous().clone()
        packed = torch.cat([a, b], dim=1).contiguous()
        out = torch.empty((world_size * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(out, packed)
        return out[:, :ctx.pa].contiguous(), out[:, ctx.pa:].contiguous()

    @staticmethod
    def backward(ctx, ga, gb):
        lo, hi = ctx.bs * ctx.rank, ctx.bs * (ctx.rank + 1)
        return (None if ga is None else ga[lo:hi]), (None if gb is None else gb[lo:hi]), None, None


class AllGatherSlice(torch.autograd.Function):
    """all_gather_into_tensor forward; backward = the local row slice, unreduced (trainer_dist.py:40-45)."""

    @staticmethod
    def forward(ctx, tensor, rank, world_size):
        import torch.distributed as dist
        ctx.rank, ctx.bs = rank, tensor.shape[0]
        t = tensor.contiguous()
        if world_size == 1 or not (dist.is_available() and dist.is_initialized()):
            return t.clone()
        out = torch.empty((world_size * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out

    @staticmethod
    def backward(ctx, grad):
        return grad[ctx.bs * ctx.rank: ctx.bs * (ctx.rank + 1)], None, None


# ---------------------------------------------------------------------------------------------- variant heads
class _LinearFn(torch.autograd.Function):
    """y = [relu](x) @ w^T + b on the tcgen05 GEMM (bf16 operands, fp32 accumulate): the projections the variant heads
    apply to token features (vid_proj over every patch token, oa_model_region_mem.py:142-145) and to the CLIP text-region
    embeddings (txt_proj_2 = ReLU -> Linear(512, 256), ibid. :70-72,118). x (rows, K) fp32, w (N, K), b (N)."""

    @staticmethod
    def forward(ctx, x, w, b, relu):
        rows, K = x.shape
        N = w.shape[0]
        x = x.contiguous().float()
        x16 = torch.empty(rows, K, dtype=torch.bfloat16, device=x.device)
        ops.cast_bf16(x, x16, relu=relu)
        w16 = torch.empty(N, K, dtype=torch.bfloat16, device=x.device)
        ops.cast_bf16(w.detach().contiguous().float(), w16)
        y = torch.empty(rows, N, dtype=torch.float32, device=x.device)
        ops.gemm(x16, w16, bias=None if b is None else b.detach().contiguous().float(), out_f32=y)
        ctx.save_for_backward(x if relu else None, x16, w16)
        ctx.relu, ctx.has_bias = relu, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x16, w16 = ctx.saved_tensors
        rows, K = x16.shape
        N = w16.shape[0]
        dy16 = torch.empty(rows, N, dtype=torch.bfloat16, device=dy.device)
        ops.cast_bf16(dy.contiguous().float(), dy16)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if ctx.relu:
                dx16 = torch.empty(rows, K, dtype=torch.bfloat16, device=dy.device)
                ops.gemm(dy16, w16, b_major=1, out_bf16=dx16)
                dx = torch.empty(rows, K, dtype=torch.float32, device=dy.device)
                ops.relu_bwd(x, dx16, dx, rows=rows, cols=K, ldx=K)
            else:
                dx = torch.empty(rows, K, dtype=torch.float32, device=dy.device)
                ops.gemm(dy16, w16, b_major=1, out_f32=dx)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(N, K, dtype=torch.float32, device=dy.device)
            ops.gemm(dy16, x16, a_major=1, b_major=1, out_f32=dw, accumulate=True)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, dtype=torch.float32, device=dy.device)
            ops.colsum_bf16(dy16, db)
        return dx, dw, db, None


def linear(x, w, b=None, relu=False):
    """(..., K) -> (..., N): nn.Linear (optionally behind a ReLU) on liboat, differentiable."""
    lead = x.shape[:-1]
    y = _LinearFn.apply(x.reshape(-1, x.shape[-1]), w, b, relu)
    return y.view(*lead, w.shape[0])


class _ObjectPatchAttnFn(torch.autograd.Function):
    """oat_object_patch_attn / _bwd: returns (weights (B,O,L), out (B,O,Cv))."""

    @staticmethod
    def forward(ctx, q, k, v, masks, mode):
        cf = lambda t: None if t is None else t.contiguous().float()
        q, k, v, masks = cf(q), cf(k), cf(v), cf(masks)
        weights, out = ops.object_patch_attention(q, k, v, mode=mode, masks=masks)
        ctx.mode = mode
        ctx.save_for_backward(q, k, v, weights)
        ctx.set_materialize_grads(False)
        return weights, out

    @staticmethod
    def backward(ctx, dweights, dout):
        q, k, v, weights = ctx.saved_tensors
        cf = lambda t: None if t is None else t.contiguous().float()
        if dweights is None and dout is None:
            return None, None, None, None, None
        if ctx.mode == "mask":
            dweights = None           # the masks are data
            if dout is None:
                return None, None, None, None, None
        dq, dk, dv = ops.object_patch_attention_bwd(q, k, v, weights, ctx.mode, dweights=cf(dweights), dout=cf(dout))
        return dq, dk, dv, None, None


def object_patch_attention(q, k, v=None, mode="softmax", masks=None):
    """Differentiable object -> patch attention: 'mask' pooling (oa_model_global_local.py:178), 'sigmoid' region
    similarity (oa_model_region_mem.py:147-151), 'softmax'. q (B,O,C), k (B,L,C), v (B,L,Cv), masks (B,O,L)."""
    return _ObjectPatchAttnFn.apply(q, k, v, masks, mode)


class _TokenPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls, tok, a, b):
        ctx.dims = tuple(tok.shape) + (a, b, cls is not None)
        return ops.token_pool(None if cls is None else cls.float(), tok.float(), a, b)

    @staticmethod
    def backward(ctx, dout):
        B, L, P, a, b, has_cls = ctx.dims
        dcls, dtok = ops.token_pool_bwd(dout.contiguous().float(), B, L, P, a, b, need_cls=has_cls)
        return dcls, dtok, None, None


def token_pool(cls, tok, a=0.5, b=0.5):
    """a * cls + b * mean(tok, dim=1): `(video_embeddings + torch.mean(video_region_feature, dim=1)) / 2`
    (oa_model_region_mem.py:119); cls=None, b=1 is `torch.mean(region_feat, dim=1)` (trainer_global_local.py:207)."""
    return _TokenPoolFn.apply(cls, tok, a, b)


class _BceSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, target, scale):
        loss, dp = ops.bce_sum(p.contiguous().float(), target.contiguous().float(), scale)
        ctx.save_for_backward(dp)
        ctx.shape = p.shape
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dp,) = ctx.saved_tensors
        return (dp * g).view(ctx.shape), None, None


def bce_sum(p, target, scale=1.0):
    """scale * nn.BCELoss(reduction='sum')(p, target) (trainer/trainer_region_mem.py:97,166)."""
    return _BceSumFn.apply(p, target, scale)
