"""Forward/backward schedules of the two towers over liboat kernels.

PyTorch supplies device buffers and the stream; every arithmetic step is a C-ABI launch (ops.py). The schedules
follow the reference modules line by line:
  video: SpaceTimeTransformer.forward_features (OATrans/model/video_transformer.py:303-351) with SpaceTimeBlock
         (:161-176, frozen-in-time residual wiring) and VarAttention (:99-135), then vid_proj (oa_model.py:129-133)
  text : HF DistilBertModel as called by FrozenInTime.compute_text (OATrans/model/oa_model.py:106-123)

Arithmetic contract (mirrored by oracle/oracle.py in bf16 mode): GEMM / attention operands are bf16, accumulation
fp32; residual stream, LayerNorm, softmax, logits, loss and every parameter gradient are fp32; activation gradients
that feed a GEMM are bf16.

Split-bf16 rows (forward only, OAT_SPLIT=0 turns it off): the logits depend DIRECTLY on the CLS row of the video
tower, on the (tiny) text tower and on the two projections; operand rounding there (2^-9 per product) is what moves
bf16 logits by ~1e-3 at the benchmark geometry, while the other F*n token rows only reach the CLS row through
attention averages over hundreds of keys. Those rows therefore take three bf16 MMAs per product instead of one:
x = hi + lo, w = hi + lo, x.w ~= hi.hi + hi.lo + lo.hi, laid out so that ONE K-concatenated GEMM does it
(activation rows [hi | hi | lo], weight rows [hi | lo | hi]); where the plain GEMM already produced hi.hi (+ bias +
residual) in fp32, only the correction [hi | lo] x [lo | hi] is accumulated on top. Cost: ~1 % of the step.
"""
import os

import torch

from . import ops

BF = torch.bfloat16
F32 = torch.float32
HEAD_DIM = 64
Q_SCALE = HEAD_DIM ** -0.5
OBJ_DIM = 2054          # 2048 ROI feature + 6 box numbers (base/base_dataset.py:593-650)
OBJ_PITCH = 2112        # padded to a multiple of 64 for TMA (16-byte row pitch) and whole k-blocks
ONES_PAD = 8            # activation rows are [x | 1 0 ... 0]: the weight-gradient GEMM then yields the bias gradient too
                        # (8: keeps 16-byte row pitches and fits the first half of a CTA pair's N = 16 instruction)
# Bias gradients of qkv / fc1 out of the weight-gradient GEMM's ones column instead of a column-sum kernel. Off by default:
# measured 60.3 vs 59.2 ms/step - the column sums already run on the side stream under tensor-bound chain kernels, while
# the extra (narrow) column block still streams every dY tile through shared memory a second time.
BIAS_VIA_WGRAD = os.environ.get("OAT_BIAS_VIA_WGRAD", "0") != "0"
SIDE_STREAM = os.environ.get("OAT_SIDE_STREAM", "1") != "0"   # weight gradients on a second stream (engine backward)
SPLIT = os.environ.get("OAT_SPLIT", "1") != "0"               # split-bf16 forward products on the CLS / text rows
# Replay the video tower's forward / backward as CUDA graphs. Off by default: measured 529 vs 540 pairs/s (graph / eager)
# on the benchmark step - the step is GPU-bound and the host already runs 3x ahead of the device, so there is no launch
# gap to reclaim; worth turning on when the host is the slower side (small batches, busy hosts).
GRAPH = os.environ.get("OAT_GRAPH", "0") != "0"
# delta = rowsum(dO * O) of the attention backward from the epilogue of the dO-producing GEMM (oat_gemm_bf16 act 4) instead
# of re-reading O inside the attention kernels. OAT_DELTA_EPI=0: the attention kernels compute it themselves.
DELTA_EPI = os.environ.get("OAT_DELTA_EPI", "1") != "0"
# qkv bias gradient of the video tower's attentions from a column sum over the dq columns only: the dk columns sum to zero
# (rows of dS sum to zero) and the dv columns sum to sum_i dO_i = db_proj . W_proj (rows of P sum to one) - a third of the
# column-sum traffic (see oat_vecmat_f32 in include/oat.h). OAT_QKV_BIAS_IDENTITY=0: column sums over all of dqkv.
QKV_BIAS_IDENTITY = os.environ.get("OAT_QKV_BIAS_IDENTITY", "1") != "0"
MAX_GRAPHS = 8


class _GraphedSchedule:
    """CUDA-graph replay of one schedule (the video tower's forward or backward: ~190 / ~280 launches whose arguments do
    not change from step to step because every activation lives in a name-keyed buffer). First call with a given key
    runs eagerly (lazy initialisation, buffer allocation), the second is captured, later ones replay. The key carries
    everything a captured kernel argument depends on: input addresses and shapes, the parameter storage, the flags.
    Kernels inside a replayed graph are still counted in ops.LAUNCHES (they do launch on the device)."""

    def __init__(self):
        self.entries = {}

    def run(self, key, fn):
        """fn() -> result (tensors living in static buffers). Returns (result, replayed)."""
        e = self.entries.get(key)
        if e is None:
            if len(self.entries) >= MAX_GRAPHS:          # too many distinct input addresses: stay eager for new ones
                return fn(), False
            self.entries[key] = {"graph": None}
            return fn(), False
        if e["graph"] is None:
            n0 = ops.LAUNCHES
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                e["result"] = fn()
            e["launches"] = ops.LAUNCHES - n0
            e["graph"] = g
            ops.LAUNCHES = n0                              # capture enqueued nothing; the replay below does the work
        e["graph"].replay()
        ops._count(e["launches"])
        ops.GRAPH_REPLAYS += 1
        return e["result"], True


class _Buffers:
    """Name-keyed cache of device buffers so that a training loop re-uses its activation storage every step. A buffer
    is re-allocated only when a request needs more elements than it holds and is otherwise viewed to the requested
    shape, so varying batch sizes / padded text lengths (tokenizer padding=True, trainer_dist.py:152) do not pile up
    one activation set per distinct shape."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.inited = {}

    def get(self, name, shape, dtype, zero=False, init=None):
        """init(t): run whenever the buffer is (re)allocated or viewed to a new shape (contents that the schedule
        never rewrites, e.g. the ones column of the extended activations)."""
        key = (name, dtype)
        numel = 1
        for d in shape:
            numel *= int(d)
        flat = self.bufs.get(key)
        if flat is None or flat.numel() < numel:
            flat = torch.empty(max(numel, 1), dtype=dtype, device=self.device)
            self.bufs[key] = flat
        t = flat[:numel].view(tuple(shape))
        if zero:
            t.zero_()
        if init is not None and (self.inited.get(key) != (flat.data_ptr(), tuple(shape))):
            init(t)
            self.inited[key] = (flat.data_ptr(), tuple(shape))
        return t


def _w16(bufs, name, w, rows=None, cols=None, pitch=None, plan=None, split=False):
    """bf16 operand copy of an fp32 weight (re-packed every step: the optimizer owns the fp32 master). With a CastPlan
    the copy is only registered; plan.run() performs every registered copy of the tower in one launch.
    split=True: the copy is [rows, 3*cols] = [hi | lo | hi] (hi = bf16(w), lo = bf16(w - hi)); its first `cols`
    columns are the ordinary bf16 operand."""
    w2 = w.detach().reshape(w.shape[0], -1)
    rows = w2.shape[0] if rows is None else rows
    cols = w2.shape[1] if cols is None else cols
    if split:
        assert plan is not None and pitch is None
        dst = bufs.get(name, (rows, 3 * cols), BF)
        plan.add(w2, dst, rows=rows, cols=cols, split=True)
        return dst
    dst = bufs.get(name, (rows, pitch or cols), BF)
    if plan is not None:
        plan.add(w2, dst, rows=rows, cols=cols)
    else:
        ops.cast_bf16(w2, dst, rows=rows, cols=cols)
    return dst


def _ones_column(D):
    def init(t):
        t.zero_()
        t[:, D] = 1.0
    return init


def _hi(w3):
    """[N, K] bf16 operand inside a split weight copy [N, 3K] = [hi | lo | hi]."""
    return w3[:, :w3.shape[1] // 3]


def _lo(w3):
    k = w3.shape[1] // 3
    return w3[:, k:2 * k]


def _spread(N, K, sms=148):
    """split-K factor that spreads a few-row accumulate GEMM over the whole chip: its cost is streaming the [N, K]
    weight, which a handful of output tiles cannot do fast."""
    tiles = max(1, (N + 255) // 256)
    return max(1, min((K + 63) // 64, sms // tiles))


def _lohi(w3):
    """[lo | hi]: against activation columns [hi | lo] this is the correction hi.lo + lo.hi."""
    return w3[:, w3.shape[1] // 3:]


class GradBook:
    """Flat fp32 gradient storage with one view per parameter (zeroed once per backward)."""

    def __init__(self, named_params, device):
        self.names = [n for n, _ in named_params]
        sizes = [p.numel() for _, p in named_params]
        self.flat = torch.zeros(sum(sizes), dtype=F32, device=device)
        self.views = {}
        off = 0
        for (n, p), s in zip(named_params, sizes):
            self.views[n] = self.flat[off:off + s].view(p.shape)
            off += s

    def zero(self):
        self.flat.zero_()

    def fresh_view(self, name):
        """A new tensor object over the same storage (see functional.TowerRunner.bwd)."""
        v = self.views[name]
        return self.flat.as_strided(v.shape, v.stride(), v.storage_offset())

    def __getitem__(self, name):
        return self.views[name]


# ==================================================================================================== video tower
class VideoEngine:
    """Space-time ViT over [CLS] + F x (N patch tokens + O object tokens), then the 768 -> P projection."""

    def __init__(self, device, heads=12, eps=1e-6, patch=16):
        self.device = device
        self.H = heads
        self.eps = eps
        self.patch = patch
        self.bufs = _Buffers(device)
        self.saved = None
        self._side = None
        self._plan = ops.CastPlan()
        self._graphs_fwd = _GraphedSchedule()
        self._graphs_bwd = _GraphedSchedule()
        self._saved_key = None

    # ------------------------------------------------------------------ forward
    def _fingerprint(self, p):
        return hash(tuple(v.data_ptr() for v in p.values()))

    def forward(self, p, video, objects=None, proj=("vid_proj.0.weight", "vid_proj.0.bias"), prefix="video_model.",
                save=True, tokens=None, region_layer=6):
        """See _forward. With OAT_GRAPH (default) a training forward of a repeating shape is replayed as a CUDA graph."""
        graphable = (GRAPH and save and tokens is None and ops.PROFILE is None and video.is_cuda
                     and getattr(self, "layer_grad_hook", None) is None and not torch.cuda.is_current_stream_capturing())
        if not graphable:
            return self._forward(p, video, objects, proj, prefix, save, tokens, region_layer)
        video = video.contiguous()
        objects = None if objects is None else objects.contiguous()
        key = ("fwd", video.data_ptr(), tuple(video.shape), None if objects is None else objects.data_ptr(),
               None if objects is None else tuple(objects.shape), proj, prefix, SPLIT, self._fingerprint(p))
        box = {}

        def fn():
            out = self._forward(p, video, objects, proj, prefix, True, None, region_layer)
            box["saved"] = self.saved
            return out, self.saved
        (out, saved), replayed = self._graphs_fwd.run(key, fn)
        self.saved = saved
        self._saved_key = key if replayed else None
        return out.clone() if replayed else out

    def _forward(self, p, video, objects=None, proj=("vid_proj.0.weight", "vid_proj.0.bias"), prefix="video_model.",
                 save=True, tokens=None, region_layer=6):
        """p: dict name -> fp32 parameter tensor. video fp32 (B,F,3,H,W); objects fp32 (B,F,O,2054) or None.
        Returns projected CLS embeddings fp32 (B, P).
        tokens="final": also returns the final LayerNorm of EVERY token row, fp32 (B, T, D) - forward_features'
          second result is its [:, 1:] (video_transformer.py:346-351).
        tokens="region": also returns region_norm(x) after `region_layer` blocks, fp32 (B, T, D)
          (oa_video_transformer_region.py:364-376: K = 6, region_feature = region_norm(x)[:, 1:])."""
        assert tokens in (None, "final", "region")
        bufs = self.bufs
        B, Fr, C, Hh, Ww = video.shape
        ps = self.patch
        N = (Hh // ps) * (Ww // ps)
        D = p[prefix + "cls_token"].shape[-1]
        H = self.H
        assert D == H * HEAD_DIM, "embed dim %d must be heads(%d) x 64" % (D, H)
        assert Fr <= p[prefix + "temporal_embed"].shape[1]        # video_transformer.py:73
        assert N + 1 == p[prefix + "pos_embed"].shape[1], "input resolution does not match pos_embed"
        O = 0 if objects is None else objects.shape[2]
        n = N + O
        T = 1 + Fr * n
        M = B * T
        depth = 0
        while (prefix + "blocks.%d.norm1.weight" % depth) in p:
            depth += 1
        video = video.contiguous()

        # --- bf16 operand copies of every weight of the tower: one launch
        plan = self._plan
        split = SPLIT
        W, W3 = {}, {}          # W: ordinary bf16 operands; W3: split copies [hi | lo | hi] (W[key] is then a view of W3[key])

        def weight(key, name, param):
            if split:
                W3[key] = _w16(bufs, name + ".s3", param, plan=plan, split=True)
                W[key] = _hi(W3[key])
            else:
                W[key] = _w16(bufs, name, param, plan=plan)

        W["patch"] = _w16(bufs, "w.patch", p[prefix + "patch_embed.proj.weight"], plan=plan)
        if O > 0:
            W["object"] = _w16(bufs, "w.object", p[prefix + "object_embed.weight"], pitch=OBJ_PITCH, plan=plan)
        for i in range(depth):
            b = "%sblocks.%d." % (prefix, i)
            for tag, aname in (("t", "timeattn"), ("s", "attn")):
                weight((i, tag, "qkv"), "w.%s.qkv.%d" % (tag, i), p[b + aname + ".qkv.weight"])
                weight((i, tag, "proj"), "w.%s.proj.%d" % (tag, i), p[b + aname + ".proj.weight"])
            weight((i, "fc1"), "w.fc1.%d" % i, p[b + "mlp.fc1.weight"])
            weight((i, "fc2"), "w.fc2.%d" % i, p[b + "mlp.fc2.weight"])
        if proj is not None:
            weight("vid_proj", "w.vid_proj", p[proj[0]])
        plan.run()

        def cls(t):
            """The B CLS rows (token 0 of every video) of an [M, w] token buffer, as a strided [B, w] view."""
            return t.view(B, T * t.shape[1])[:, :t.shape[1]]

        # --- patch / object embedding + token assembly (video_transformer.py:71-76, 303-325)
        K0 = C * ps * ps
        cols = bufs.get("cols", (B * Fr * N, K0), BF)
        ops.im2col_patches(video, cols, ps)
        wp = W["patch"]
        patch = bufs.get("patch", (B * Fr * N, D), F32)
        ops.gemm(cols, wp, bias=p[prefix + "patch_embed.proj.bias"], out_f32=patch)
        objemb = obj16 = None
        if O > 0:
            assert objects.shape[-1] == OBJ_DIM
            obj16 = bufs.get("obj16", (B * Fr * O, OBJ_PITCH), BF)
            ops.cast_bf16(objects.contiguous().reshape(-1, OBJ_DIM), obj16)
            wo = W["object"]
            objemb = bufs.get("objemb", (B * Fr * O, D), F32)
            ops.gemm(obj16, wo, bias=p[prefix + "object_embed.bias"], out_f32=objemb)
        type_embed = p.get(prefix + "token_type_embeddings.weight")
        xs = [bufs.get("x.%d" % i, (M, D), F32) for i in range(depth + 1)]
        ops.assemble_tokens(patch, objemb, p[prefix + "cls_token"], p[prefix + "pos_embed"],
                            p[prefix + "temporal_embed"], type_embed, xs[0], B, Fr, N, O, D)

        layers = []
        for i in range(depth):
            b = "%sblocks.%d." % (prefix, i)
            L = {}
            x = xs[i]

            def ln(tag, src, wname):
                """LayerNorm of every token row -> bf16 GEMM operand; the CLS rows also leave as [hi | hi | lo]."""
                hx = bufs.get("h%s.%d" % (tag, i), (M, D + ONES_PAD), BF, init=_ones_column(D))
                h = hx[:, :D]
                mean = bufs.get("mean%s.%d" % (tag, i), (M,), F32)
                rstd = bufs.get("rstd%s.%d" % (tag, i), (M,), F32)
                c3 = bufs.get("c3." + tag, (B, 3 * D), BF) if split else None
                ops.layernorm_fwd(src, p[b + wname + ".weight"], p[b + wname + ".bias"], self.eps, y_bf16=h,
                                  mean=mean, rstd=rstd, y_split=c3, split_period=T)
                L["h%sx" % tag] = hx
                return h, mean, rstd, c3

            def attention(tag, mode, h, c3, aname, resid, out):
                wqkv = W[(i, tag, "qkv")]
                qkv = bufs.get("qkv%s.%d" % (tag, i), (M, 3 * D), BF)
                ops.gemm(h, wqkv, bias=p[b + aname + ".qkv.bias"], scale_cols=D, scale=Q_SCALE, out_bf16=qkv)
                if split:       # CLS rows again, three-term product
                    ops.gemm(c3, W3[(i, tag, "qkv")], bias=p[b + aname + ".qkv.bias"], scale_cols=D, scale=Q_SCALE,
                             out_bf16=cls(qkv))
                a = bufs.get("a%s.%d" % (tag, i), (M, D), BF)
                lse = bufs.get("lse%s.%d" % (tag, i), (B * H * T,), F32)
                ws = bufs.get("attn_ws", (max(1, ops.attn_fwd_workspace_floats(ops.MODE_SPACE, B, H, Fr, n),
                                              ops.attn_fwd_workspace_floats(ops.MODE_TIME, B, H, Fr, n)),), F32)
                ops.attn_fwd(mode, B, T, H, Fr, n, qkv, a, lse, cls_ws=ws)
                wproj = W[(i, tag, "proj")]
                ops.gemm(a, wproj, bias=p[b + aname + ".proj.bias"], residual=resid, out_f32=out)
                if split:       # the attention output is bf16 already (lo = 0): add a . w_lo on the CLS rows
                    ops.gemm(cls(a), _lo(W3[(i, tag, "proj")]), out_f32=cls(out), accumulate=True,
                             split_k=_spread(D, D))
                return wqkv, wproj, qkv, a, lse

            # time attention on norm3(x); residual from x                    (video_transformer.py:164-165)
            L["h3"], L["m3"], L["r3"], c3 = ln("3", x, "norm3")
            tr = bufs.get("tr.%d" % i, (M, D), F32)
            L["wqkv_t"], L["wproj_t"], L["qkv_t"], L["a_t"], L["lse_t"] = attention("t", ops.MODE_TIME, L["h3"], c3,
                                                                                    "timeattn", x, tr)
            # space attention on norm1(time_residual); residual from x again  (:167-170)
            L["h1"], L["m1"], L["r1"], c3 = ln("1", tr, "norm1")
            sr = bufs.get("sr.%d" % i, (M, D), F32)
            L["wqkv_s"], L["wproj_s"], L["qkv_s"], L["a_s"], L["lse_s"] = attention("s", ops.MODE_SPACE, L["h1"], c3,
                                                                                    "attn", x, sr)
            # MLP on norm2(space_residual)                                   (:174, Mlp :45-51)
            L["h2"], L["m2"], L["r2"], c3 = ln("2", sr, "norm2")
            L["w1"] = W[(i, "fc1")]
            L["w2"] = W[(i, "fc2")]
            u = bufs.get("u.%d" % i, (M, 4 * D), BF)
            g = bufs.get("g.%d" % i, (M, 4 * D), BF)
            ops.gemm(L["h2"], L["w1"], bias=p[b + "mlp.fc1.bias"], act=ops.ACT_GELU, out_bf16=g, out2_bf16=u)
            g3 = None
            if split:           # CLS rows: three-term fc1, GELU kept in fp32 and split for fc2
                g32 = bufs.get("g32", (B, 4 * D), F32)
                ops.gemm(c3, W3[(i, "fc1")], bias=p[b + "mlp.fc1.bias"], act=ops.ACT_GELU, out_f32=g32,
                         out_bf16=cls(g), out2_bf16=cls(u))
                g3 = bufs.get("g3", (B, 12 * D), BF)
                ops.split3_bf16(g32, g3)
            ops.gemm(g, L["w2"], bias=p[b + "mlp.fc2.bias"], residual=sr, out_f32=xs[i + 1])
            if split:           # + g_hi . w_lo + g_lo . w_hi on the CLS rows
                ops.gemm(g3[:, 4 * D:], _lohi(W3[(i, "fc2")]), out_f32=cls(xs[i + 1]), accumulate=True,
                         split_k=_spread(D, 8 * D))
            L["tr"], L["sr"], L["u"], L["g"] = tr, sr, u, g
            layers.append(L)
            if tokens == "region" and i + 1 == region_layer:
                tok32, tmean, trstd = self._token_norm(p, prefix + "region_norm", xs[i + 1], M, D)

        if tokens == "region":
            assert 1 <= region_layer <= depth, "region_layer %d outside the %d blocks" % (region_layer, depth)
        if tokens == "final":
            tok32, tmean, trstd = self._token_norm(p, prefix + "norm", xs[depth], M, D)

        # final LayerNorm: only the CLS row is consumed (:346-351), then vid_proj (oa_model.py:131)
        cls32 = bufs.get("cls32", (B, D), F32)
        cls16 = bufs.get("cls16", (B, D), BF)
        mf = bufs.get("meanf", (B,), F32)
        rf = bufs.get("rstdf", (B,), F32)
        cls3 = bufs.get("cls3", (B, 3 * D), BF) if (split and proj is not None) else None
        ops.layernorm_fwd(xs[depth], p[prefix + "norm.weight"], p[prefix + "norm.bias"], self.eps, rows=B, ldx=T * D,
                          y_bf16=cls16, y_f32=cls32, mean=mf, rstd=rf, y_split=cls3, split_period=1)
        out = cls32.clone() if proj is None else None
        wv = None
        if proj is not None:
            wv = W["vid_proj"]
            P = p[proj[0]].shape[0]
            out = torch.empty((B, P), dtype=F32, device=self.device)
            if split:
                ops.gemm(cls3, W3["vid_proj"], bias=p[proj[1]], out_f32=out)
            else:
                ops.gemm(cls16, wv, bias=p[proj[1]], out_f32=out)
        if save:
            self.saved = dict(B=B, Fr=Fr, N=N, O=O, n=n, T=T, M=M, D=D, depth=depth, xs=xs, layers=layers, cols=cols,
                              obj16=obj16, cls16=cls16, mf=mf, rf=rf, wv=wv, proj=proj, prefix=prefix,
                              has_type=type_embed is not None, tokens=tokens, region_layer=region_layer,
                              tmean=tmean if tokens else None, trstd=trstd if tokens else None)
        if tokens is not None:
            return out, tok32.view(B, T, D)
        return out

    def _token_norm(self, p, wname, x, M, D):
        """LayerNorm of every token row (fp32 result): the patch / region features the variant heads consume."""
        tok32 = torch.empty((M, D), dtype=F32, device=self.device)      # returned to autograd: not a recycled buffer
        mean = self.bufs.get("tok.mean", (M,), F32)
        rstd = self.bufs.get("tok.rstd", (M,), F32)
        ops.layernorm_fwd(x, p[wname + ".weight"], p[wname + ".bias"], self.eps, y_f32=tok32, mean=mean, rstd=rstd)
        return tok32, mean, rstd

    # ------------------------------------------------------------------ backward
    def backward(self, p, grads, dout, dtokens=None):
        """See _backward. Replayed as a CUDA graph when the forward it belongs to was (same shapes, same buffers)."""
        graphable = (GRAPH and dtokens is None and ops.PROFILE is None and dout.is_cuda and self._saved_key is not None
                     and hasattr(grads, "flat") and getattr(self, "layer_grad_hook", None) is None
                     and not torch.cuda.is_current_stream_capturing())
        if not graphable:
            return self._backward(p, grads, dout, dtokens)
        static = self.bufs.get("graph.dout", tuple(dout.shape), F32)
        static.copy_(dout)
        # every forward graph of one shape shares the activation buffers, so one backward graph serves them all
        key = ("bwd", self._saved_key[2], self._saved_key[4], self._saved_key[5:], grads.flat.data_ptr(), tuple(dout.shape),
               SIDE_STREAM)
        saved = self.saved
        self._graphs_bwd.run(key, lambda: self._backward(p, grads, static, None, saved))

    def _backward(self, p, grads, dout, dtokens=None, saved=None):
        """dout: fp32 (B, P) gradient of the projected embeddings; dtokens: fp32 (B, T, D) gradient of the token
        features returned with tokens="final" / "region" (None: unused). Fills `grads` (GradBook-like: name -> fp32
        view, pre-zeroed) with every parameter gradient."""
        S = self.saved if saved is None else saved
        assert S is not None, "backward without a saved forward"
        bufs = self.bufs
        B, Fr, N, O, n, T, M, D, depth = (S[k] for k in ("B", "Fr", "N", "O", "n", "T", "M", "D", "depth"))
        H = self.H
        prefix = S["prefix"]
        xs, layers = S["xs"], S["layers"]

        # Weight-gradient GEMMs and bias column sums have no consumer inside the backward chain: they run on a second
        # stream so that the HBM-bound kernels of the chain (LayerNorm backward, attention backward) overlap with them.
        # Gradient activations they read rotate over two buffer sets (layer parity); the chain waits for the side
        # stream's layer i+2 before overwriting set i%2.
        use_side = SIDE_STREAM and torch.cuda.is_available()
        main = torch.cuda.current_stream(self.device) if use_side else None
        if use_side:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            side = self._side
            side.wait_stream(main)              # gradient book zeroed, forward finished
        side_done = {}

        def wgrad(dy16, act16, name, bias=True, ext=None, proj=None):
            # bias=False: the bias gradient was already reduced (fp32) by the LayerNorm-backward kernel that produced dy.
            # ext: the activation rows extended by a ones column ([x | 1 0 .. 0]); dY^T . ext then carries the bias
            # gradient in its last column block (an N = 16 MMA) instead of a second pass over dY (oat_unpack_wgrad).
            def run():
                if bias and ext is not None and BIAS_VIA_WGRAD:
                    n_out, k_in = dy16.shape[1], act16.shape[1]
                    scratch = bufs.get("wgrad.scratch", (4 * D, D + ONES_PAD), F32, init=lambda t: t.zero_())
                    scratch = scratch[:n_out]
                    ops.gemm(dy16, ext, a_major=1, b_major=1, out_f32=scratch, accumulate=True)
                    ops.unpack_wgrad(scratch, k_in, grads[name + ".weight"].view(n_out, k_in), grads[name + ".bias"])
                    return
                ops.gemm(dy16, act16, a_major=1, b_major=1, out_f32=grads[name + ".weight"].view(dy16.shape[1], -1),
                         accumulate=True)
                if bias and proj is not None and QKV_BIAS_IDENTITY:
                    # proj: name of the output projection behind this attention; its bias gradient (sum of dY_proj over the
                    # token rows, from the LayerNorm backward's dxsum) is complete before the attention backward runs
                    db = grads[name + ".bias"]
                    ops.colsum_bf16(dy16[:, :D], db[:D])                                 # q: a real column sum
                    ops.vecmat_f32(grads[proj + ".bias"], p[proj + ".weight"], db[2 * D:])   # v; k stays exactly zero
                elif bias:
                    ops.colsum_bf16(dy16, grads[name + ".bias"])
            if not use_side:
                return run()
            ev = torch.cuda.Event()
            ev.record(main)                      # dy16 is complete on the chain
            side.wait_event(ev)
            with torch.cuda.stream(side):
                run()

        # bf16 d(block output): read by the fc2 weight gradient one layer later, so three buffers rotate
        dy = bufs.get("dy.a", (M, D), F32, zero=True)         # d(block output), non-zero in the CLS rows only
        dy16 = bufs.get("dy16.%d" % (depth % 3), (M, D), BF, zero=True)
        dcls16 = bufs.get("dcls16", (B, D), BF)
        if S["proj"] is not None:
            wname, bname = S["proj"]
            P = dout.shape[1]
            d16 = bufs.get("dout16", (B, P), BF)
            ops.cast_bf16(dout.contiguous(), d16)
            ops.gemm(d16, S["wv"], b_major=1, out_bf16=dcls16)
            ops.gemm(d16, S["cls16"], a_major=1, b_major=1, out_f32=grads[wname], accumulate=True)
            ops.colsum_bf16(d16, grads[bname])
        else:
            ops.cast_bf16(dout.contiguous(), dcls16)
        # LayerNorm backward is linear in dy, so the gradient arriving through the token features is a second pass over
        # ALL rows (final norm, or region_norm when it sits on the last block) added onto the CLS-row pass; only the
        # last pass accumulates the column sums of the finished dx (= fc2 bias gradient of the last block).
        tok_mode = S["tokens"] if dtokens is not None else None
        region_at = S["region_layer"] if tok_mode == "region" else -1
        tok_here = tok_mode == "final" or region_at == depth
        fc2b_last = grads["%sblocks.%d.mlp.fc2.bias" % (prefix, depth - 1)] if depth > 0 else None
        ops.layernorm_bwd(xs[depth], S["mf"], S["rf"], p[prefix + "norm.weight"], dy_bf16=dcls16, rows=B, ldx=T * D,
                          dx=dy, dx_bf16=dy16, lddx=T * D, lddxb=T * D, dgamma=grads[prefix + "norm.weight"],
                          dbeta=grads[prefix + "norm.bias"], dxsum=None if tok_here else fc2b_last)
        if dtokens is not None:
            dtokens = dtokens.contiguous().view(M, D)
        if tok_here:
            wn = prefix + ("norm" if tok_mode == "final" else "region_norm")
            ops.layernorm_bwd(xs[depth], S["tmean"], S["trstd"], p[wn + ".weight"], dy_f32=dtokens, add1=dy, dx=dy,
                              dx_bf16=dy16, dgamma=grads[wn + ".weight"], dbeta=grads[wn + ".bias"], dxsum=fc2b_last)

        nset = 2 if use_side else 1
        du = [bufs.get("du.%d" % k, (M, 4 * D), BF) for k in range(nset)]
        dqkv_s = [bufs.get("dqkv_s.%d" % k, (M, 3 * D), BF) for k in range(nset)]
        dqkv_t = [bufs.get("dqkv_t.%d" % k, (M, 3 * D), BF) for k in range(nset)] if use_side else dqkv_s
        dsr16 = [bufs.get("dsr16.%d" % k, (M, D), BF) for k in range(nset)]
        dtr16 = [bufs.get("dtr16.%d" % k, (M, D), BF) for k in range(nset)]
        dh = bufs.get("dh", (M, D), BF)
        da = bufs.get("da", (M, D), BF)
        dsr = bufs.get("dsr", (M, D), F32)
        dtr = bufs.get("dtr", (M, D), F32)
        acc = bufs.get("cls_acc", (B * H * 3 * HEAD_DIM,), F32)
        dyb = bufs.get("dy.b", (M, D), F32)
        # delta = rowsum(dO * O) per head comes out of the epilogue of the GEMM that produces dO (the dgrad of the output
        # projection), so the attention backward kernels never read O
        delta = bufs.get("attn_delta", (H, M), F32) if (DELTA_EPI and D % 256 == 0) else None

        def dgrad_proj(dy_, w_, o_):
            if delta is None:
                ops.gemm(dy_, w_, b_major=1, out_bf16=da)
            else:
                ops.gemm(dy_, w_, b_major=1, act=ops.ACT_ROWDOT, aux=o_, rowdot=delta, out_bf16=da)

        for i in reversed(range(depth)):
            b = "%sblocks.%d." % (prefix, i)
            L = layers[i]
            k = i % nset
            dy16b = bufs.get("dy16.%d" % (i % 3), (M, D), BF)
            if use_side and (i + 2) in side_done:
                main.wait_event(side_done[i + 2])            # buffer set k is free again
            # ---- x_out = sr + fc2(gelu(fc1(norm2(sr))))
            ops.gemm(dy16, L["w2"], b_major=1, act=ops.ACT_GELU_BWD, aux=L["u"], out_bf16=du[k])
            wgrad(dy16, L["g"], b + "mlp.fc2", bias=False)
            ops.gemm(du[k], L["w1"], b_major=1, out_bf16=dh)
            wgrad(du[k], L["h2"], b + "mlp.fc1", ext=L["h2x"])
            ops.layernorm_bwd(L["sr"], L["m2"], L["r2"], p[b + "norm2.weight"], dy_bf16=dh, add1=dy, dx=dsr,
                              dx_bf16=dsr16[k], dgamma=grads[b + "norm2.weight"], dbeta=grads[b + "norm2.bias"],
                              dxsum=grads[b + "attn.proj.bias"])
            # ---- sr = x + proj_s(space_attn(qkv_s(norm1(tr))))
            dgrad_proj(dsr16[k], L["wproj_s"], L["a_s"])
            wgrad(dsr16[k], L["a_s"], b + "attn.proj", bias=False)
            ops.attn_bwd(ops.MODE_SPACE, B, T, H, Fr, n, L["qkv_s"], L["a_s"], L["lse_s"], da, dqkv_s[k], Q_SCALE, acc,
                         delta=delta)
            ops.gemm(dqkv_s[k], L["wqkv_s"], b_major=1, out_bf16=dh)
            wgrad(dqkv_s[k], L["h1"], b + "attn.qkv", ext=L["h1x"], proj=b + "attn.proj")
            ops.layernorm_bwd(L["tr"], L["m1"], L["r1"], p[b + "norm1.weight"], dy_bf16=dh, dx=dtr, dx_bf16=dtr16[k],
                              dgamma=grads[b + "norm1.weight"], dbeta=grads[b + "norm1.bias"],
                              dxsum=grads[b + "timeattn.proj.bias"])
            # ---- tr = x + proj_t(time_attn(qkv_t(norm3(x))))
            dgrad_proj(dtr16[k], L["wproj_t"], L["a_t"])
            wgrad(dtr16[k], L["a_t"], b + "timeattn.proj", bias=False)
            ops.attn_bwd(ops.MODE_TIME, B, T, H, Fr, n, L["qkv_t"], L["a_t"], L["lse_t"], da, dqkv_t[k], Q_SCALE, acc,
                         delta=delta)
            ops.gemm(dqkv_t[k], L["wqkv_t"], b_major=1, out_bf16=dh)
            wgrad(dqkv_t[k], L["h3"], b + "timeattn.qkv", ext=L["h3x"], proj=b + "timeattn.proj")
            if i == region_at:      # x_i also fed region_norm: dsr += region_norm'(dtokens), in place, before the sum below
                wn = prefix + "region_norm"
                ops.layernorm_bwd(xs[i], S["tmean"], S["trstd"], p[wn + ".weight"], dy_f32=dtokens, add1=dsr, dx=dsr,
                                  dgamma=grads[wn + ".weight"], dbeta=grads[wn + ".bias"])
            # ---- dx = dsr (space skip) + dtr (time skip) + norm3'(dh)
            ops.layernorm_bwd(xs[i], L["m3"], L["r3"], p[b + "norm3.weight"], dy_bf16=dh, add1=dsr, add2=dtr, dx=dyb,
                              dx_bf16=dy16b, dgamma=grads[b + "norm3.weight"], dbeta=grads[b + "norm3.bias"],
                              dxsum=grads["%sblocks.%d.mlp.fc2.bias" % (prefix, i - 1)] if i > 0 else None)
            if use_side:
                ev = torch.cuda.Event()
                ev.record(side)
                side_done[i] = ev
            hook = getattr(self, "layer_grad_hook", None)
            if hook is not None:        # every gradient of block i has been enqueued (fc2.bias came from block i+1's LN3)
                if use_side:
                    # ... the weight gradients on the side stream, the LayerNorm / bias gradients on the chain: the hook
                    # runs on the side stream once it has also seen the chain up to here, so whatever it launches (a
                    # gradient all-reduce) is ordered after both and the chain itself never waits
                    ev = torch.cuda.Event()
                    ev.record(main)
                    side.wait_event(ev)
                    with torch.cuda.stream(side):
                        hook(grads, "%sblocks.%d." % (prefix, i))
                else:
                    hook(grads, "%sblocks.%d." % (prefix, i))
            dy, dyb = dyb, dy
            dy16 = dy16b

        # ---- token assembly / embeddings
        dpatch = bufs.get("dpatch", (B * Fr * N, D), BF)
        dobj = bufs.get("dobj", (B * Fr * O, D), BF) if O > 0 else None
        ops.assemble_tokens_bwd(dy, dpatch, dobj, grads[prefix + "cls_token"], grads[prefix + "pos_embed"],
                                grads[prefix + "temporal_embed"],
                                grads[prefix + "token_type_embeddings.weight"] if S["has_type"] else None,
                                B, Fr, N, O, D)
        wgrad(dpatch, S["cols"], prefix + "patch_embed.proj")
        if O > 0:
            dwo = bufs.get("dw.object", (D, OBJ_PITCH), F32, zero=True)
            ops.gemm(dobj, S["obj16"], a_major=1, b_major=1, out_f32=dwo, accumulate=True)
            grads[prefix + "object_embed.weight"].copy_(dwo[:, :OBJ_DIM])
            ops.colsum_bf16(dobj, grads[prefix + "object_embed.bias"])
        if use_side:
            main.wait_stream(side)


# ==================================================================================================== text tower
class TextEngine:
    """DistilBERT (post-LN, 6 layers) -> CLS -> ReLU -> Linear(768, P)."""

    def __init__(self, device, heads=12, eps=1e-12):
        self.device = device
        self.H = heads
        self.eps = eps
        self.bufs = _Buffers(device)
        self.saved = None
        self._plan = ops.CastPlan()

    def forward(self, p, input_ids, attention_mask=None, proj=("txt_proj.1.weight", "txt_proj.1.bias"),
                prefix="text_model.", save=True, dropout=None):
        """dropout = {"p": hidden dropout, "p_attn": attention dropout, "seed": int} reproduces DistilBERT's training
        mode (the reference keeps text_model.train(), model/oa_model.py:28): after the embedding LayerNorm (site 0), on
        the softmax weights of layer i (site 1 + 3 i) and on the FFN output before the residual add (site 2 + 3 i).
        The masks are Philox draws of (seed, site, element): nothing is stored, backward re-draws them."""
        bufs = self.bufs
        pd = float(dropout["p"]) if dropout else 0.0
        pa = float(dropout["p_attn"]) if dropout else 0.0
        seed = int(dropout["seed"]) if dropout else 0
        B, Lq = input_ids.shape
        H = self.H
        word = p[prefix + "embeddings.word_embeddings.weight"]
        pos = p[prefix + "embeddings.position_embeddings.weight"]
        D = word.shape[1]
        assert D == H * HEAD_DIM
        assert Lq <= pos.shape[0]
        M = B * Lq
        layers_n = 0
        while (prefix + "transformer.layer.%d.sa_layer_norm.weight" % layers_n) in p:
            layers_n += 1
        ids = input_ids.contiguous().to(torch.int64)
        key_mask = None
        if attention_mask is not None:
            key_mask = attention_mask.to(torch.int32).contiguous().view(-1)

        # --- bf16 operand copies of every weight (q / k / v packed into one [3D, D] operand, biases likewise): one launch
        # With SPLIT every weight copy is [hi | lo | hi] and every activation operand [hi | hi | lo]: the whole (tiny)
        # text tower runs its forward products in three-term split-bf16 arithmetic (module docstring).
        plan = self._plan
        split = SPLIT
        S3 = 3 if split else 1
        W = {}
        for i in range(layers_n):
            b = "%stransformer.layer.%d." % (prefix, i)
            wqkv = bufs.get("w.qkv.%d.%d" % (i, S3), (3 * D, S3 * D), BF)
            bqkv = bufs.get("b.qkv.%d" % i, (3 * D,), F32)
            for j, nm in enumerate(("q_lin", "k_lin", "v_lin")):
                plan.add(p[b + "attention.%s.weight" % nm], wqkv[j * D:(j + 1) * D], split=split)
                plan.add(p[b + "attention.%s.bias" % nm].detach().view(1, D), bqkv[j * D:(j + 1) * D].view(1, D))
            W[(i, "qkv")], W[(i, "bqkv")] = wqkv, bqkv
            W[(i, "o")] = _w16(bufs, "w.o.%d" % i, p[b + "attention.out_lin.weight"], plan=plan, split=split)
            W[(i, "l1")] = _w16(bufs, "w.l1.%d" % i, p[b + "ffn.lin1.weight"], plan=plan, split=split)
            W[(i, "l2")] = _w16(bufs, "w.l2.%d" % i, p[b + "ffn.lin2.weight"], plan=plan, split=split)
        if proj is not None:
            W["txt_proj"] = _w16(bufs, "w.txt_proj", p[proj[0]], plan=plan, split=split)
        plan.run()
        hi = _hi if split else (lambda w: w)

        emb = bufs.get("emb", (M, D), F32)
        ops.text_embed(ids, word, pos, emb, Lq)

        def ln(tag, src, wname, drop_site=None):
            """-> (GEMM operand [M, S3*D], its plain bf16 part [M, D], fp32 value, mean, rstd); with drop_site the
            normalised rows pass through dropout before they become the residual stream and the GEMM operand."""
            y32 = bufs.get("y32." + tag, (M, D), F32)
            mean = bufs.get("mean." + tag, (M,), F32)
            rstd = bufs.get("rstd." + tag, (M,), F32)
            dropped = drop_site is not None and pd > 0.0
            y3 = bufs.get("y3." + tag, (M, 3 * D), BF) if split else None
            y16 = None if split else bufs.get("y16." + tag, (M, D), BF)
            if dropped:
                ops.layernorm_fwd(src, p[wname + ".weight"], p[wname + ".bias"], self.eps, y_f32=y32, mean=mean, rstd=rstd)
                ops.dropout_fwd(y32, pd, seed, drop_site, out=y32, out_bf16=y16, out_split3=y3)
            else:
                ops.layernorm_fwd(src, p[wname + ".weight"], p[wname + ".bias"], self.eps, y_bf16=y16, y_f32=y32,
                                  mean=mean, rstd=rstd, y_split=y3, split_period=1)
            if split:
                return y3, y3[:, :D], y32, mean, rstd
            return y16, y16, y32, mean, rstd

        xop, x16, x32, m0, r0 = ln("emb", emb, prefix + "embeddings.LayerNorm", drop_site=0)
        layers = []
        for i in range(layers_n):
            b = "%stransformer.layer.%d." % (prefix, i)
            L = {"x16": x16, "x32": x32}
            wqkv, bqkv = W[(i, "qkv")], W[(i, "bqkv")]
            qkv = bufs.get("qkv.%d" % i, (M, 3 * D), BF)
            ops.gemm(xop, wqkv, bias=bqkv, scale_cols=D, scale=Q_SCALE, out_bf16=qkv)
            ctx = bufs.get("ctx.%d" % i, (M, D), BF)
            lse = bufs.get("lse.%d" % i, (B * H * Lq,), F32)
            ops.attn_fwd(ops.MODE_PLAIN, B, Lq, H, 0, 0, qkv, ctx, lse, key_mask,
                         dropout=(pa, seed, 1 + 3 * i) if pa > 0.0 else None)
            wo = W[(i, "o")]
            sa_sum = bufs.get("sa_sum.%d" % i, (M, D), F32)
            ops.gemm(ctx, hi(wo), bias=p[b + "attention.out_lin.bias"], residual=x32, out_f32=sa_sum)
            if split:           # ctx is bf16 already: + ctx . w_lo
                ops.gemm(ctx, _lo(wo), out_f32=sa_sum, accumulate=True)
            yop, y16, y32, m1, r1 = ln("sa.%d" % i, sa_sum, b + "sa_layer_norm")
            w1, w2 = W[(i, "l1")], W[(i, "l2")]
            Hd = w1.shape[0]
            u = bufs.get("u.%d" % i, (M, Hd), BF)
            g = bufs.get("g.%d" % i, (M, Hd), BF)
            ffn_sum = bufs.get("ffn_sum.%d" % i, (M, D), F32)
            # FFN.ff_chunk: lin2 output -> dropout -> + residual; without dropout the residual add rides in the GEMM epilogue
            ffn_out = bufs.get("ffn_tmp", (M, D), F32) if pd > 0.0 else ffn_sum
            ffn_res = None if pd > 0.0 else y32
            if split:
                g32 = bufs.get("g32", (M, Hd), F32)
                ops.gemm(yop, w1, bias=p[b + "ffn.lin1.bias"], act=ops.ACT_GELU, out_f32=g32, out_bf16=g, out2_bf16=u)
                g3 = bufs.get("g3", (M, 3 * Hd), BF)
                ops.split3_bf16(g32, g3)
                ops.gemm(g3, w2, bias=p[b + "ffn.lin2.bias"], residual=ffn_res, out_f32=ffn_out)
            else:
                ops.gemm(y16, w1, bias=p[b + "ffn.lin1.bias"], act=ops.ACT_GELU, out_bf16=g, out2_bf16=u)
                ops.gemm(g, w2, bias=p[b + "ffn.lin2.bias"], residual=ffn_res, out_f32=ffn_out)
            if pd > 0.0:
                ops.dropout_fwd(ffn_out, pd, seed, 2 + 3 * i, residual=y32, out=ffn_sum)
            xop, x16, x32, m2, r2 = ln("out.%d" % i, ffn_sum, b + "output_layer_norm")
            L.update(wqkv=hi(wqkv), qkv=qkv, ctx=ctx, lse=lse, wo=hi(wo), sa_sum=sa_sum, y16=y16, m1=m1, r1=r1,
                     w1=hi(w1), w2=hi(w2), u=u, g=g, ffn_sum=ffn_sum, m2=m2, r2=r2)
            layers.append(L)

        # last_hidden_state[:, 0] -> ReLU -> Linear (oa_model.py:113-121, 68-69)
        out = x32.view(B, Lq, D)[:, 0].clone() if proj is None else None
        r16 = wt = None
        if proj is not None:
            wt = hi(W["txt_proj"])
            out = torch.empty((B, p[proj[0]].shape[0]), dtype=F32, device=self.device)
            if split:
                r3 = bufs.get("relu3", (B, 3 * D), BF)
                ops.split3_bf16(x32, r3, rows=B, cols=D, lds=Lq * D, relu=True)
                r16 = r3[:, :D]
                ops.gemm(r3, W["txt_proj"], bias=p[proj[1]], out_f32=out)
            else:
                r16 = bufs.get("relu16", (B, D), BF)
                ops.cast_bf16(x32, r16, rows=B, cols=D, lds=Lq * D, relu=True)
                ops.gemm(r16, wt, bias=p[proj[1]], out_f32=out)
        if save:
            self.saved = dict(B=B, Lq=Lq, M=M, D=D, ids=ids, key_mask=key_mask, emb=emb, m0=m0, r0=r0, layers=layers,
                              last32=x32, r16=r16, wt=wt, proj=proj, prefix=prefix, drop=(pd, pa, seed))
        return out

    def backward(self, p, grads, dout):
        S = self.saved
        assert S is not None, "backward without a saved forward"
        bufs = self.bufs
        B, Lq, M, D = S["B"], S["Lq"], S["M"], S["D"]
        H = self.H
        prefix = S["prefix"]
        layers = S["layers"]
        pd, pa, seed = S["drop"]

        def wgrad_into(dy16, act16, wview, bview):
            ops.gemm(dy16, act16, a_major=1, b_major=1, out_f32=wview, accumulate=True)
            if bview is not None:       # None: already reduced in fp32 by the producing LayerNorm-backward kernel
                ops.colsum_bf16(dy16, bview)

        dX = bufs.get("dX", (M, D), F32, zero=True)      # fp32 gradient w.r.t. the layer output (CLS rows only at first)
        if S["proj"] is not None:
            wname, bname = S["proj"]
            d16 = bufs.get("dout16", (B, dout.shape[1]), BF)
            ops.cast_bf16(dout.contiguous(), d16)
            dr16 = bufs.get("dr16", (B, D), BF)
            ops.gemm(d16, S["wt"], b_major=1, out_bf16=dr16)
            wgrad_into(d16, S["r16"], grads[wname], grads[bname])
            ops.relu_bwd(S["last32"], dr16, dX, rows=B, cols=D, ldx=Lq * D, lddx=Lq * D)   # straight onto the CLS rows
        else:
            dX.view(B, Lq, D)[:, 0].copy_(dout)
        dX16 = None                                       # optional bf16 part of the same gradient

        dffn = bufs.get("dffn", (M, D), F32)
        dffn16 = bufs.get("dffn16", (M, D), BF)
        dsa = bufs.get("dsa", (M, D), F32)
        dsa16 = bufs.get("dsa16", (M, D), BF)
        dya = bufs.get("dya16", (M, D), BF)
        dxa = bufs.get("dxa16", (M, D), BF)
        dctx = bufs.get("dctx", (M, D), BF)
        dqkv = bufs.get("dqkv", (M, 3 * D), BF)
        dwqkv = bufs.get("dwqkv", (3 * D, D), F32)
        dbqkv = bufs.get("dbqkv", (3 * D,), F32)

        for i in reversed(range(len(layers))):
            b = "%stransformer.layer.%d." % (prefix, i)
            L = layers[i]
            Hd = L["u"].shape[1]
            du = bufs.get("du", (M, Hd), BF)
            # x_out = LN(ffn_sum),  ffn_sum = lin2(gelu(lin1(y))) + y
            if pd > 0.0:    # the FFN branch sees the gradient through its dropout mask; the skip (dffn) does not
                ops.layernorm_bwd(L["ffn_sum"], L["m2"], L["r2"], p[b + "output_layer_norm.weight"], dy_bf16=dX16,
                                  dy_f32=dX, dx=dffn, dgamma=grads[b + "output_layer_norm.weight"],
                                  dbeta=grads[b + "output_layer_norm.bias"])
                ops.dropout_bwd(pd, seed, 2 + 3 * i, dy_f32=dffn, dx_bf16=dffn16)
                ops.colsum_bf16(dffn16, grads[b + "ffn.lin2.bias"])
            else:
                ops.layernorm_bwd(L["ffn_sum"], L["m2"], L["r2"], p[b + "output_layer_norm.weight"], dy_bf16=dX16,
                                  dy_f32=dX, dx=dffn, dx_bf16=dffn16, dgamma=grads[b + "output_layer_norm.weight"],
                                  dbeta=grads[b + "output_layer_norm.bias"], dxsum=grads[b + "ffn.lin2.bias"])
            ops.gemm(dffn16, L["w2"], b_major=1, act=ops.ACT_GELU_BWD, aux=L["u"], out_bf16=du)
            wgrad_into(dffn16, L["g"], grads[b + "ffn.lin2.weight"], None)
            ops.gemm(du, L["w1"], b_major=1, out_bf16=dya)
            wgrad_into(du, L["y16"], grads[b + "ffn.lin1.weight"], grads[b + "ffn.lin1.bias"])
            # y = LN(sa_sum),  sa_sum = out_lin(attn) + x ; dy = dffn (skip) + dya (through lin1)
            ops.layernorm_bwd(L["sa_sum"], L["m1"], L["r1"], p[b + "sa_layer_norm.weight"], dy_bf16=dya, dy_f32=dffn,
                              dx=dsa, dx_bf16=dsa16, dgamma=grads[b + "sa_layer_norm.weight"],
                              dbeta=grads[b + "sa_layer_norm.bias"], dxsum=grads[b + "attention.out_lin.bias"])
            ops.gemm(dsa16, L["wo"], b_major=1, out_bf16=dctx)
            wgrad_into(dsa16, L["ctx"], grads[b + "attention.out_lin.weight"], None)
            ops.attn_bwd(ops.MODE_PLAIN, B, Lq, H, 0, 0, L["qkv"], L["ctx"], L["lse"], dctx, dqkv, Q_SCALE, None,
                         S["key_mask"], dropout=(pa, seed, 1 + 3 * i) if pa > 0.0 else None)
            ops.gemm(dqkv, L["wqkv"], b_major=1, out_bf16=dxa)
            dwqkv.zero_()
            dbqkv.zero_()
            wgrad_into(dqkv, L["x16"], dwqkv, dbqkv)
            for j, nm in enumerate(("q_lin", "k_lin", "v_lin")):
                grads[b + "attention.%s.weight" % nm].copy_(dwqkv[j * D:(j + 1) * D])
                grads[b + "attention.%s.bias" % nm].copy_(dbqkv[j * D:(j + 1) * D])
            # gradient w.r.t. the layer input x: dsa (skip, fp32) + dxa (through q/k/v, bf16)
            dX, dsa = dsa, dX
            dX16 = dxa      # consumed by the first LayerNorm backward of the next iteration, rewritten after it

        demb = dffn
        if pd > 0.0:        # embedding dropout sits between the LayerNorm and everything that consumed its output
            ops.dropout_bwd(pd, seed, 0, dy_f32=dX, dy_bf16=dX16, dx_f32=dsa)
            dX, dX16 = dsa, None
        ops.layernorm_bwd(S["emb"], S["m0"], S["r0"], p[prefix + "embeddings.LayerNorm.weight"], dy_bf16=dX16,
                          dy_f32=dX, dx=demb, dgamma=grads[prefix + "embeddings.LayerNorm.weight"],
                          dbeta=grads[prefix + "embeddings.LayerNorm.bias"])
        ops.text_embed_bwd(S["ids"], demb, grads[prefix + "embeddings.word_embeddings.weight"],
                           grads[prefix + "embeddings.position_embeddings.weight"], Lq)
