"""Training step of the CompTransTTS path: training-mode forward + backward over libctts_b200's C ABI.

Reference: train.py:104-123 (`output = model(*batch[2:], step=step)`, `Loss(batch, output, step)`, `.backward()`), with
the forward of model/CompTransTTS.py:64-152 in `model.train()` mode -- dropout, BatchNorm on batch statistics
(modules.py:140-148, conformer.py:465), `predictor_grad` gradient scaling (modules.py:1026,893), soft upsampling before
`binarization_start_steps` (modules.py:1047-1049).

The reference gets the backward pass from torch.autograd.  Here the forward records every kernel call on a TAPE
(`Ctx.tape`: closures over the saved activations) and the backward replays the tape in reverse; every arithmetic step
of both passes is a libctts_b200 kernel.  The whole step is ONE torch.autograd.Function (`StepFunction`), so the
reference's `model/loss.py` -- plain PyTorch on the outputs -- drives it unchanged.

Gradients of the parameters are accumulated by the kernels directly into a flat fp32 ARENA (`GradArena`); `param.grad`
is a view into it, so the data-parallel all-reduce of train.py:58 is one (bucketed, overlappable) NCCL call on the arena
with no packing step (ctts_b200/dist.py).
"""
import math
import os

import torch

from . import capi
from .capi import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SWISH, ACT_TANH
from . import engine
from .engine import Planes, _ACTS, _f32, _i64, pad_mask


def _st():
    return engine._stream()


def _ll(*vals):
    import ctypes
    return (ctypes.c_longlong * len(vals))(*[int(v) for v in vals])


def _ints(*vals):
    import ctypes
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


class Var:
    """An activation on the tape: value (fp32), gradient (allocated by its first producer in the backward pass), cached
    bf16 operand planes of the value."""
    __slots__ = ("v", "g", "needs_grad", "_planes", "_tplanes")

    def __init__(self, v, needs_grad=True):
        self.v = v
        self.g = None
        self.needs_grad = needs_grad
        self._planes = {}
        self._tplanes = None

    @property
    def shape(self):
        return self.v.shape


# ---------------------------------------------------------------------------------------------------------------------
class GradArena:
    """One flat fp32 buffer holding the gradient of every trainable parameter; `views[name]` aliases it.

    Parameters are laid out in REVERSE registration order (postnet, mel_linear, decoder, variance adaptor, encoder): the
    order in which the backward pass finishes them, so a prefix of the arena is final early and its all-reduce can overlap
    the rest of the backward (dist.ArenaAllReduce).  Tied parameters (one Parameter under several names) share one slot;
    frozen ones have none."""

    def __init__(self, module):
        named = list(module.named_parameters(remove_duplicate=False))
        seen, order = set(), []
        for name, prm in reversed(named):
            if prm.requires_grad and id(prm) not in seen:
                seen.add(id(prm))
                order.append((name, prm))
        total = sum((p.numel() + 3) // 4 * 4 for _, p in order)     # every slot 16-byte aligned (float4 epilogues)
        dev = order[0][1].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.views, self.offsets, self.params, self.order = {}, {}, [], []
        by_id, off = {}, 0
        for name, prm in order:
            n = prm.numel()
            by_id[id(prm)] = (off, self.flat[off:off + n].view_as(prm))
            self.params.append((prm, by_id[id(prm)][1]))
            self.order.append((name, off, n))
            off += (n + 3) // 4 * 4
        for name, prm in named:
            if prm.requires_grad:
                self.offsets[name], self.views[name] = by_id[id(prm)]
        self.device = dev
        self.sig = self.signature(module)

    def boundary_after(self, prefix):
        """End offset (elements) of the last slot whose name starts with one of `prefix` -- the arena is laid out in
        backward order, so everything before this offset is final once those parameters' gradients are."""
        end = 0
        for name, off, n in self.order:
            if name.startswith(prefix):
                end = max(end, off + (n + 3) // 4 * 4)
        return end

    @staticmethod
    def signature(module):
        return tuple((id(p), p.device) for _, p in module.named_parameters(remove_duplicate=False))

    def begin_backward(self):
        """Zero the arena unless the caller is accumulating over micro-batches (param.grad still aliases it)."""
        accumulating = all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in self.params)
        if not accumulating:
            self.flat.zero_()
        return accumulating

    def publish(self):
        for p, v in self.params:
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v


# ---------------------------------------------------------------------------------------------------------------------
class TrainWeights:
    """Kernel-layout copies of the weights the training step uses (tap-major packed Conv1d weights, bf16 operand planes,
    dgrad re-packs, concatenated projections), built lazily PER WEIGHT and re-built only when that weight's storage or
    version changed -- i.e. once per optimizer step for the weights a step actually touches."""

    def __init__(self):
        self.cache = {}

    def get(self, key, sources, build):
        sig = tuple((t.data_ptr(), t._version) for t in sources)
        e = self.cache.get(key)
        if e is None or e[0] != sig:
            e = (sig, build())
            self.cache[key] = e
        return e[1]


_CONSTANTS = {}


def cwt_scale_weights(device):
    """(i + 1 + 2.5) ** -2.5 for the 10 CWT scales, utils/pitch_tools.py:260."""
    k = ("cwt_scale_w", str(device))
    if k not in _CONSTANTS:
        _CONSTANTS[k] = ((torch.arange(0, 10).float() + 1 + 2.5) ** (-2.5)).to(device)
    return _CONSTANTS[k]


class Ctx:
    """State of one training step: parameters, kernel-layout weights, the tape, dropout stream."""

    def __init__(self, module, arena):
        self.module = module
        self.prep = module._prepared          # positional tables only (they do not depend on the weights)
        P = {k: v for k, v in module.named_parameters(remove_duplicate=False)}
        P.update({k: v for k, v in module.named_buffers(remove_duplicate=False)})
        self.P = P                            # per-step dict: blocks add virtual (concatenated) weights
        self.src = {}                         # virtual weight name -> the parameters it is built from
        self.tw = module._train_weights
        self.G = dict(arena.views)
        self.tape = []
        self.cfg = module.model_config
        self.pcfg = module.preprocess_config
        self.tcfg = module.train_config
        self.enc_math = "tc3" if module.encoder_math == "bf16x6" else "fp32"
        self.dec_math = "tc2" if module.decoder_math == "bf16x3" else "fp32"
        self.bwd_tc = os.environ.get("CTTS_TRAIN_BWD_MATH", "tc") == "tc"
        self.dropout_on = os.environ.get("CTTS_DROPOUT", "1") != "0"
        self.seed = module._dropout_seed
        self.offset = module._dropout_offset
        self.offset_dev = None   # device step counter added to every dropout offset (CUDA-graph mode)
        self.hooks = []          # (tape position, callable): fired when the backward pass reaches that position
        self.marks = {}          # stage name -> index of its first tape closure

    def record(self, fn):
        self.tape.append(fn)

    # -- kernel-layout weights -------------------------------------------------------------------------------------
    def _sources(self, name):
        return self.src.get(name) or [self.P[name]]

    def concat(self, vname, parts):
        """Virtual weight: rows of several projections stacked (q | k | v)."""
        srcs = [self.P[n] for n in parts]
        self.src[vname] = srcs
        self.P[vname] = self.tw.get(("cat", vname), srcs, lambda: torch.cat([_f32(t.detach()) for t in srcs], 0).contiguous())
        return self.P[vname]

    def packed(self, name):
        """fp32 weight as [N, taps*Cin] (tap-major for Conv1d)."""
        t = self.P[name]
        if t.dim() != 3:
            return t

        def build():
            n, cin, taps = t.shape
            out = torch.empty(n, taps * cin, device=t.device, dtype=torch.float32)
            capi.call("ctts_pack_conv_weight", _f32(t.detach()), n, cin, taps, out, _st())
            return out
        return self.tw.get(("packed", name), self._sources(name), build)

    def weight_planes(self, name, n):
        return self.tw.get(("planes", name, n), self._sources(name), lambda: engine.split_planes(self.packed(name), n))

    def dgrad_packed(self, name):
        """[Cin, taps*N] with flipped taps: dx = conv(dz, wd)."""
        t = self.P[name]

        def build():
            N, Cin = t.shape[0], t.shape[1]
            taps = t.shape[2] if t.dim() == 3 else 1
            wd = torch.empty(Cin, taps * N, device=t.device, dtype=torch.float32)
            capi.call("ctts_pack_conv_weight_dgrad", _f32(t.detach()), N, Cin, taps, wd, _st())
            return wd
        return self.tw.get(("dgrad", name), self._sources(name), build)

    def dgrad_planes(self, name):
        return self.tw.get(("dgrad_planes", name), self._sources(name), lambda: engine.split_planes(self.dgrad_packed(name), 2))

    def next_offset(self):
        self.offset += 1
        return self.offset


def planes_of(var, n):
    p = var._planes.get(n)
    if p is None:
        p = var._planes[n] = engine.split_planes(var.v, n)
    return p


def accumulate_into(var, t, scale=1.0):
    """var.g += scale * t (the first contribution adopts the buffer when no scaling is needed)."""
    if not var.needs_grad:
        return
    if var.g is None:
        if scale == 1.0:
            var.g = t
        else:
            var.g = torch.empty_like(t)
            capi.call("ctts_axpy", t, float(scale), t.numel(), 0, var.g, _st())
    else:
        capi.call("ctts_axpy", t, float(scale), t.numel(), 1, var.g, _st())


def accumulate_copy(var, t, scale=1.0):
    """var.g += scale * t without ever adopting `t` (for the second consumer of one gradient buffer)."""
    if not var.needs_grad:
        return
    acc = 1
    if var.g is None:
        var.g = torch.empty_like(t)
        acc = 0
    capi.call("ctts_axpy", t, float(scale), t.numel(), acc, var.g, _st())


def grad_buffer(var):
    """(buffer, accumulate flag) for a kernel that writes the WHOLE gradient of var."""
    if var.g is None:
        var.g = torch.empty_like(var.v)
        return var.g, 0
    return var.g, 1


# ---------------------------------------------------------------------------------------------------------------------
# dense layers
# ---------------------------------------------------------------------------------------------------------------------
def _generic(a, b, y, Z, zmod, M, N, K, a_str, b_str, y_str, Kin=0, shift0=0, shift_z=0, alpha=1.0, accumulate=0):
    capi.call("ctts_gemm_generic", a, b, y, Z, zmod, M, N, K, _ll(*a_str), _ll(*b_str), _ll(*y_str), Kin, shift0, shift_z,
              float(alpha), int(accumulate), _st())


def _tc_dgrad_ok(ctx, N, Cin):
    return ctx.bwd_tc and N % 8 == 0 and Cin % 4 == 0 and N >= 16


def _dgrad(ctx, dz, wname, taps, x, dz_planes=None):
    """x.g (+)= conv(dz, w^T flipped): the data gradient of y = conv(x, w)."""
    B, T, N = dz.shape
    Cin = x.v.shape[-1]
    out, acc = grad_buffer(x)
    res = out if acc else None
    if _tc_dgrad_ok(ctx, N, Cin):
        dzp = dz_planes if dz_planes is not None else engine.split_planes(dz, 2)
        engine.gemm_tc(dzp, ctx.dgrad_planes(wname), residual=res, taps=taps, out=out)
    elif N % 16 == 0:
        engine.conv_gemm(dz, ctx.dgrad_packed(wname), residual=res, taps=taps, out=out)
    else:
        assert taps == 1, "generic dgrad: taps == 1 only"
        w = ctx.P[wname]
        _generic(dz, w, out, 1, 1, B * T, Cin, N, (0, 0, N, 1, 0), (0, 0, 1, Cin, 0), (0, 0, Cin, 1), accumulate=acc)


def transposed_planes(t, taps=1, n=2):
    """fp32 [B, T, C] -> Planes [B, taps, C, Tp] (time contiguous, Tp = T rounded up to 8, tap shifts applied)."""
    B, T, C = t.shape
    Tp = (T + 7) // 8 * 8
    p = Planes.empty((B, taps, C, Tp), t.device, n)
    capi.call("ctts_split_transpose", t, B, T, C, C, 0, Tp, taps, n, capi.ptr_array(p.p), _st())
    return p


WGRAD_ROWMAJOR = os.environ.get("CTTS_WGRAD_ROWMAJOR", "1") != "0"


def _wgrad_rowmajor_ok(ctx, N, Cin):
    """ctts_gemm_wgrad_rowmajor: the weight gradient straight from the row-major planes (no transposed copies)."""
    return WGRAD_ROWMAJOR and ctx.bwd_tc and Cin % 128 == 0 and N % 8 == 0


def _wgrad(ctx, dz, x, wname, taps, dzT=None, dzp=None):
    """G[w] += dz^T (*) x : the weight gradient of y = conv(x, w)."""
    G = ctx.G.get(wname)
    if G is None:
        return
    B, T, N = dz.shape
    Cin = x.v.shape[-1]
    st = _st()
    if _wgrad_rowmajor_ok(ctx, N, Cin):
        if dzp is None:
            dzp = engine.split_planes(dz, 2)
        xp = planes_of(x, 2)
        if taps == 1:
            capi.call("ctts_gemm_wgrad_rowmajor", 2, capi.ptr_array(dzp.p), capi.ptr_array(xp.p), B, T, Cin, N, 1, 1.0, 1, G, st)
        else:
            tmp = torch.empty(N, taps * Cin, device=dz.device, dtype=torch.float32)
            capi.call("ctts_gemm_wgrad_rowmajor", 2, capi.ptr_array(dzp.p), capi.ptr_array(xp.p), B, T, Cin, N, taps, 1.0, 0, tmp,
                      st)
            capi.call("ctts_unpack_conv_wgrad", tmp, N, Cin, taps, 1, G, st)
        return
    if ctx.bwd_tc and Cin % 4 == 0:
        Tp = (T + 7) // 8 * 8
        if dzT is None:
            dzT = transposed_planes(dz)
        if x._tplanes is None:
            x._tplanes = {}
        xT = x._tplanes.get(taps)
        if xT is None:
            xT = x._tplanes[taps] = transposed_planes(x.v, taps)
        if taps == 1:
            capi.call("ctts_gemm_wgrad", 2, capi.ptr_array(dzT.p), capi.ptr_array(xT.p), B, T, Tp, Cin, N, 1, 1.0, 1, G, st)
        else:
            tmp = torch.empty(N, taps * Cin, device=dz.device, dtype=torch.float32)
            capi.call("ctts_gemm_wgrad", 2, capi.ptr_array(dzT.p), capi.ptr_array(xT.p), B, T, Tp, Cin, N, taps, 1.0, 0, tmp,
                      st)
            capi.call("ctts_unpack_conv_wgrad", tmp, N, Cin, taps, 1, G, st)
        return
    if taps == 1:
        _generic(dz, x.v, G, 1, 1, N, Cin, B * T, (0, 0, 1, N, 0), (0, 0, 1, Cin, 0), (0, 0, Cin, 1), accumulate=1)
    else:
        tmp = torch.empty(N, taps * Cin, device=dz.device, dtype=torch.float32)
        _generic(dz, x.v, tmp, taps, 1, N, Cin, B * T, (0, 0, 1, N, T * N), (0, 0, 1, Cin, T * Cin),
                 (Cin, 0, taps * Cin, 1), Kin=T, shift0=-(taps // 2), shift_z=1)
        capi.call("ctts_unpack_conv_wgrad", tmp, N, Cin, taps, 1, G, st)


def linear(ctx, x, wname, bname=None, alpha=1.0, act=ACT_NONE, residual=None, lens=None, taps=1, math="fp32",
           out_planes=0):
    """y = act((conv_taps(x, W) + b) * alpha) [+ residual] [* keep]  (nn.Conv1d 'same' / nn.Linear) on the tape.
    For GELU / Swish the pre-activation is kept for the backward pass."""
    P = ctx.P
    B, T, Cin = x.v.shape
    N = P[wname].shape[0]
    bias = P[bname] if bname else None
    keep_pre = act in (ACT_GELU, ACT_SWISH)
    assert not (keep_pre and (residual is not None or lens is not None))
    assert residual is None or act == ACT_NONE
    fwd_act = ACT_NONE if keep_pre else act
    res_t = residual.v if residual is not None else None
    use_tc = math in ("tc2", "tc3") and Cin % 8 == 0 and N % 4 == 0 and N >= 16
    yp = None
    if use_tc:
        n = 2 if math == "tc2" else 3
        out, yp = engine.gemm_tc(planes_of(x, n), ctx.weight_planes(wname, n), bias, alpha, act=fwd_act, residual=res_t,
                                 lens=lens, taps=taps, want_planes=(out_planes == n and not keep_pre))
    elif Cin % 16 == 0:
        out = engine.conv_gemm(x.v, ctx.packed(wname), bias, alpha, act=fwd_act, residual=res_t, lens=lens, taps=taps)
    else:
        assert taps == 1 and lens is None and act == ACT_NONE and alpha == 1.0
        out = torch.empty(B, T, N, device=x.v.device, dtype=torch.float32)
        capi.call("ctts_linear_smallk", x.v, ctx.packed(wname), bias, res_t, B * T, Cin, N, out, _st())
    pre = None
    if keep_pre:
        pre = out
        out = torch.empty_like(pre)
        pl = Planes.empty(pre.shape, pre.device, out_planes) if out_planes else None
        capi.call("ctts_act_fwd", pre, pre.numel(), int(act), out, out_planes, capi.ptr_array(pl.p) if pl else None, _st())
        yp = pl
    y = Var(out)
    if yp is not None:
        y._planes[yp.n] = yp

    def bwd():
        if y.g is None:
            return
        dz = y.g
        mask = lens
        res_done = False
        if residual is not None and alpha != 1.0:     # y = alpha * (..) + residual: the residual sees the unscaled gradient
            if mask is not None:
                capi.call("ctts_mask_rows", dz, mask, B, T, N, _st())
                mask = None
            accumulate_copy(residual, dz)
            res_done = True
        ref = pre if keep_pre else (y.v if act in (ACT_RELU, ACT_TANH) else None)
        dzp = dzT = None
        want_w = ctx.G.get(wname) is not None and ctx.bwd_tc and Cin % 4 == 0
        want_d = x.needs_grad and _tc_dgrad_ok(ctx, N, Cin)
        if want_w and want_d and N % 4 == 0:
            # one pass: mask + activation' + alpha + bias gradient + the operand planes of both backward GEMMs
            Tp = (T + 7) // 8 * 8
            dzp = Planes.empty((B, T, N), dz.device, 2)
            if not _wgrad_rowmajor_ok(ctx, N, Cin):      # (the row-major weight gradient needs no time-major copy)
                dzT = Planes.empty((B, 1, N, Tp), dz.device, 2)
            capi.call("ctts_act_bwd_planes", dz, ref, int(act), float(alpha), mask, B, T, N, Tp,
                      dz if residual is not None and not res_done else None, 2, capi.ptr_array(dzp.p),
                      capi.ptr_array(dzT.p) if dzT is not None else None, ctx.G.get(bname) if bname else None, _st())
        elif mask is not None or act != ACT_NONE or alpha != 1.0 or bias is not None:
            capi.call("ctts_act_bwd", dz, ref, int(act), float(alpha), mask, 1, T, B * T, N, dz,
                      ctx.G.get(bname) if bname else None, _st())
        if x.needs_grad:
            _dgrad(ctx, dz, wname, taps, x, dzp)
        _wgrad(ctx, dz, x, wname, taps, dzT, dzp)
        if residual is not None and not res_done:
            accumulate_into(residual, dz)
        y.g = None

    ctx.record(bwd)
    return y


def layer_norm(ctx, x, wname, bname, eps, lens=None, planes=0):
    P = ctx.P
    B, T, C = x.v.shape
    if planes:
        out, yp = engine.layernorm_planes(x.v, P[wname], P[bname], eps, lens, want_fp32=True, n=planes)
    else:
        out, yp = engine.layernorm(x.v, P[wname], P[bname], eps, lens), None
    y = Var(out)
    if yp is not None:
        y._planes[planes] = yp

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        dx, acc = grad_buffer(x)
        capi.call("ctts_layernorm_bwd", x.v, P[wname], y.g, float(eps), lens, B, T, C, dx, acc, ctx.G.get(wname),
                  ctx.G.get(bname), _st())
        y.g = None

    ctx.record(bwd)
    return y


def dropout(ctx, x, p):
    """F.dropout in training mode on the library's Philox stream; identity when p == 0 (the parity configuration)."""
    if p <= 0.0 or not ctx.dropout_on:
        return x
    off = ctx.next_offset()
    out = torch.empty_like(x.v)
    capi.call("ctts_dropout", x.v, x.v.numel(), float(p), ctx.seed, off, ctx.offset_dev, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        capi.call("ctts_dropout", y.g, y.g.numel(), float(p), ctx.seed, off, ctx.offset_dev, y.g, _st())
        accumulate_into(x, y.g)
        y.g = None

    ctx.record(bwd)
    return y


def residual_add(ctx, res, y_in, lens):
    """(res + y) * keep -- the un-fused form used when a dropout sits between the GEMM and the residual add."""
    B, T, C = res.v.shape
    out = torch.empty_like(res.v)
    capi.call("ctts_binary", res.v, y_in.v, 0, 0, lens, B, T, C, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        if lens is not None:
            capi.call("ctts_mask_rows", y.g, lens, B, T, C, _st())
        accumulate_copy(res, y.g)
        accumulate_into(y_in, y.g)
        y.g = None

    ctx.record(bwd)
    return y


def dropout_residual(ctx, res, x, p, lens):
    """(res + dropout(x)) * keep in one pass (ctts_dropout_add)."""
    B, T, C = res.v.shape
    off = ctx.next_offset()
    out = torch.empty_like(res.v)
    capi.call("ctts_dropout_add", x.v, res.v, lens, B, T, C, float(p), ctx.seed, off, ctx.offset_dev, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        if lens is not None:
            capi.call("ctts_mask_rows", y.g, lens, B, T, C, _st())
        accumulate_copy(res, y.g)
        capi.call("ctts_dropout", y.g, y.g.numel(), float(p), ctx.seed, off, ctx.offset_dev, y.g, _st())
        accumulate_into(x, y.g)
        y.g = None

    ctx.record(bwd)
    return y


def sublayer(ctx, x, h, wname, bname, lens, p_drop, math, alpha=1.0, taps=1):
    """x + dropout(alpha * linear(h)) masked: fused into the GEMM epilogue when there is no dropout."""
    if p_drop > 0.0 and ctx.dropout_on:
        y = linear(ctx, h, wname, bname, alpha=alpha, taps=taps, math=math)
        return dropout_residual(ctx, x, y, p_drop, lens)
    return linear(ctx, h, wname, bname, alpha=alpha, residual=x, lens=lens, taps=taps, math=math)


# ---------------------------------------------------------------------------------------------------------------------
# attention (fs2 / transformer: softmax(q k^T * scale + key padding mask) v on a packed qkv tensor)
# ---------------------------------------------------------------------------------------------------------------------
def _attention_bwd_fp32(qkv, dO, lens, n_head, scale):
    """Backward of softmax(q k^T * scale + key mask) v with the scores re-materialised in FP32 (CUDA-core GEMMs)."""
    B, T, C3 = qkv.v.shape
    C = C3 // 3
    DH = C // n_head
    st = _st()
    dev = qkv.v.device
    Z = B * n_head
    q = qkv.v
    k = q.view(-1)[C:]
    v = q.view(-1)[2 * C:]
    qs = (T * C3, DH, C3, 1, 0)                      # strides (zo, zi, row, k, kb) of a head slice of qkv
    S = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
    _generic(q, k, S, Z, n_head, T, T, DH, qs, qs, (n_head * T * T, T * T, T, 1), alpha=scale)
    Pm = torch.empty_like(S)
    capi.call("ctts_masked_softmax", S, lens, n_head, Z, T, T, T, 1, Pm, st)
    os_ = (T * C, DH, C, 1, 0)
    dP = S                                           # reuse
    _generic(dO, v, dP, Z, n_head, T, T, DH, os_, qs, (n_head * T * T, T * T, T, 1))
    dqkv = torch.empty_like(q)
    dq, dk, dv = dqkv, dqkv.view(-1)[C:], dqkv.view(-1)[2 * C:]
    ys = (T * C3, DH, C3, 1)
    # dV[s, d] = sum_t P[t, s] dO[t, d]
    _generic(Pm, dO, dv, Z, n_head, T, DH, T, (n_head * T * T, T * T, 1, T, 0), (T * C, DH, 1, C, 0), ys)
    capi.call("ctts_softmax_bwd", Pm, dP, Z, T, T, T, 1.0, dP, st)
    dS = dP
    # dQ[t, d] = scale * sum_s dS[t, s] K[s, d];  dK[s, d] = scale * sum_t dS[t, s] Q[t, d]
    _generic(dS, k, dq, Z, n_head, T, DH, T, (n_head * T * T, T * T, T, 1, 0), (T * C3, DH, 1, C3, 0), ys, alpha=scale)
    _generic(dS, q, dk, Z, n_head, T, DH, T, (n_head * T * T, T * T, 1, T, 0), (T * C3, DH, 1, C3, 0), ys, alpha=scale)
    return dqkv


def _bgemm(a_planes, a_view, w_planes, w_view, addr, y_outer, y_inner, alpha, Z, T, K, N, y):
    capi.call("ctts_gemm_batched_planes", 2, capi.ptr_array(a_planes.p), _ll(*a_view), capi.ptr_array(w_planes.p), _ll(*w_view),
              _ints(*addr), int(y_outer), int(y_inner), float(alpha), None, None, Z, T, K, N, y, None, _st())


def _attention_bwd_tc(qkv, dO, lens, n_head, scale):
    """The same backward on tcgen05 (2 bf16 planes per operand, ctts_gemm_batched_planes): five batched GEMMs over the
    re-materialised probabilities; the operands whose reduction index is time come from ctts_split_transpose."""
    B, T, C3 = qkv.v.shape
    C = C3 // 3
    H = n_head
    DH = C // H
    Z = B * H
    Tp = (T + 7) // 8 * 8
    st = _st()
    dev = qkv.v.device
    qp = planes_of(qkv, 2)                                        # [B, T, 3C]
    qkv_view = (C3, T, B, C3, T * C3)
    big = (H * T * Tp, T * Tp)
    # 1. S = scale * q k^T  -> [Z, T, Tp] fp32
    S = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
    _bgemm(qp, qkv_view, qp, qkv_view, (H, H, 0, DH, H, C, DH, 1, Tp), big[0], big[1], scale, Z, T, DH, Tp, S)
    # 2. P (fp32), its planes and transposed planes
    Pm = torch.empty_like(S)
    capi.call("ctts_masked_softmax", S, lens, H, Z, T, T, Tp, 1, Pm, st)
    PT = Planes.empty((Z, 1, T, Tp), dev, 2)
    capi.call("ctts_split_transpose", Pm, Z, T, T, Tp, 0, Tp, 1, 2, capi.ptr_array(PT.p), st)
    # 3. dP = dO v^T
    dOp = engine.split_planes(dO, 2)
    _bgemm(dOp, (C, T, B, C, T * C), qp, qkv_view, (H, H, 0, DH, H, 2 * C, DH, 1, Tp), big[0], big[1], 1.0, Z, T, DH, Tp, S)
    # 4. dS = P * (dP - sum P dP)
    capi.call("ctts_softmax_bwd", Pm, S, Z, T, T, Tp, 1.0, S, st)
    dSp = engine.split_planes(S, 2)                               # [Z, T, Tp]
    dST = Planes.empty((Z, 1, T, Tp), dev, 2)
    capi.call("ctts_split_transpose", S, Z, T, T, Tp, 0, Tp, 1, 2, capi.ptr_array(dST.p), st)
    # time-major copies of q, k, dO: [B, C, Tp] == [Z, DH, Tp]
    def tr(src, ld, c0):
        p = Planes.empty((B, 1, C, Tp), dev, 2)
        capi.call("ctts_split_transpose", src, B, T, C, ld, c0, Tp, 1, 2, capi.ptr_array(p.p), st)
        return p
    qT, kT, dOT = tr(qkv.v, C3, 0), tr(qkv.v, C3, C), tr(dO, C, 0)
    sq_view = (T, T, Z, Tp, T * Tp)                               # [Z][T rows][T valid of Tp]
    hT_view = (T, DH, Z, Tp, DH * Tp)
    dqkv = torch.empty_like(qkv.v)
    addr = (H, 1, 0, 0, 1, 0, 0, 1, C3)
    # 5. dV[s, d] = sum_t P^T[s, t] dO^T[d, t]
    _bgemm(PT, sq_view, dOT, hT_view, addr, T * C3, DH, 1.0, Z, T, T, DH, dqkv.view(-1)[2 * C:])
    # 6. dQ[t, d] = scale * sum_s dS[t, s] K^T[d, s]
    _bgemm(dSp, sq_view, kT, hT_view, addr, T * C3, DH, scale, Z, T, T, DH, dqkv)
    # 7. dK[s, d] = scale * sum_t dS^T[s, t] Q^T[d, t]
    _bgemm(dST, sq_view, qT, hT_view, addr, T * C3, DH, scale, Z, T, T, DH, dqkv.view(-1)[C:])
    return dqkv


def attention(ctx, qkv, lens, n_head, math):
    B, T, C3 = qkv.v.shape
    C = C3 // 3
    DH = C // n_head
    scale = 1.0 / math_sqrt(DH)
    if math in ("tc2", "tc3") and DH % 64 == 0:
        n = 2 if math == "tc2" else 3
        ap = engine.attention_tc(planes_of(qkv, n), lens, n_head)
        out = torch.empty(B, T, C, device=qkv.v.device, dtype=torch.float32)
        capi.call("ctts_merge_planes", n, capi.ptr_array(ap.p), out.numel(), out, _st())   # fp32 copy for the wgrad operand
        a = Var(out)
        a._planes[n] = ap
    else:
        a = Var(engine.attention(qkv.v, lens, n_head))

    def bwd():
        if a.g is None:
            return
        if ctx.bwd_tc and DH % 64 == 0:
            dqkv = _attention_bwd_tc(qkv, a.g, lens, n_head, scale)
        else:
            dqkv = _attention_bwd_fp32(qkv, a.g, lens, n_head, scale)
        accumulate_into(qkv, dqkv)
        a.g = None

    ctx.record(bwd)
    return a


def math_sqrt(v):
    return math.sqrt(v)


# ---------------------------------------------------------------------------------------------------------------------
# embeddings, positions, length regulator
# ---------------------------------------------------------------------------------------------------------------------
def embed_tokens(ctx, tokens, lens, wname, scale, pos_mode, table=None):
    P = ctx.P
    B, S = tokens.shape
    tab = P[wname]
    C = tab.shape[1]
    pe = table if table is not None else ctx.prep.table_fs2(C, S + 1, tokens.device)
    x = torch.empty(B, S, C, device=tokens.device, dtype=torch.float32)
    word = torch.empty_like(x)
    capi.call("ctts_embed_tokens", tokens, tab, pe, pe.shape[0], float(scale), B, S, C, tab.shape[0], x, word, lens, pos_mode,
              _st())
    xv, wv = Var(x), Var(word)

    def bwd():
        G = ctx.G.get(wname)
        if G is None:
            return
        st = _st()
        if xv.g is not None:    # x = (scale * E[tok] + pe) * keep ; padding_idx 0 receives no gradient (blocks.py:10-15)
            capi.call("ctts_scatter_add_rows", xv.g, tokens, lens, S, B * S, C, tab.shape[0], 0, float(scale), G, st)
        if wv.g is not None:
            capi.call("ctts_scatter_add_rows", wv.g, tokens, None, S, B * S, C, tab.shape[0], 0, float(scale), G, st)
        xv.g = wv.g = None

    ctx.record(bwd)
    return xv, wv


def add_positions(ctx, x, alpha_name, lens, pos_mode=0, table=None):
    """y = (x + alpha * pe[pos]) * keep (transformer_fs2.py:54-60, modules.py:1349-1350)."""
    P = ctx.P
    B, T, C = x.v.shape
    pe = table if table is not None else ctx.prep.table_fs2(C, T + 1, x.v.device)
    out = torch.empty_like(x.v)
    alpha = P[alpha_name] if alpha_name else None
    capi.call("ctts_add_positions", x.v, pe, pe.shape[0], alpha, lens, B, T, C, pos_mode, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        st = _st()
        if alpha_name and alpha_name in ctx.G:
            capi.call("ctts_add_positions_bwd", y.g, x.v, pe, pe.shape[0], lens, B, T, C, pos_mode, ctx.G[alpha_name], st)
        if lens is not None:
            capi.call("ctts_mask_rows", y.g, lens, B, T, C, st)
        accumulate_into(x, y.g)
        y.g = None

    ctx.record(bwd)
    return y


def add_row(ctx, x, row):
    """x + row[b] broadcast over time (speaker / prosody vectors, modules.py:985-988)."""
    B, T, C = x.v.shape
    out = torch.empty_like(x.v)
    capi.call("ctts_add_row_broadcast", x.v, row.v, B, T, C, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        if row.needs_grad:
            if row.g is None:
                row.g = torch.zeros_like(row.v)
            capi.call("ctts_act_bwd", y.g, None, ACT_NONE, 1.0, None, B, T, T, C, None, row.g, _st())
        accumulate_into(x, y.g)
        y.g = None

    ctx.record(bwd)
    return y


def scale_grad(ctx, x, g):
    """x.detach() + g * (x - x.detach()): value unchanged, gradient scaled by `predictor_grad` (modules.py:1026,893)."""
    y = Var(x.v)
    y._planes = x._planes

    def bwd():
        if y.g is None:
            return
        accumulate_into(x, y.g, g)
        y.g = None

    ctx.record(bwd)
    return y


def first_rows(ctx, x):
    """x[:, 0, :] as [1, B, C] (input of the CWT statistics MLP, modules.py:909-912)."""
    B, T, C = x.v.shape
    out = torch.empty(1, B, C, device=x.v.device, dtype=torch.float32)
    capi.call("ctts_copy_rows", x.v, T * C, B, C, out, C, 0, _st())
    y = Var(out)

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        if x.g is None:
            x.g = torch.zeros_like(x.v)
        capi.call("ctts_copy_rows", y.g, C, B, C, x.g, T * C, 1, _st())
        y.g = None

    ctx.record(bwd)
    return y


def length_expand(ctx, x, cum_lr, M):
    B, S, C = x.v.shape
    out = torch.empty(B, M, C, device=x.v.device, dtype=torch.float32)
    capi.call("ctts_length_expand", x.v, None, None, cum_lr, B, S, C, M, 0, out, None, None, 0, _st())
    y = Var(out)

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        dx, acc = grad_buffer(x)
        capi.call("ctts_length_expand_bwd", y.g, cum_lr, B, S, C, M, acc, dx, _st())
        y.g = None

    ctx.record(bwd)
    return y


# ---------------------------------------------------------------------------------------------------------------------
# transformer_fs2 blocks (transformer_fs2.py:47-72,176-239)
# ---------------------------------------------------------------------------------------------------------------------
def fft_layers_fs2(ctx, pre, x, lens, n_layers, n_head, kernel, act, p_drop, math):
    n = {"tc2": 2, "tc3": 3}.get(math, 0)
    for i in range(n_layers):
        lp = "%slayers.%d.op." % (pre, i)
        h = layer_norm(ctx, x, lp + "layer_norm1.weight", lp + "layer_norm1.bias", 1e-12, planes=n)
        qkv = linear(ctx, h, lp + "self_attn.in_proj_weight", math=math, out_planes=n)
        a = attention(ctx, qkv, lens, n_head, math)
        x = sublayer(ctx, x, a, lp + "self_attn.out_proj.weight", None, lens, p_drop, math)
        h = layer_norm(ctx, x, lp + "layer_norm2.weight", lp + "layer_norm2.bias", 1e-12, planes=n)
        f = linear(ctx, h, lp + "ffn.ffn_1.weight", lp + "ffn.ffn_1.bias", alpha=kernel ** -0.5, act=act, taps=kernel,
                   math=math, out_planes=n)
        f = dropout(ctx, f, p_drop)
        x = sublayer(ctx, x, f, lp + "ffn.ffn_2.weight", lp + "ffn.ffn_2.bias", lens, p_drop, math)
    return layer_norm(ctx, x, pre + "layer_norm.weight", pre + "layer_norm.bias", 1e-5, lens, planes=n)


def encoder_fs2(ctx, tokens, src_lens):
    c = ctx.cfg["transformer_fs2"]
    C = c["encoder_hidden"]
    x, word = embed_tokens(ctx, tokens, src_lens, "encoder.embed_tokens.weight", math.sqrt(C), 0)
    x = dropout(ctx, x, c["encoder_dropout"])
    act = _ACTS[ctx.cfg["variance_predictor"]["ffn_act"]]
    x = fft_layers_fs2(ctx, "encoder.", x, src_lens, c["encoder_layer"], c["encoder_head"], c["ffn_kernel_size"], act,
                       c["encoder_dropout"], ctx.enc_math)
    return x, word


def decoder_fs2(ctx, x, mel_lens):
    c = ctx.cfg["transformer_fs2"]
    x = add_positions(ctx, x, "decoder.pos_embed_alpha", mel_lens)
    x = dropout(ctx, x, c["decoder_dropout"])
    act = _ACTS[ctx.cfg["variance_predictor"]["ffn_act"]]
    return fft_layers_fs2(ctx, "decoder.", x, mel_lens, c["decoder_layer"], c["decoder_head"], c["ffn_kernel_size"], act,
                          c["decoder_dropout"], ctx.dec_math)


ENCODERS = {"transformer_fs2": encoder_fs2}
DECODERS = {"transformer_fs2": decoder_fs2}


# ---------------------------------------------------------------------------------------------------------------------
# variance adaptor (modules.py:962-1114)
# ---------------------------------------------------------------------------------------------------------------------
def predictor_stack(ctx, pre, x, n_layers, kernel, lens, p_drop):
    """[pad, Conv1d, ReLU, LayerNorm(channels), Dropout] x n (modules.py:1277-1288,1330-1338)."""
    math = ctx.enc_math
    n = 3 if math == "tc3" else 0
    for l in range(n_layers):
        h = linear(ctx, x, "%sconv.%d.1.weight" % (pre, l), "%sconv.%d.1.bias" % (pre, l), act=ACT_RELU, taps=kernel,
                   math=math)
        x = layer_norm(ctx, h, "%sconv.%d.3.weight" % (pre, l), "%sconv.%d.3.bias" % (pre, l), 1e-12, lens,
                       planes=n if l + 1 < n_layers else 0)
        x = dropout(ctx, x, p_drop)
    return x


def duration_predictor(ctx, x, src_lens):
    vp = ctx.cfg["variance_predictor"]
    pre = "variance_adaptor.duration_predictor."
    h = predictor_stack(ctx, pre, x, vp["dur_predictor_layers"], vp["dur_predictor_kernel"], src_lens, vp["dropout"])
    return linear(ctx, h, pre + "linear.weight", pre + "linear.bias", lens=src_lens)      # [B, S, 1]


def pitch_style_predictor(ctx, pre, xs, alpha=1.0):
    """PitchPredictor / EnergyPredictor.forward, modules.py:1343-1356."""
    vp = ctx.cfg["variance_predictor"]
    xp = add_positions(ctx, xs, pre + "pos_embed_alpha", None)
    h = predictor_stack(ctx, pre, xp, vp["predictor_layers"], vp["predictor_kernel"], None, vp["dropout"])
    return linear(ctx, h, pre + "linear.weight", pre + "linear.bias", alpha=alpha)


def alignment_encoder(ctx, mel, text_embedding, src_lens, attn_prior, spk):
    """AlignmentEncoder.forward, modules.py:1176-1213 -> (attn_soft, attn_logprob) Vars [B, 1, M, S]."""
    pre = "variance_adaptor.aligner."
    P = ctx.P
    B, M, _ = mel.v.shape
    S = text_embedding.v.shape[1]
    st = _st()
    keys, queries = text_embedding, mel
    if spk is not None:
        sv = _reshape(ctx, spk, (1, B, -1))
        ks = linear(ctx, sv, pre + "key_spk_proj.linear.weight")
        qs = linear(ctx, sv, pre + "query_spk_proj.linear.weight")
        keys = add_row(ctx, keys, _reshape(ctx, ks, (B, -1)))
        queries = add_row(ctx, queries, _reshape(ctx, qs, (B, -1)))
    k = linear(ctx, keys, pre + "key_proj.0.conv.weight", pre + "key_proj.0.conv.bias", act=ACT_RELU, taps=3)
    k = linear(ctx, k, pre + "key_proj.2.conv.weight", pre + "key_proj.2.conv.bias")
    q = linear(ctx, queries, pre + "query_proj.0.conv.weight", pre + "query_proj.0.conv.bias", act=ACT_RELU, taps=3)
    q = linear(ctx, q, pre + "query_proj.2.conv.weight", pre + "query_proj.2.conv.bias", act=ACT_RELU)
    q = linear(ctx, q, pre + "query_proj.4.conv.weight", pre + "query_proj.4.conv.bias")
    C = q.v.shape[2]
    temp = float(ctx.cfg["duration_modeling"]["aligner_temperature"])
    soft = torch.empty(B, 1, M, S, device=mel.v.device, dtype=torch.float32)
    logprob = torch.empty_like(soft)
    capi.call("ctts_aligner_attention", q.v, k.v, attn_prior, src_lens, temp, B, M, S, C, soft, logprob, st)
    sv_, lv_ = Var(soft), Var(logprob)

    def bwd():
        if sv_.g is None and lv_.g is None:
            return
        st2 = _st()
        dev = soft.device
        da = torch.empty(B, M, S, device=dev, dtype=torch.float32)
        capi.call("ctts_aligner_attention_bwd", soft, logprob, attn_prior, sv_.g, lv_.g, src_lens, B, M, S, da, st2)
        # a = -temp |q - k|^2 ; sum_s da = 0  =>  dq = 2 temp * da k ;  dk = 2 temp * (da^T q - k * colsum(da))
        dq, accq = grad_buffer(q)
        _generic(da, k.v, dq, B, 1, M, C, S, (M * S, 0, S, 1, 0), (S * C, 0, 1, C, 0), (M * C, 0, C, 1), alpha=2 * temp,
                 accumulate=accq)
        dk, acck = grad_buffer(k)
        _generic(da, q.v, dk, B, 1, S, C, M, (M * S, 0, 1, S, 0), (M * C, 0, 1, C, 0), (S * C, 0, C, 1), alpha=2 * temp,
                 accumulate=acck)
        ca = torch.zeros(B, S, device=dev, dtype=torch.float32)
        capi.call("ctts_act_bwd", da, None, ACT_NONE, 1.0, None, B, M, M, S, None, ca, st2)
        capi.call("ctts_rowscale_axpy", k.v, ca, -2 * temp, B * S, C, 1, dk, st2)
        sv_.g = lv_.g = None

    ctx.record(bwd)
    return sv_, lv_


def _reshape(ctx, x, shape):
    """View with another shape sharing value and gradient storage."""
    y = Var(x.v.view(*shape), x.needs_grad)

    def bwd():
        if y.g is None:
            return
        accumulate_into(x, y.g.view(x.v.shape))
        y.g = None

    ctx.record(bwd)
    return y


def soft_upsample(ctx, attn_soft, x):
    """x_up = bmm(attn_soft, x) (modules.py:1047-1049), attn_soft [B,1,M,S], x [B,S,C]."""
    B, S, C = x.v.shape
    M = attn_soft.v.shape[2]
    out = torch.empty(B, M, C, device=x.v.device, dtype=torch.float32)
    _generic(attn_soft.v, x.v, out, B, 1, M, C, S, (M * S, 0, S, 1, 0), (S * C, 0, 1, C, 0), (M * C, 0, C, 1))
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        da, acca = grad_buffer(attn_soft)        # dA[m, s] = sum_c dy[m, c] x[s, c]
        _generic(y.g, x.v, da, B, 1, M, S, C, (M * C, 0, C, 1, 0), (S * C, 0, C, 1, 0), (M * S, 0, S, 1), accumulate=acca)
        if x.needs_grad:
            dx, accx = grad_buffer(x)            # dx[s, c] = sum_m A[m, s] dy[m, c]
            _generic(attn_soft.v, y.g, dx, B, 1, S, C, M, (M * S, 0, 1, S, 0), (M * C, 0, 1, C, 0), (S * C, 0, C, 1),
                     accumulate=accx)
        y.g = None

    ctx.record(bwd)
    return y


def variance_adaptor(ctx, spk, text, text_embedding, src_lens, mel, mel_lens, max_len, pitch_target, energy_target,
                     duration_target, attn_prior, p_control, e_control, step):
    """VarianceAdaptor.forward in training mode (targets given).  Returns a dict of Vars / tensors."""
    cfg, pcfg, tcfg, P = ctx.cfg, ctx.pcfg, ctx.tcfg, ctx.P
    B, S, C = text.v.shape
    st = _st()
    dev = text.v.device
    g_pred = float(cfg["variance_predictor"]["predictor_grad"])
    x = add_row(ctx, text, spk) if spk is not None else text
    prosody_info = None
    if cfg["prosody_modeling"]["model_type"] == "liu2021":
        from . import train_prosody
        x, prosody_info = train_prosody.liu2021(ctx, x, src_lens, mel, mel_lens)
    elif cfg["prosody_modeling"]["model_type"] != "none":
        raise NotImplementedError("prosody model %r" % cfg["prosody_modeling"]["model_type"])
    log_d = duration_predictor(ctx, scale_grad(ctx, x, g_pred), src_lens)

    attn = None
    if attn_prior is not None:
        assert cfg["duration_modeling"]["learn_alignment"] and duration_target is None and mel is not None
        attn_soft, attn_logprob = alignment_encoder(ctx, Var(_f32(mel), False), text_embedding, src_lens, _f32(attn_prior),
                                                    spk)
        M_in = mel.shape[1]
        prev_ws = torch.empty(B * M_in * S, device=dev, dtype=torch.uint8)
        attn_hard = torch.empty(B, 1, M_in, S, device=dev, dtype=torch.float32)
        attn_hard_dur = torch.empty(B, S, device=dev, dtype=torch.float32)
        capi.call("ctts_mas", attn_soft.v, src_lens, mel_lens, B, M_in, S, prev_ws, attn_hard, attn_hard_dur, st)
        attn = (attn_soft, attn_hard, attn_hard_dur, attn_logprob)
        duration_rounded = attn_hard_dur
    else:
        assert duration_target is not None, "training needs duration targets or attention priors"
        assert not cfg["duration_modeling"]["learn_alignment"]
        duration_rounded = duration_target
    cum_lr, cum_m2p, lens2 = engine.length_scan(duration_rounded, src_lens, B, S, dev)
    soft = attn_prior is not None and step < tcfg["duration"]["binarization_start_steps"]
    need_m2p = attn_prior is not None
    if max_len is None:
        maxes = lens2.view(2, B).max(dim=1).values.tolist()       # host sync (only when the caller gives no max_mel_len)
        M, M2 = int(maxes[0]), int(maxes[1])
    elif need_m2p:
        # MAS assigns every frame of utterance b to exactly one phoneme, so the durations sum to mel_lens[b] and the length
        # of mel2ph is max(mel_lens) -- which is what train.py passes as max_mel_len (utils/tools.py:69-147): no host sync
        M = M2 = int(max_len)
    else:
        M, M2 = int(max_len), 0
    mel2ph = torch.empty(B, M2, device=dev, dtype=torch.int64) if (need_m2p and M2 > 0) else None
    if mel2ph is not None:
        dummy = torch.empty(B, 1, C, device=dev, dtype=torch.float32)
        capi.call("ctts_length_expand", x.v, None, None, cum_lr, B, S, C, 1, 0, dummy, cum_m2p, mel2ph, M2, st)
    if soft:
        xe = soft_upsample(ctx, attn[0], x)
        mel_len = mel_lens
        M = xe.v.shape[1]
    else:
        xe = length_expand(ctx, x, cum_lr, M)
        mel_len = lens2[:B]
    if attn_prior is not None:
        m2p = mel2ph if mel2ph is not None else torch.zeros(B, 0, device=dev, dtype=torch.int64)
        pitch_target["mel2ph"] = m2p[:, :max_len]

    x_sum = Var(xe.v.clone())
    scatter_jobs = []
    pitch_pred = energy_pred = None
    pre = "variance_adaptor."
    pitch_cfg = pcfg["preprocessing"]["pitch"]
    if cfg["variance_embedding"]["use_pitch_embed"] and pitch_cfg["pitch_type"] != "cwt":
        # pitch_type 'frame' / 'ph' (modules.py:890-906,927-938): one PitchPredictor; the embedding comes from the targets
        assert pitch_cfg["pitch_norm"] == "log" and pitch_target is not None, "training needs pitch targets"
        m2p = _i64(pitch_target["mel2ph"])
        assert m2p.shape[1] == M, "mel2ph length %d != regulated length %d" % (m2p.shape[1], M)
        idx = torch.empty(B, M, device=dev, dtype=torch.int64)
        if pitch_cfg["pitch_type"] == "frame":
            ppred = pitch_style_predictor(ctx, pre + "pitch_predictor.", scale_grad(ctx, xe, g_pred), alpha=p_control)
            f0t = _f32(pitch_target["f0"])
            f0_denorm = torch.empty(B, M, device=dev, dtype=torch.float32)
            capi.call("ctts_frame_pitch", None, 0, f0t, _f32(pitch_target["uv"]), m2p, 1 if pitch_cfg["use_uv"] else 0, B * M,
                      f0t, f0_denorm, idx, st)
            pitch_target["f0"] = f0t
        else:
            ppred = pitch_style_predictor(ctx, pre + "pitch_predictor.", scale_grad(ctx, x, g_pred), alpha=p_control)
            f0 = torch.empty(B, S, device=dev, dtype=torch.float32)
            capi.call("ctts_phoneme_pitch", _f32(pitch_target["f0"]), m2p, src_lens, _i64(mel_len), B, S, M, f0, st)
            pitch_target["f0"] = f0
            f0_denorm = torch.empty(B, S, device=dev, dtype=torch.float32)
            idx_ph = torch.empty(B, S, device=dev, dtype=torch.int64)
            capi.call("ctts_f0_to_pitch", f0, None, B * S, f0_denorm, idx_ph, st)
            capi.call("ctts_gather_index", idx_ph, m2p, B, S, M, idx, st)
        emb = P[pre + "pitch_embed.weight"]
        capi.call("ctts_gather_add", emb, idx, B * M, C, emb.shape[0], x_sum.v, st)
        scatter_jobs.append(("frame", pre + "pitch_embed.weight", idx))
        pitch_pred = {"pitch_pred": ppred, "f0_denorm": f0_denorm, "cwt": None, "stats": None}
    elif cfg["variance_embedding"]["use_pitch_embed"]:
        assert pitch_cfg["pitch_norm"] == "log"
        h = linear(ctx, scale_grad(ctx, xe, g_pred), pre + "cwt_predictor.0.weight", pre + "cwt_predictor.0.bias",
                   math=ctx.enc_math)
        cwt = pitch_style_predictor(ctx, pre + "cwt_predictor.1.", h, alpha=p_control)
        first = first_rows(ctx, x)
        s = linear(ctx, first, pre + "cwt_stats_layers.0.weight", pre + "cwt_stats_layers.0.bias", act=ACT_RELU)
        s = linear(ctx, s, pre + "cwt_stats_layers.2.weight", pre + "cwt_stats_layers.2.bias", act=ACT_RELU)
        stats = linear(ctx, s, pre + "cwt_stats_layers.4.weight", pre + "cwt_stats_layers.4.bias")     # [1, B, 2]
        assert pitch_target is not None, "training needs pitch targets"
        m2p = pitch_target["mel2ph"]
        assert m2p.shape[1] == M, "mel2ph length %d != regulated length %d" % (m2p.shape[1], M)
        f0n = torch.empty(B, M, device=dev, dtype=torch.float32)
        f0_denorm = torch.empty(B, M, device=dev, dtype=torch.float32)
        idx = torch.empty(B, M, device=dev, dtype=torch.int64)
        spec = _f32(pitch_target["cwt_spec"])
        capi.call("ctts_cwt_to_pitch", spec, spec.shape[-1], cwt_scale_weights(dev), _f32(pitch_target["f0_mean"]),
                  _f32(pitch_target["f0_std"]), 1, 1.0, float(pitch_cfg["pitch_norm_eps"]), _f32(pitch_target["uv"]),
                  1 if pitch_cfg["use_uv"] else 0, B, M, f0n, f0_denorm, idx, st)
        pitch_target["f0"] = f0n
        pitch_target["f0_cwt"] = f0n
        emb = P[pre + "pitch_embed.weight"]
        capi.call("ctts_gather_add", emb, idx, B * M, C, emb.shape[0], x_sum.v, st)
        scatter_jobs.append(("frame", pre + "pitch_embed.weight", idx))
        pitch_pred = {"pitch_pred": None, "f0_denorm": f0_denorm, "cwt": cwt, "stats": stats}
    if cfg["variance_embedding"]["use_energy_embed"]:
        level = pcfg["preprocessing"]["energy"]["feature"]
        bins = P[pre + "energy_bins"]
        emb = P[pre + "energy_embedding.weight"]
        assert energy_target is not None, "training needs energy targets"
        if level == "frame_level":
            pred = pitch_style_predictor(ctx, pre + "energy_predictor.", xe)
            eidx = torch.empty(B, M, device=dev, dtype=torch.int64)
            capi.call("ctts_bucketize", _f32(energy_target), 1.0, bins, bins.shape[0], B * M, eidx, st)
            capi.call("ctts_gather_add", emb, eidx, B * M, C, emb.shape[0], x_sum.v, st)
            scatter_jobs.append(("frame", pre + "energy_embedding.weight", eidx))
        else:
            if attn_prior is not None:   # frame-level target -> phoneme level by the hard durations (modules.py:1096-1097)
                et = _f32(energy_target)
                M_e = et.shape[1]
                work = torch.empty(B * M_e, device=dev, dtype=torch.float32)
                energy_target = torch.empty(B, S, device=dev, dtype=torch.float32)
                capi.call("ctts_phoneme_energy", attn[2], src_lens, et, B, S, M_e, work, energy_target, st)
            pred = pitch_style_predictor(ctx, pre + "energy_predictor.", x)
            eidx = torch.empty(B, S, device=dev, dtype=torch.int64)
            capi.call("ctts_bucketize", _f32(energy_target), 1.0, bins, bins.shape[0], B * S, eidx, st)
            capi.call("ctts_length_expand", None, emb, eidx, cum_lr, B, S, C, M, 1, x_sum.v, None, None, 0, st)
            scatter_jobs.append(("phoneme", pre + "energy_embedding.weight", eidx))
        energy_pred = pred

    def bwd_sum():
        if x_sum.g is None:
            return
        st2 = _st()
        for kind, name, index in scatter_jobs:
            G = ctx.G.get(name)
            if G is None:
                continue
            if kind == "frame":
                capi.call("ctts_scatter_add_rows", x_sum.g, index, None, M, B * M, C, G.shape[0], 0, 1.0, G, st2)
            else:
                tmp = torch.empty(B, S, C, device=dev, dtype=torch.float32)
                capi.call("ctts_length_expand_bwd", x_sum.g, cum_lr, B, S, C, M, 0, tmp, st2)
                capi.call("ctts_scatter_add_rows", tmp, index, None, S, B * S, C, G.shape[0], 0, 1.0, G, st2)
        accumulate_into(xe, x_sum.g)
        x_sum.g = None

    # x_sum's consumers (the decoder) are recorded later, so this closure runs after them in the backward pass; but it
    # must run BEFORE the closures of xe's other consumers have finished?  No: gradient fan-in is additive and every
    # producer closure of xe (length_expand / soft_upsample) was recorded before this point, hence runs after it.
    ctx.record(bwd_sum)
    return dict(x=x_sum, log_d=log_d, duration_rounded=duration_rounded, mel_len=mel_len, attn=attn,
                pitch_target=pitch_target, pitch_pred=pitch_pred, energy_target=energy_target, energy_pred=energy_pred,
                prosody_info=prosody_info, M=M)


# ---------------------------------------------------------------------------------------------------------------------
# mel head: mel_linear + PostNet with batch-statistics BatchNorm (CompTransTTS.py:133-135, modules.py:140-148)
# ---------------------------------------------------------------------------------------------------------------------
def batch_norm_act(ctx, z, pre, act, out_planes=0, momentum=0.1, eps=1e-5):
    """y = act(BatchNorm1d(z)) with batch statistics over all B*T rows; updates the running buffers like nn.BatchNorm1d."""
    P = ctx.P
    B, T, C = z.v.shape
    rows = B * T
    st = _st()
    dev = z.v.device
    mean = torch.empty(C, device=dev, dtype=torch.float32)
    var = torch.empty(C, device=dev, dtype=torch.float32)
    capi.call("ctts_bn_stats", z.v, rows, C, mean, var, st)
    out = torch.empty_like(z.v)
    pl = Planes.empty(z.v.shape, dev, out_planes) if out_planes else None
    capi.call("ctts_bn_act_fwd", z.v, mean, var, P[pre + "weight"], P[pre + "bias"], float(eps), int(act), rows, C, out,
              out_planes, capi.ptr_array(pl.p) if pl else None, st)
    capi.call("ctts_bn_update_running", mean, var, rows, float(momentum), C, P[pre + "running_mean"], P[pre + "running_var"],
              P[pre + "num_batches_tracked"], st)
    y = Var(out)
    if pl is not None:
        y._planes[out_planes] = pl

    def bwd():
        if y.g is None:
            return
        ws = torch.empty(2 * C, device=dev, dtype=torch.float32)
        dz = torch.empty_like(z.v)
        capi.call("ctts_bn_bwd", y.g, z.v, mean, var, P[pre + "weight"], P[pre + "bias"], float(eps), int(act), rows, C, dz,
                  ctx.G.get(pre + "weight"), ctx.G.get(pre + "bias"), ws, _st())
        accumulate_into(z, dz)
        y.g = None

    ctx.record(bwd)
    return y


def add(ctx, a, b):
    """a + b (PostNet residual, CompTransTTS.py:135)."""
    out = torch.empty_like(a.v)
    capi.call("ctts_axpy", a.v, 1.0, a.v.numel(), 0, out, _st())
    capi.call("ctts_axpy", b.v, 1.0, b.v.numel(), 1, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        accumulate_copy(b, y.g)        # two consumers of one buffer must not alias (both may be updated in place)
        accumulate_into(a, y.g)
        y.g = None

    ctx.record(bwd)
    return y


def mel_head(ctx, dec):
    math = ctx.dec_math
    n = 2 if math == "tc2" else 0
    mel = linear(ctx, dec, "mel_linear.weight", "mel_linear.bias", math=math, out_planes=n)
    h = mel
    for i in range(5):
        pre = "postnet.convolutions.%d." % i
        z = linear(ctx, h, pre + "0.conv.weight", pre + "0.conv.bias", taps=5, math=math)
        h = batch_norm_act(ctx, z, pre + "1.", ACT_TANH if i < 4 else ACT_NONE, out_planes=n if i < 4 else 0)
        h = dropout(ctx, h, 0.5)       # hard-coded in the reference (modules.py:144-145)
    return mel, add(ctx, h, mel)


# ---------------------------------------------------------------------------------------------------------------------
# the step
# ---------------------------------------------------------------------------------------------------------------------
def forward_train(ctx, speakers, texts, src_lens, max_src_len, mels, mel_lens, max_mel_len, p_targets, e_targets, d_targets,
                  attn_priors, spker_embeds, p_control, e_control, d_control, step):
    """Training-mode CompTransTTS.forward on the tape.  Returns (outputs pytree with Vars at the differentiable leaves)."""
    module, cfg, P = ctx.module, ctx.cfg, ctx.P
    block = cfg["block_type"]
    if block not in ENCODERS:
        from . import train_blocks  # noqa: F401  (registers the other block types)
    if block not in ENCODERS:
        raise NotImplementedError("block_type %r: training kernels not built" % block)
    texts = _i64(texts)
    src_lens = _i64(src_lens)
    mel_lens_t = _i64(mel_lens) if mel_lens is not None else None
    B, S = texts.shape
    enc, word = ENCODERS[block](ctx, texts, src_lens)
    ctx.marks["variance_adaptor"] = len(ctx.tape)     # closures below this index belong to the encoder
    spk = None
    if module.has_speaker_emb:
        if module.embedder_type == "none":
            spk_idx = _i64(speakers)
            tab = P["speaker_emb.weight"]
            sv = torch.zeros(B, tab.shape[1], device=texts.device, dtype=torch.float32)
            capi.call("ctts_gather_add", tab, spk_idx, B, tab.shape[1], tab.shape[0], sv, _st())
            spk = Var(sv)

            def bwd_spk():
                if spk.g is not None and "speaker_emb.weight" in ctx.G:
                    capi.call("ctts_scatter_add_rows", spk.g, spk_idx, None, 1, B, tab.shape[1], tab.shape[0], -1, 1.0,
                              ctx.G["speaker_emb.weight"], _st())
            ctx.record(bwd_spk)
        else:
            assert spker_embeds is not None, "Speaker embedding should not be None"
            e = Var(_f32(spker_embeds).view(1, B, -1), False)
            spk = _reshape(ctx, linear(ctx, e, "speaker_emb.weight", "speaker_emb.bias"), (B, -1))
    va = variance_adaptor(ctx, spk, enc, word, src_lens, mels, mel_lens_t, max_mel_len, p_targets, e_targets, d_targets,
                          attn_priors, p_control, e_control, step)
    ctx.marks["decoder"] = len(ctx.tape)               # closures from here on: decoder + mel head
    dec = DECODERS[block](ctx, va["x"], va["mel_len"])
    mel, post = mel_head(ctx, dec)
    return va, mel, post


class StepFunction(torch.autograd.Function):
    """The whole training-mode forward as one autograd node: forward runs the tape-recording engine, backward replays the
    tape.  Inputs: a dummy tensor that requires grad (so that autograd calls backward) + the context."""

    @staticmethod
    def forward(fctx, anchor, holder):
        outs = holder["run"]()
        fctx.holder = holder
        return tuple(outs)

    @staticmethod
    def backward(fctx, *grads):
        holder = fctx.holder
        holder["backward"](grads)
        return None, None


def _seed_gradients(out_vars, grads, world):
    """Private copies of the incoming output gradients with DDP's 1/world mean folded in (train.py:58)."""
    st = _st()
    inv = 1.0 / world
    for var, g in zip(out_vars, grads):
        if var is None or g is None:
            continue
        buf = torch.empty(var.v.shape, device=var.v.device, dtype=torch.float32)
        gc = g if (g.is_contiguous() and g.dtype == torch.float32) else g.float().contiguous()
        capi.call("ctts_axpy", gc, inv, gc.numel(), 0, buf, st)
        var.g = buf


def _segments(ctx, reducer):
    """Tape ranges [hi, lo] (replayed downwards) between the points where a slice of the gradient arena becomes final."""
    n = len(ctx.tape)
    if reducer is None or reducer.world == 1 or not reducer.enabled:
        return [(n - 1, 0, None)]
    cuts = sorted({min(max(i, 0), n) for i, _ in ctx.hooks}, reverse=True)
    fire = {}
    for i, fn in ctx.hooks:
        fire.setdefault(min(max(i, 0), n), []).append(fn)
    segs, hi = [], n - 1
    for c in cuts:
        if c <= 0:
            continue
        segs.append((hi, c, fire[c]))
        hi = c - 1
    segs.append((hi, 0, fire.get(0)))
    return segs


def run_backward(ctx, arena, out_vars, grads, world=1, reducer=None):
    """Eager backward: replay the tape in reverse; launch the arena all-reduce of a stage as soon as the stage is done."""
    arena.begin_backward()
    _seed_gradients(out_vars, grads, world)
    for hi, lo, hooks in _segments(ctx, reducer):
        for i in range(hi, lo - 1, -1):
            ctx.tape[i]()
        for h in hooks or ():
            h()
    ctx.tape = []
    if reducer is not None:
        reducer.finish()
    arena.publish()


def _leaves_of(va, mel, post):
    leaves = [("mel", mel), ("post", post), ("log_d", va["log_d"])]
    if va["pitch_pred"] is not None:
        if va["pitch_pred"]["cwt"] is not None:
            leaves += [("cwt", va["pitch_pred"]["cwt"]), ("stats", va["pitch_pred"]["stats"])]
        if va["pitch_pred"].get("pitch_pred") is not None:
            leaves.append(("pitch_pred", va["pitch_pred"]["pitch_pred"]))
    if va["energy_pred"] is not None:
        leaves.append(("e_pred", va["energy_pred"]))
    if va["attn"] is not None:
        leaves += [("attn_soft", va["attn"][0]), ("attn_logprob", va["attn"][3])]
    if va["prosody_info"] is not None:
        for i, v in enumerate(va["prosody_info"]):
            if isinstance(v, Var):
                leaves.append(("prosody.%d" % i, v))
    return leaves


def _side_outputs(va):
    """The non-differentiable tensors of the 14-tuple that the forward computes (pytree of tensors)."""
    return dict(f0_denorm=va["pitch_pred"]["f0_denorm"] if va["pitch_pred"] is not None else None,
                attn_hard=va["attn"][1] if va["attn"] is not None else None,
                attn_hard_dur=va["attn"][2] if va["attn"] is not None else None,
                prosody=tuple(None if isinstance(v, Var) else v for v in va["prosody_info"])
                if va["prosody_info"] is not None else None,
                duration_rounded=va["duration_rounded"], mel_len=va["mel_len"], pitch_target=va["pitch_target"],
                energy_target=va["energy_target"])


class TrainGraphs:
    """Shape-keyed cache of training steps captured as CUDA graphs.  A step is run eagerly the first time its key is seen,
    captured the second time (forward graph; the backward graph(s) at the first backward call) and replayed afterwards:
    the eager step is bound by ~1000 Python -> ctypes launches, not by the GPU.  With a data-parallel reducer the backward
    is captured as one graph per stage so that the arena all-reduce of a stage is launched between them."""

    def __init__(self, max_entries=4):
        self.entries = {}
        self.max_entries = max_entries
        self.pool = None

    def clear(self):
        self.entries.clear()


def _unflatten_inputs(t):
    pt = {k[len("p_targets."):]: v for k, v in t.items() if k.startswith("p_targets.")}
    return dict(speakers=t.get("speakers"), texts=t["texts"], src_lens=t["src_lens"], mels=t.get("mels"),
                mel_lens=t.get("mel_lens"), p_targets=pt if pt else None, e_targets=t.get("e_targets"),
                d_targets=t.get("d_targets"), attn_priors=t.get("attn_priors"), spker_embeds=t.get("spker_embeds"))


class _Step:
    """One training-mode forward on the tape, and its backward."""

    def __init__(self, module, arena, scalars, graph_mode=False):
        self.module, self.arena, self.scalars = module, arena, scalars
        self.graph_mode = graph_mode
        self.ctx = None
        self.bwd_graphs = None
        self.bwd_pattern = None
        self.pending = False

    def record(self, t):
        """Run forward_train on the tensors `t` (tape recorded in self.ctx)."""
        module = self.module
        max_src_len, max_mel_len, p_control, e_control, d_control, step = self.scalars
        ctx = Ctx(module, self.arena)
        if self.graph_mode:
            ctx.tw = TrainWeights()                 # re-layout kernels are captured and re-run with every replay
            ctx.offset_dev = module.dropout_counter(t["texts"].device)
            ctx.offset = 0
        a = _unflatten_inputs(t)
        with torch.no_grad():
            va, mel, post = forward_train(ctx, a["speakers"], a["texts"], a["src_lens"], max_src_len, a["mels"], a["mel_lens"],
                                          max_mel_len, a["p_targets"], a["e_targets"], a["d_targets"], a["attn_priors"],
                                          a["spker_embeds"], p_control, e_control, d_control, step)
        if not self.graph_mode:
            module._dropout_offset = ctx.offset
        self.ctx = ctx
        self.n_dropout_sites = ctx.offset
        leaves = _leaves_of(va, mel, post)
        self.names = [n for n, _ in leaves]
        self.vars = [v for _, v in leaves]
        self.side = _side_outputs(va)
        return [v.v for v in self.vars]

    # -- backward ------------------------------------------------------------------------------------------------
    def backward(self, grads):
        module, arena, ctx = self.module, self.arena, self.ctx
        reducer = module._reducer
        world = reducer.world if reducer is not None else 1
        if reducer is not None:
            reducer.begin(ctx, arena)
        if not self.graph_mode:
            run_backward(ctx, arena, self.vars, grads, world=world, reducer=reducer)
            return
        pattern = tuple(g is not None for g in grads)
        arena.begin_backward()
        gs = [None if g is None else (g if (g.is_contiguous() and g.dtype == torch.float32) else g.float().contiguous())
              for g in grads]
        if self.bwd_graphs is None or self.bwd_pattern != pattern:
            self.bwd_pattern = pattern
            self.static_grads = [None if g is None else g.clone() for g in gs]
            self.bwd_graphs = []
            segs = _segments(ctx, reducer)
            cache = module._train_graphs
            torch.cuda.synchronize()
            for si, (hi, lo, hooks) in enumerate(segs):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=cache.pool):
                    if si == 0:
                        _seed_gradients(self.vars, self.static_grads, world)
                    for i in range(hi, lo - 1, -1):
                        ctx.tape[i]()
                self.bwd_graphs.append((g, hooks))
            ctx.tape = None        # the closures live on in the graphs
        for dst, src in zip(self.static_grads, gs):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        for g, hooks in self.bwd_graphs:
            g.replay()
            for h in hooks or ():
                h()
        if reducer is not None:
            reducer.finish()
        arena.publish()
        self.pending = False


def _clone_tree(v):
    return engine._tree_map(lambda t: t.clone(), v)


def forward(module, speakers, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None, p_targets=None,
            e_targets=None, d_targets=None, attn_priors=None, spker_embeds=None, p_control=1.0, e_control=1.0,
            d_control=1.0, step=None):
    """CompTransTTS.forward in model.train() mode (model/CompTransTTS.py:64-152): the reference's 14-tuple, connected to
    autograd through StepFunction."""
    capi.require_device()
    capi.require_cuda_tensor(texts)
    arena = module.grad_arena()
    tcfg = module.train_config
    soft = attn_priors is not None and step is not None and step < tcfg["duration"]["binarization_start_steps"]
    tin = {}
    engine._flatten("", dict(speakers=speakers if module.has_speaker_emb and module.embedder_type == "none" else None,
                             texts=_i64(texts), src_lens=_i64(src_lens), mels=mels if attn_priors is not None or
                             module.model_config["prosody_modeling"]["model_type"] != "none" else None,
                             mel_lens=_i64(mel_lens) if mel_lens is not None else None, p_targets=p_targets,
                             e_targets=e_targets, d_targets=d_targets, attn_priors=attn_priors,
                             spker_embeds=spker_embeds if module.has_speaker_emb and module.embedder_type != "none" else None),
                    tin)
    scalars = (max_src_len, max_mel_len, float(p_control), float(e_control), float(d_control), step)
    graphs = module._train_graphs if (module.use_cuda_graphs and texts.is_cuda and
                                      os.environ.get("CTTS_TRAIN_GRAPHS", "1") != "0") else None
    st = None
    outs_are_static = False
    if graphs is not None and max_mel_len is not None:
        if graphs.sig != arena.sig:
            graphs.clear()
            graphs.sig = arena.sig
        key = (max_src_len, max_mel_len, float(p_control), float(e_control), float(d_control), bool(soft),
               os.environ.get("CTTS_DROPOUT", "1"), os.environ.get("CTTS_TRAIN_BWD_MATH", "tc"), module._reducer is not None
               and module._reducer.enabled, tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(tin.items())))
        e = graphs.entries.get(key)
        if e is None:
            if len(graphs.entries) >= graphs.max_entries:
                graphs.entries.pop(next(iter(graphs.entries)))
            graphs.entries[key] = False
        elif e is False:
            st = _Step(module, arena, scalars, graph_mode=True)
            st.static_in = {k: v.clone() for k, v in tin.items()}
            module.dropout_counter(texts.device)     # allocate OUTSIDE the capture (a captured zeros() would re-zero it)
            torch.cuda.synchronize()
            if graphs.pool is None:
                graphs.pool = torch.cuda.graph_pool_handle()
            st.fwd_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st.fwd_graph, pool=graphs.pool):
                st.static_out = st.record(st.static_in)
            graphs.entries[key] = st
        elif not e.pending:
            st = e
            graphs.entries[key] = graphs.entries.pop(key)       # LRU
        if st is not None:
            for k, v in tin.items():
                st.static_in[k].copy_(v, non_blocking=True)
            if st.n_dropout_sites:
                module.dropout_counter(texts.device).add_(st.n_dropout_sites)
            st.fwd_graph.replay()
            st.pending = True
            outs_are_static = True
    if st is None:
        st = _Step(module, arena, scalars)

    def run():
        if outs_are_static:
            return [t.clone() for t in st.static_out]      # the static buffers are overwritten by the next replay
        return st.record(tin)

    holder = {"run": run, "backward": st.backward}
    outs = StepFunction.apply(module.autograd_anchor(texts.device), holder)
    o = dict(zip(st.names, outs))
    side = _clone_tree(st.side) if outs_are_static else st.side
    B = texts.shape[0]
    src_lens = _i64(src_lens)
    src_masks = pad_mask(src_lens, max_src_len)
    mel_masks = pad_mask(_i64(mel_lens), max_mel_len) if mel_lens is not None else None
    p_pred = None
    if "cwt" in o:
        stats = o["stats"].view(B, 2)
        p_pred = {"pitch_pred": None, "f0_denorm": side["f0_denorm"], "cwt": o["cwt"], "f0_mean": stats[:, 0],
                  "f0_std": stats[:, 1]}
    elif "pitch_pred" in o:
        p_pred = {"pitch_pred": o["pitch_pred"], "f0_denorm": side["f0_denorm"], "cwt": None, "f0_mean": None, "f0_std": None}
    e_pred = o["e_pred"].squeeze(-1) if "e_pred" in o else None
    attn_outs = (None, None, None, None)
    if "attn_soft" in o:
        attn_outs = (o["attn_soft"], side["attn_hard"], side["attn_hard_dur"], o["attn_logprob"])
    prosody = None
    if side["prosody"] is not None:
        prosody = tuple(o.get("prosody.%d" % i, v) for i, v in enumerate(side["prosody"]))
    d_rounded = d_targets if (d_targets is not None and attn_priors is None) else side["duration_rounded"]
    p_out = side["pitch_target"]
    if p_targets is not None and p_out is not None and p_out is not p_targets:
        p_targets.update(p_out)        # the reference mutates the caller's dict (modules.py:1053,1078-1083; train.py:107)
        p_out = p_targets
    return (o["mel"], o["post"], p_pred, e_pred, o["log_d"].squeeze(-1), d_rounded, src_masks, mel_masks, src_lens,
            side["mel_len"], attn_outs, prosody, p_out, side["energy_target"])
