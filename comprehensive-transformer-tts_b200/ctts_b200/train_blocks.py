"""Training step of the `transformer`, `fastformer` and `conformer` block types (train_engine.py holds the tape, the
shared layers, the fs2 stack, the VarianceAdaptor and the mel head).  References: model/transformers/transformer.py:157-288,
fastformer.py:140-376, conformer.py:162-560; the decoders' training-mode `max_seq_len` truncation
(transformer.py:137-145) is refused, like in the oracle, rather than silently diverging."""
import math

import torch

from . import capi, engine_blocks
from .capi import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SWISH
from . import train_engine as TE
from .train_engine import (Var, accumulate_copy, accumulate_into, add, add_positions, batch_norm_act, dropout, embed_tokens,
                           grad_buffer, layer_norm, linear, sublayer, _generic, _st)


def _check_len(ctx, T):
    if T > ctx.cfg["max_seq_len"]:
        raise NotImplementedError("training mode: sequence length %d > max_seq_len %d (the reference truncates the batch, "
                                  "transformer.py:137-145)" % (T, ctx.cfg["max_seq_len"]))


def _embed_abs(ctx, tokens, d):
    B, S = tokens.shape
    _check_len(ctx, S)
    pe = ctx.P["encoder.position_enc"][0]
    return embed_tokens(ctx, tokens, None, "encoder.src_word_emb.weight", 1.0, 1, table=pe)


def _add_abs(ctx, x):
    _check_len(ctx, x.v.shape[1])
    return add_positions(ctx, x, None, None, pos_mode=1, table=ctx.P["decoder.position_enc"][0])


def linear_cat(ctx, x, vname, parts, math, out_planes=0):
    """One GEMM over the concatenation of several [C, C] projections (prep.w[vname], built by engine_blocks.PREPARE): the
    weight gradient of the concatenation is distributed to the parts' arena slots afterwards."""
    cat = ctx.concat(vname, parts)
    holder = {}

    def distribute():
        t = holder.get("g")
        if t is None:
            return
        off = 0
        for name in parts:
            G = ctx.G.get(name)
            n = ctx.P[name].shape[0]
            if G is not None:
                capi.call("ctts_axpy", t[off:off + n], 1.0, G.numel(), 1, G, _st())
            off += n
        ctx.G.pop(vname, None)

    ctx.record(distribute)
    y = linear(ctx, x, vname, math=math, out_planes=out_planes)

    def alloc():
        holder["g"] = ctx.G[vname] = torch.zeros_like(cat)

    ctx.record(alloc)
    return y


# ---------------------------------------------------------------------------------------------------------------------
# transformer (post-LN)
# ---------------------------------------------------------------------------------------------------------------------
def _stack_transformer(ctx, pre, x, lens, n_layers, n_head, kernel, p_drop, math):
    n = {"tc2": 2, "tc3": 3}.get(math, 0)
    for i in range(n_layers):
        a = "%slayer_stack.%d.slf_attn." % (pre, i)
        f = "%slayer_stack.%d.pos_ffn." % (pre, i)
        qkv = linear_cat(ctx, x, a + "qkv", [a + "w_qs.linear.weight", a + "w_ks.linear.weight", a + "w_vs.linear.weight"],
                         math, out_planes=n)
        att = TE.attention(ctx, qkv, lens, n_head, math)
        o = sublayer(ctx, x, att, a + "fc.linear.weight", None, None, p_drop, math)
        x = layer_norm(ctx, o, a + "layer_norm.weight", a + "layer_norm.bias", 1e-5, lens, planes=n)
        h = linear(ctx, x, f + "w_1.weight", f + "w_1.bias", act=ACT_RELU, taps=kernel[0], math=math, out_planes=n)
        o = sublayer(ctx, x, h, f + "w_2.weight", f + "w_2.bias", None, p_drop, math, taps=kernel[1])
        x = layer_norm(ctx, o, f + "layer_norm.weight", f + "layer_norm.bias", 1e-5, lens, planes=n)
    return x


def encoder_transformer(ctx, tokens, src_lens):
    c = ctx.cfg["transformer"]
    x, word = _embed_abs(ctx, tokens, c["encoder_hidden"])
    x = _stack_transformer(ctx, "encoder.", x, src_lens, c["encoder_layer"], c["encoder_head"], c["conv_kernel_size"],
                           c["encoder_dropout"], ctx.enc_math)
    return x, word


def decoder_transformer(ctx, x, mel_lens):
    c = ctx.cfg["transformer"]
    x = _add_abs(ctx, x)
    return _stack_transformer(ctx, "decoder.", x, mel_lens, c["decoder_layer"], c["decoder_head"], c["conv_kernel_size"],
                              c["decoder_dropout"], ctx.dec_math)


# ---------------------------------------------------------------------------------------------------------------------
# fastformer
# ---------------------------------------------------------------------------------------------------------------------
def _pool(ctx, logits, values, lens, heads, hs):
    B, T, _ = logits.v.shape
    pooled = torch.empty(B, heads * hs, device=logits.v.device, dtype=torch.float32)
    capi.call("ctts_fastformer_pool", logits.v, values.v, lens, B, T, heads, hs, pooled, _st())
    y = Var(pooled)

    def bwd():
        if y.g is None:
            return
        dl = torch.empty_like(logits.v)
        dv = torch.empty_like(values.v)
        capi.call("ctts_fastformer_pool_bwd", logits.v, values.v, lens, y.g, B, T, heads, hs, dl, dv, _st())
        accumulate_into(logits, dl)
        accumulate_into(values, dv)
        y.g = None

    ctx.record(bwd)
    return y


def _mul_row(ctx, a, row):
    """a[b, t, :] * row[b, :]"""
    B, T, C = a.v.shape
    out = torch.empty_like(a.v)
    capi.call("ctts_binary", a.v, row.v, 1, 1, None, B, T, C, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        da = torch.empty_like(a.v)
        db = torch.zeros_like(row.v)
        capi.call("ctts_mul_bwd", y.g, a.v, row.v, 1, None, B, T, C, da, db, _st())
        accumulate_into(a, da)
        accumulate_into(row, db)
        y.g = None

    ctx.record(bwd)
    return y


def _stack_fastformer(ctx, pre, x, lens, n_layers, heads, kernel, p_drop, math):
    C = x.v.shape[-1]
    hs = C // heads
    n = {"tc2": 2, "tc3": 3}.get(math, 0)
    tied = "%slayer_stack.layers.0.0.fn." % pre
    for i in range(n_layers):
        a = "%slayer_stack.layers.%d.0." % (pre, i)
        f = "%slayer_stack.layers.%d.1." % (pre, i)
        h = layer_norm(ctx, x, a + "norm.weight", a + "norm.bias", 1e-5, planes=n)
        q = linear(ctx, h, a + "fn.query.weight", a + "fn.query.bias", math=math, out_planes=n)
        k = linear(ctx, h, a + "fn.key.weight", a + "fn.key.bias", math=math)
        ql = linear(ctx, q, tied + "to_q_attn_logits.weight", tied + "to_q_attn_logits.bias", math=math)
        pooled_q = _pool(ctx, ql, q, lens, heads, hs)
        qk = _mul_row(ctx, k, pooled_q)
        kl = linear(ctx, qk, tied + "to_k_attn_logits.weight", tied + "to_k_attn_logits.bias", math=math)
        pooled_k = _pool(ctx, kl, qk, lens, heads, hs)
        wv = _mul_row(ctx, q, pooled_k)
        if p_drop > 0.0 and ctx.dropout_on:      # dropout(transform(wv) + q) + x
            t = add(ctx, linear(ctx, wv, a + "fn.transform.weight", a + "fn.transform.bias", math=math), q)
            x = TE.residual_add(ctx, x, dropout(ctx, t, p_drop), lens)
        else:
            r = add(ctx, q, x)
            x = linear(ctx, wv, a + "fn.transform.weight", a + "fn.transform.bias", residual=r, lens=lens, math=math)
        h = layer_norm(ctx, x, f + "norm.weight", f + "norm.bias", 1e-5, planes=n)
        g = linear(ctx, h, f + "fn.w_1.weight", f + "fn.w_1.bias", act=ACT_GELU, taps=kernel[0], math=math, out_planes=n)
        x = sublayer(ctx, x, g, f + "fn.w_2.weight", f + "fn.w_2.bias", lens, p_drop, math, taps=kernel[1])
    return x


def encoder_fastformer(ctx, tokens, src_lens):
    c = ctx.cfg["transformer"]
    x, word = _embed_abs(ctx, tokens, c["encoder_hidden"])
    heads = c["encoder_hidden"] // c["encoder_head"]
    return _stack_fastformer(ctx, "encoder.", x, src_lens, c["encoder_layer"], heads, c["conv_kernel_size"],
                             c["encoder_dropout"], ctx.enc_math), word


def decoder_fastformer(ctx, x, mel_lens):
    c = ctx.cfg["transformer"]
    x = _add_abs(ctx, x)
    heads = c["decoder_hidden"] // c["decoder_head"]
    return _stack_fastformer(ctx, "decoder.", x, mel_lens, c["decoder_layer"], heads, c["conv_kernel_size"],
                             c["decoder_dropout"], ctx.dec_math)


# ---------------------------------------------------------------------------------------------------------------------
# conformer
# ---------------------------------------------------------------------------------------------------------------------
def _add_bias_row(ctx, x, bname):
    """x[r, :] + bias (the u / v biases of the relative attention, broadcast over batch and time)."""
    B, T, C = x.v.shape
    out = torch.empty_like(x.v)
    capi.call("ctts_add_row_broadcast", x.v, ctx.P[bname], 1, B * T, C, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        G = ctx.G.get(bname)
        if G is not None:
            capi.call("ctts_act_bwd", y.g, None, ACT_NONE, 1.0, None, 1, T, B * T, C, None, G, _st())
        accumulate_into(x, y.g)
        y.g = None

    ctx.record(bwd)
    return y


def _relpos_attention(ctx, a, qkv, pos_proj, n_head, p_drop):
    """RelativeMultiHeadAttention.forward (conformer.py:397-431): content + shifted positional scores, softmax WITHOUT
    padding mask, (dropout), P.V.  Scores are materialised in FP32 ([B*H, T, T]), like the inference path."""
    B, T, C3 = qkv.v.shape
    C = C3 // 3
    dh = C // n_head
    Z = B * n_head
    dev = qkv.v.device
    st = _st()
    qc = torch.empty(B, T, C, device=dev, dtype=torch.float32)
    capi.call("ctts_copy_rows", qkv.v, C3, B * T, C, qc, C, 0, st)
    q = Var(qc)
    qu = _add_bias_row(ctx, q, a + "attention.u_bias")
    qv = _add_bias_row(ctx, q, a + "attention.v_bias")
    kk = qkv.v.view(-1)[C:]
    vv = qkv.v.view(-1)[2 * C:]
    TT = T * T
    hs_q = (T * C, dh, C, 1, 0)            # (zo, zi, row, k, kb) of a head slice of a [B, T, C] tensor
    hs_kv = (T * C3, dh, C3, 1, 0)         # ... of the k / v part of qkv
    sc = (n_head * TT, TT, T, 1)
    content = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
    pscore = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
    _generic(qu.v, kk, content, Z, n_head, T, T, dh, hs_q, hs_kv, sc)
    _generic(qv.v, pos_proj.v, pscore, Z, n_head, T, T, dh, hs_q, (0, dh, C, 1, 0), sc)
    prob = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
    capi.call("ctts_relshift_softmax", content, pscore, Z, T, T, math.sqrt(C), prob, st)
    drop_off = None
    prob_used = prob
    if p_drop > 0.0 and ctx.dropout_on:
        drop_off = ctx.next_offset()
        prob_used = torch.empty_like(prob)
        capi.call("ctts_dropout", prob, prob.numel(), float(p_drop), ctx.seed, drop_off, ctx.offset_dev, prob_used, st)
    out = torch.empty(B, T, C, device=dev, dtype=torch.float32)
    # ctx[b, t, h*dh + d] = sum_s P[z][t, s] v[b, s, h*dh + d]
    _generic(prob_used, vv, out, Z, n_head, T, dh, T, (n_head * TT, TT, T, 1, 0), (T * C3, dh, 1, C3, 0), (T * C, dh, C, 1))
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        st2 = _st()
        dO = y.g
        dqkv = torch.zeros_like(qkv.v)
        dk, dv = dqkv.view(-1)[C:], dqkv.view(-1)[2 * C:]
        # dP[t, s] = sum_d dO[t, d] v[s, d];  dV[s, d] = sum_t P[t, s] dO[t, d]
        dP = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
        _generic(dO, vv, dP, Z, n_head, T, T, dh, hs_q, hs_kv, sc)
        _generic(prob_used, dO, dv, Z, n_head, T, dh, T, (n_head * TT, TT, 1, T, 0), (T * C, dh, 1, C, 0), (T * C3, dh, C3, 1))
        if drop_off is not None:
            capi.call("ctts_dropout", dP, dP.numel(), float(p_drop), ctx.seed, drop_off, ctx.offset_dev, dP, st2)
        capi.call("ctts_softmax_bwd", prob, dP, Z, T, T, T, 1.0, dP, st2)
        dcontent = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
        dpos = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
        capi.call("ctts_relshift_bwd", dP, Z, T, T, T, math.sqrt(C), dcontent, dpos, st2)
        # content = qu k^T:  dqu[t, d] = sum_s dcontent[t, s] k[s, d];  dk[s, d] = sum_t dcontent[t, s] qu[t, d]
        dqu, acc = grad_buffer(qu)
        _generic(dcontent, kk, dqu, Z, n_head, T, dh, T, (n_head * TT, TT, T, 1, 0), (T * C3, dh, 1, C3, 0), (T * C, dh, C, 1),
                 accumulate=acc)
        _generic(dcontent, qu.v, dk, Z, n_head, T, dh, T, (n_head * TT, TT, 1, T, 0), (T * C, dh, 1, C, 0), (T * C3, dh, C3, 1))
        # pscore = qv pos^T:  dqv[t, d] = sum_j dpos[t, j] pos[j, h, d];  dpos_proj[j, h, d] = sum_{b,t} dpos[b,h][t, j] qv[b,t,h,d]
        dqv, acc = grad_buffer(qv)
        _generic(dpos, pos_proj.v, dqv, Z, n_head, T, dh, T, (n_head * TT, TT, T, 1, 0), (0, dh, 1, C, 0), (T * C, dh, C, 1),
                 accumulate=acc)
        dpp, acc = grad_buffer(pos_proj)
        _generic(dpos, qv.v, dpp, n_head, 1, T, dh, B * T, (TT, 0, 1, T, n_head * TT), (dh, 0, 1, C, T * C), (dh, 0, C, 1),
                 Kin=T, accumulate=acc)
        accumulate_into(qkv, dqkv)
        y.g = None

    ctx.record(bwd)

    def bwd_q():          # recorded after the consumers of q: runs before them?  No -- see below
        pass

    # q is a private copy of the first third of qkv: its gradient is routed back by a closure recorded BEFORE q's
    # consumers (so that it runs after them in the backward pass)
    return y, q


def _relpos_attention_tc(ctx, a, qkv, pos_proj, n_head, p_drop):
    """The same attention on tcgen05 for the training step (2 bf16 planes): the 32-wide heads are zero padded to one
    64-wide k-block (ctts_pad_heads_planes); content / positional scores, P.V and all six backward products are batched plane
    GEMMs (ctts_gemm_batched_planes); the operands whose reduction index is time come from ctts_split_transpose.  The
    [B*H, T, T] score tensors are materialised in fp32, as in the reference."""
    from .engine import Planes, gemm_batched_planes, split_planes
    B, T, C3 = qkv.v.shape
    C = C3 // 3
    H = n_head
    dh = C // H
    Z = B * H
    DHp = 64
    Cp = H * DHp
    Tp = (T + 7) // 8 * 8
    dev = qkv.v.device
    st = _st()
    qc = torch.empty(B, T, C, device=dev, dtype=torch.float32)
    capi.call("ctts_copy_rows", qkv.v, C3, B * T, C, qc, C, 0, st)
    q = Var(qc)
    qu = _add_bias_row(ctx, q, a + "attention.u_bias")
    qv = _add_bias_row(ctx, q, a + "attention.v_bias")

    def padded(src, rows, ld, c0):
        out = Planes.empty((rows, Cp), dev, 2)
        capi.call("ctts_pad_heads_planes", src, None, rows, ld, c0, H, dh, DHp, 2, capi.ptr_array(out.p), _st())
        return out

    def time_major(src, Zs, ld, c0, cols):
        """fp32 [Zs, T, ld] columns [c0, c0+cols) -> planes [Zs, cols, Tp]"""
        out = Planes.empty((Zs, 1, cols, Tp), dev, 2)
        capi.call("ctts_split_transpose", src, Zs, T, cols, ld, c0, Tp, 1, 2, capi.ptr_array(out.p), _st())
        return out

    act_view = (Cp, T, B, Cp, T * Cp)
    big = (H * T * Tp, T * Tp)
    sq_view = (T, T, Z, Tp, T * Tp)          # [Z][T rows][T valid of Tp columns]
    hT_view = (T, dh, Z, Tp, dh * Tp)        # [Z][dh rows][time]
    kp = padded(qkv.v, B * T, C3, C)
    content = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
    pscore = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
    gemm_batched_planes(padded(qu.v, B * T, C, 0), act_view, kp, act_view, (H, H, 0, DHp, H, 0, DHp, 1, Tp), big[0], big[1], 1.0,
                        Z, T, DHp, Tp, y=content)
    gemm_batched_planes(padded(qv.v, B * T, C, 0), act_view, padded(pos_proj.v, T, C, 0), (Cp, T, 1, Cp, T * Cp),
                        (H, H, 0, DHp, 0x7fffffff, 0, DHp, 1, Tp), big[0], big[1], 1.0, Z, T, DHp, Tp, y=pscore)
    prob = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
    if Tp == T:
        capi.call("ctts_relshift_softmax", content, pscore, Z, T, Tp, math.sqrt(C), prob, st)
    else:
        # the fp32 kernel reads dense [Z, T, T] score rows: re-stride first (T not a multiple of 8)
        cd = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
        pd = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
        capi.call("ctts_copy_rows", content, Tp, Z * T, T, cd, T, 0, st)
        capi.call("ctts_copy_rows", pscore, Tp, Z * T, T, pd, T, 0, st)
        capi.call("ctts_relshift_softmax", cd, pd, Z, T, Tp, math.sqrt(C), prob, st)
    del content, pscore
    drop_off = None
    prob_used = prob
    if p_drop > 0.0 and ctx.dropout_on:
        drop_off = ctx.next_offset()
        prob_used = torch.empty_like(prob)
        capi.call("ctts_dropout", prob, prob.numel(), float(p_drop), ctx.seed, drop_off, ctx.offset_dev, prob_used, st)
    out = torch.empty(B, T, C, device=dev, dtype=torch.float32)
    gemm_batched_planes(split_planes(prob_used, 2), sq_view, time_major(qkv.v, B, C3, 2 * C, C), hT_view,
                        (H, 1, 0, 0, 1, 0, 0, 1, C), T * C, dh, 1.0, Z, T, T, dh, y=out)
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        st2 = _st()
        dO = y.g
        dqkv = torch.zeros_like(qkv.v)
        addr_c = (H, 1, 0, 0, 1, 0, 0, 1, C)         # outputs laid out like a [B, T, C] activation
        addr_3c = (H, 1, 0, 0, 1, 0, 0, 1, C3)        # ... like a third of qkv
        # dP[t, s] = sum_d dO[t, d] v[s, d]   (padded heads, K = 64)
        dP = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
        gemm_batched_planes(padded(dO, B * T, C, 0), act_view, padded(qkv.v, B * T, C3, 2 * C), act_view,
                            (H, H, 0, DHp, H, 0, DHp, 1, Tp), big[0], big[1], 1.0, Z, T, DHp, Tp, y=dP)
        # dV[s, d] = sum_t P[t, s] dO[t, d]
        PT = Planes.empty((Z, 1, T, Tp), dev, 2)
        capi.call("ctts_split_transpose", prob_used, Z, T, T, Tp, 0, Tp, 1, 2, capi.ptr_array(PT.p), st2)
        gemm_batched_planes(PT, sq_view, time_major(dO, B, C, 0, C), hT_view, addr_3c, T * C3, dh, 1.0, Z, T, T, dh,
                            y=dqkv.view(-1)[2 * C:])
        if drop_off is not None:
            capi.call("ctts_dropout", dP, dP.numel(), float(p_drop), ctx.seed, drop_off, ctx.offset_dev, dP, st2)
        capi.call("ctts_softmax_bwd", prob, dP, Z, T, T, Tp, 1.0, dP, st2)
        dcontent = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
        dpos = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
        capi.call("ctts_relshift_bwd", dP, Z, T, Tp, Tp, math.sqrt(C), dcontent, dpos, st2)
        dcp = split_planes(dcontent, 2)
        dcT = Planes.empty((Z, 1, T, Tp), dev, 2)
        capi.call("ctts_split_transpose", dcontent, Z, T, T, Tp, 0, Tp, 1, 2, capi.ptr_array(dcT.p), st2)
        # content = qu k^T:  dqu[t, d] = sum_s dcontent[t, s] k[s, d];  dk[s, d] = sum_t dcontent[t, s] qu[t, d]
        dqu, acc = grad_buffer(qu)
        gemm_batched_planes(dcp, sq_view, time_major(qkv.v, B, C3, C, C), hT_view, addr_c, T * C, dh, 1.0, Z, T, T, dh, y=dqu,
                            residual=dqu if acc else None)
        gemm_batched_planes(dcT, sq_view, time_major(qu.v, B, C, 0, C), hT_view, addr_3c, T * C3, dh, 1.0, Z, T, T, dh,
                            y=dqkv.view(-1)[C:])
        # pscore = qv pos^T:  dqv[t, d] = sum_j dpos[t, j] pos[j, h, d];  dpos_proj[j, h, d] = sum_{b,t} dpos[b,h][t, j] qv[b,t,h,d]
        dpp_ = split_planes(dpos, 2)
        pos_rep = torch.empty(B, T, C, device=dev, dtype=torch.float32)
        capi.call("ctts_copy_rows", pos_proj.v, 0, B, T * C, pos_rep, T * C, 0, st2)     # one copy per utterance
        dqv, acc = grad_buffer(qv)
        gemm_batched_planes(dpp_, sq_view, time_major(pos_rep, B, C, 0, C), hT_view, addr_c, T * C, dh, 1.0, Z, T, T, dh, y=dqv,
                            residual=dqv if acc else None)
        dposT = Planes.empty((Z, 1, T, Tp), dev, 2)
        capi.call("ctts_split_transpose", dpos, Z, T, T, Tp, 0, Tp, 1, 2, capi.ptr_array(dposT.p), st2)
        part = torch.empty(B, T, C, device=dev, dtype=torch.float32)                     # per-utterance partial sums
        gemm_batched_planes(dposT, sq_view, time_major(qv.v, B, C, 0, C), hT_view, addr_c, T * C, dh, 1.0, Z, T, T, dh, y=part)
        if pos_proj.needs_grad:
            if pos_proj.g is None:
                pos_proj.g = torch.zeros_like(pos_proj.v)
            capi.call("ctts_act_bwd", part, None, ACT_NONE, 1.0, None, 1, 1, B, T * C, None, pos_proj.g, st2)   # sum over b
        accumulate_into(qkv, dqkv)
        y.g = None

    ctx.record(bwd)
    return y, q


def _route_q(ctx, qkv, holder):
    """Recorded before the relative attention: adds dq (the gradient of the private q copy) into dqkv[:, :, :C]."""
    def bwd():
        q = holder.get("q")
        if q is None or q.g is None or qkv.g is None:
            return
        B, T, C3 = qkv.v.shape
        C = C3 // 3
        capi.call("ctts_copy_rows", q.g, C, B * T, C, qkv.g, C3, 1, _st())
        q.g = None

    ctx.record(bwd)


def _glu(ctx, h):
    B, T, C2 = h.v.shape
    C = C2 // 2
    out = torch.empty(B, T, C, device=h.v.device, dtype=torch.float32)
    capi.call("ctts_glu", h.v, B * T, C, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        dh = torch.empty_like(h.v)
        capi.call("ctts_glu_bwd", h.v, y.g, B * T, C, dh, _st())
        accumulate_into(h, dh)
        y.g = None

    ctx.record(bwd)
    return y


def _dwconv(ctx, x, wname, K):
    B, T, C = x.v.shape
    out = torch.empty_like(x.v)
    w = ctx.P[wname]
    capi.call("ctts_dwconv", x.v, w, K, B, T, C, out, _st())
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        dx = torch.empty_like(x.v)
        capi.call("ctts_dwconv_bwd", y.g, x.v, w, K, B, T, C, dx, ctx.G.get(wname), _st())
        accumulate_into(x, dx)
        y.g = None

    ctx.record(bwd)
    return y


def _conformer_ffn(ctx, p, x, p_drop, math):
    """FeedForwardModule with the half-step residual (conformer.py:205-213,264-295)."""
    n = {"tc2": 2, "tc3": 3}.get(math, 0)
    h = layer_norm(ctx, x, p + "0.weight", p + "0.bias", 1e-5, planes=n)
    g = linear(ctx, h, p + "1.linear.weight", p + "1.linear.bias", act=ACT_SWISH, math=math, out_planes=n)
    g = dropout(ctx, g, p_drop)
    return sublayer(ctx, x, g, p + "4.linear.weight", p + "4.linear.bias", None, p_drop, math, alpha=0.5)


def _stack_conformer(ctx, pre, x, lens, n_layers, n_head, kernel, p_drop, math):
    B, T, C = x.v.shape
    dev = x.v.device
    for i in range(n_layers):
        lp = "%slayer_stack.%d.sequential." % (pre, i)
        x = _conformer_ffn(ctx, lp + "0.module.sequential.", x, p_drop, math)
        a = lp + "1.module."
        h = layer_norm(ctx, x, a + "layer_norm.weight", a + "layer_norm.bias", 1e-5)
        qkv = linear_cat(ctx, h, a + "attention.qkv", [a + "attention.query_proj.linear.weight",
                                                       a + "attention.key_proj.linear.weight",
                                                       a + "attention.value_proj.linear.weight"], math)
        pos = Var(ctx.P[a + "positional_encoding"][0, :T].contiguous().view(1, T, C), False)
        pos_proj = linear(ctx, pos, a + "attention.pos_proj.linear.weight")
        holder = {}
        _route_q(ctx, qkv, holder)
        rel = _relpos_attention_tc if (ctx.bwd_tc and C // n_head <= 64 and C % n_head == 0) else _relpos_attention
        att, q = rel(ctx, a, qkv, TE._reshape(ctx, pos_proj, (T, C)), n_head, p_drop)
        holder["q"] = q
        x = sublayer(ctx, x, att, a + "attention.out_proj.linear.weight", None, None, p_drop, math)
        m = lp + "2.module.sequential."
        h = layer_norm(ctx, x, m + "0.weight", m + "0.bias", 1e-5)
        pw = linear(ctx, h, m + "2.conv.weight", m + "2.conv.bias", math=math)
        g = _glu(ctx, pw)
        d = _dwconv(ctx, g, m + "4.conv.weight", kernel)
        bn = batch_norm_act(ctx, d, m + "5.", ACT_SWISH)
        x = sublayer(ctx, x, bn, m + "7.conv.weight", m + "7.conv.bias", None, p_drop, math)
        x = _conformer_ffn(ctx, lp + "3.module.sequential.", x, p_drop, math)
        x = layer_norm(ctx, x, lp + "4.weight", lp + "4.bias", 1e-5, lens)
    return x


def encoder_conformer(ctx, tokens, src_lens):
    c = ctx.cfg["conformer"]
    x, word = _embed_abs(ctx, tokens, c["encoder_hidden"])
    return _stack_conformer(ctx, "encoder.", x, src_lens, c["encoder_layer"], c["encoder_head"], c["conv_kernel_size"],
                            c["encoder_dropout"], ctx.enc_math), word


def decoder_conformer(ctx, x, mel_lens):
    c = ctx.cfg["conformer"]
    x = _add_abs(ctx, x)
    return _stack_conformer(ctx, "decoder.", x, mel_lens, c["decoder_layer"], c["decoder_head"], c["conv_kernel_size"],
                            c["decoder_dropout"], ctx.dec_math)


TE.ENCODERS.update({"transformer": encoder_transformer, "fastformer": encoder_fastformer, "conformer": encoder_conformer})
TE.DECODERS.update({"transformer": decoder_transformer, "fastformer": decoder_fastformer, "conformer": decoder_conformer})
