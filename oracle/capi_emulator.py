"""CPU emulation of libctts_b200's C ABI, entry point by entry point.  TEST INFRASTRUCTURE -- NOT PRODUCT.

Each function restates, in plain torch on CPU tensors, the contract `include/ctts_b200.h` gives for the entry point of the
same name (same argument order).  Two uses, both in `tests/` only:
  * `-m "not gpu"`: `install(monkeypatch)` replaces `capi.call`, so the HOST logic of the product (the tape of
    ctts_b200/train_engine.py, shapes, strides, gradient routing) runs in this GPU-less container and is checked against the
    reference's gradient fixtures (tests/golden/*_train.npz);
  * `-m gpu`: the per-kernel parity tests call the real entry point and this restatement on the same seeded inputs.
Pointer semantics are kept: a tensor argument stands for the address of its first element, and the callee addresses
`ptr + offset` into the underlying storage (`_flat`), so sliced / offset views behave as they do on the device.
The product never imports this module (tests/test_capi_symbols.py::test_product_never_imports_oracle).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_GELU, ACT_TANH, ACT_SWISH = 0, 1, 2, 3, 4


def _flat(t):
    """1-D alias of t's storage starting at t's first element (pointer semantics)."""
    n = t.untyped_storage().nbytes() // t.element_size() - t.storage_offset()
    return t.as_strided((n,), (1,), t.storage_offset())


def _v(t, *shape):
    """View of the memory at `t` as a contiguous tensor of `shape`."""
    n = int(np.prod(shape)) if shape else 1
    return _flat(t)[:n].view(*shape)


def _planes(arr, n):
    return list(arr._keepalive[:n])


def _val(arr, n, *shape):
    return sum(_v(p, *shape).float() for p in _planes(arr, n))


def _split_into(x, planes, *shape):
    rem = x.float().clone()
    for p in planes:
        h = rem.to(torch.bfloat16)
        _v(p, *shape).copy_(h)
        rem = rem - h.float()


def _act(v, act):
    if act == ACT_RELU:
        return F.relu(v)
    if act == ACT_GELU:
        return F.gelu(v)
    if act == ACT_TANH:
        return torch.tanh(v)
    if act == ACT_SWISH:
        return v * torch.sigmoid(v)
    return v


def _act_grad(ref, act):
    if act == ACT_RELU:
        return (ref > 0).float()
    if act == ACT_TANH:
        return 1 - ref * ref
    if act == ACT_GELU:
        return 0.5 * (1 + torch.erf(ref / math.sqrt(2))) + ref * torch.exp(-0.5 * ref * ref) / math.sqrt(2 * math.pi)
    if act == ACT_SWISH:
        s = torch.sigmoid(ref)
        return s * (1 + ref * (1 - s))
    return torch.ones_like(ref)


def _keep(lens, B, T):
    if lens is None:
        return torch.ones(B, T, dtype=torch.bool)
    return torch.arange(T)[None, :] < _v(lens, B)[:, None]


def _positions(flag):
    keep = flag.int()
    return (torch.cumsum(keep, 1) * keep).long()


# ---- forward entry points -------------------------------------------------------------------------------------------
def ctts_embed_tokens(tokens, table, pe, pe_rows, scale, B, S, C, vocab, x, word, lens, pos_mode, stream):
    tok = _v(tokens, B, S).clamp(0, vocab - 1)
    w = scale * _v(table, vocab, C)[tok]
    pos = _positions(tok != 0) if pos_mode == 0 else torch.arange(S)[None].expand(B, S)
    _v(word, B, S, C).copy_(w)
    _v(x, B, S, C).copy_((w + _v(pe, pe_rows, C)[pos]) * _keep(lens, B, S)[:, :, None])


def ctts_add_positions(x, pe, pe_rows, alpha, lens, B, T, C, pos_mode, y, stream):
    xv = _v(x, B, T, C)
    pos = _positions(xv[..., 0] != 0) if pos_mode == 0 else torch.arange(T)[None].expand(B, T)
    a = _v(alpha, 1)[0] if alpha is not None else 1.0
    _v(y, B, T, C).copy_((xv + a * _v(pe, pe_rows, C)[pos]) * _keep(lens, B, T)[:, :, None])


def _ln(x, gamma, beta, eps, lens, B, T, C):
    y = F.layer_norm(_v(x, B, T, C), (C,), _v(gamma, C), _v(beta, C), eps)
    return y * _keep(lens, B, T)[:, :, None]


def ctts_layernorm(x, gamma, beta, eps, lens, B, T, C, y, stream):
    _v(y, B, T, C).copy_(_ln(x, gamma, beta, eps, lens, B, T, C))


def ctts_layernorm_planes(x, gamma, beta, eps, lens, B, T, C, y, n, planes, stream):
    out = _ln(x, gamma, beta, eps, lens, B, T, C)
    if y is not None:
        _v(y, B, T, C).copy_(out)
    if n:
        _split_into(out, _planes(planes, n), B, T, C)


def _conv_core(xv, wv, bias, alpha, cs, csh, act, residual, lens, B, T, Cin, N, taps):
    w = wv.view(N, taps, Cin).permute(0, 2, 1).contiguous()          # packed [N, taps*Cin] -> torch [N, Cin, taps]
    acc = F.conv1d(xv.transpose(1, 2), w, None, padding=taps // 2).transpose(1, 2)
    if bias is not None:
        acc = acc + _v(bias, N)
    v = acc * alpha
    if cs is not None:
        v = v * _v(cs, N) + _v(csh, N)
    v = _act(v, act)
    if residual is not None:
        v = v + _v(residual, B, T, N)
    return v * _keep(lens, B, T)[:, :, None]


def ctts_conv1d_gemm(x, w, bias, alpha, cs, csh, act, residual, lens, B, T, Cin, N, taps, y, stream):
    out = _conv_core(_v(x, B, T, Cin), _v(w, N, taps * Cin), bias, alpha, cs, csh, act, residual, lens, B, T, Cin, N, taps)
    _v(y, B, T, N).copy_(out)


def ctts_gemm_split(n, xp, wp, bias, alpha, cs, csh, act, residual, lens, B, T, Cin, N, taps, y, yp, stream):
    out = _conv_core(_val(xp, n, B, T, Cin), _val(wp, n, N, taps * Cin), bias, alpha, cs, csh, act, residual, lens, B, T, Cin,
                     N, taps)
    if y is not None:
        _v(y, B, T, N).copy_(out)
    if yp is not None:
        _split_into(out, _planes(yp, n), B, T, N)


def ctts_gemm_split_ln(xp, wp, bias, alpha, residual, lens, B, T, Cin, N, taps, y, g, b, eps, ln_masked, ln_y, ln_planes, stream):
    out = _conv_core(_val(xp, 2, B, T, Cin), _val(wp, 2, N, taps * Cin), bias, alpha, None, None, 0, residual, lens, B, T, Cin,
                     N, taps)
    _v(y, B, T, N).copy_(out)
    ctts_layernorm_planes(y, g, b, eps, lens if ln_masked else None, B, T, N, ln_y, 2, ln_planes, stream)


def ctts_split_planes(x, numel, n, planes, stream):
    _split_into(_v(x, numel), _planes(planes, n), numel)


def ctts_pack_conv_weight(w, N, Cin, taps, packed, stream):
    _v(packed, N, taps, Cin).copy_(_v(w, N, Cin, taps).permute(0, 2, 1))


def _attention(qkv, lens, B, T, C, H, scale):
    dh = C // H
    q, k, v = qkv.split(C, dim=-1)
    q = q.view(B, T, H, dh).transpose(1, 2) * scale
    k = k.view(B, T, H, dh).transpose(1, 2)
    v = v.view(B, T, H, dh).transpose(1, 2)
    keep = _keep(lens, B, T)
    s = torch.matmul(q, k.transpose(-1, -2)).masked_fill(~keep[:, None, None, :], float("-inf"))
    o = torch.matmul(torch.softmax(s, -1), v).transpose(1, 2).reshape(B, T, C)
    return o * keep[:, :, None]


def ctts_attention(qkv, lens, B, T, C, H, scale, out, stream):
    _v(out, B, T, C).copy_(_attention(_v(qkv, B, T, 3 * C), lens, B, T, C, H, scale))


def ctts_transpose_v_planes(n, qkvp, B, T, C, H, vtp, stream):
    dh = C // H
    Tp = (T + 7) // 8 * 8
    for src, dst in zip(_planes(qkvp, n), _planes(vtp, n)):
        v = _v(src, B, T, 3 * C)[:, :, 2 * C:].reshape(B, T, H, dh).permute(0, 2, 3, 1)        # [B, H, dh, T]
        out = _v(dst, B * H, dh, Tp)
        out.zero_()
        out[:, :, :T] = v.reshape(B * H, dh, T)


def ctts_flash_attention_bf16x3(qh, ql, vth, vtl, lens, B, T, C, H, scale, oh, ol, stream):
    qkv = _v(qh, B, T, 3 * C).float() + _v(ql, B, T, 3 * C).float()
    _split_into(_attention(qkv, lens, B, T, C, H, scale), [oh, ol], B, T, C)


def ctts_attention_split(n, qkvp, lens, B, T, C, H, scale, scores, pp, vt, outp, out_f32, stream):
    o = _attention(_val(qkvp, n, B, T, 3 * C), lens, B, T, C, H, scale)
    if outp is not None:
        _split_into(o, _planes(outp, n), B, T, C)
    if out_f32 is not None:
        _v(out_f32, B, T, C).copy_(o)


def ctts_attention_small(qkvp, lens, B, T, C, H, scale, outp, stream):
    assert T <= 128 and C == H * 128
    _split_into(_attention(_val(qkvp, 3, B, T, 3 * C), lens, B, T, C, H, scale), _planes(outp, 3), B, T, C)


class _PA:
    """Stand-in of capi.ptr_array for the legacy two-plane entry points (hi / lo passed as separate pointers)."""

    def __init__(self, *tensors):
        self._keepalive = list(tensors)


def ctts_gemm_bf16x3(x_hi, x_lo, w_hi, w_lo, bias, alpha, cs, csh, act, residual, lens, B, T, Cin, N, taps, y, y_hi, y_lo, stream):
    ctts_gemm_split(2, _PA(x_hi, x_lo), _PA(w_hi, w_lo), bias, alpha, cs, csh, act, residual, lens, B, T, Cin, N, taps, y,
                    _PA(y_hi, y_lo) if y_hi is not None else None, stream)


def ctts_attention_bf16x3(qkv_hi, qkv_lo, lens, B, T, C, H, scale, scores, p_hi, p_lo, vt_hi, vt_lo, out_hi, out_lo, out_f32,
                          stream):
    ctts_attention_split(2, _PA(qkv_hi, qkv_lo), lens, B, T, C, H, scale, scores, _PA(p_hi, p_lo), _PA(vt_hi, vt_lo),
                         _PA(out_hi, out_lo) if out_hi is not None else None, out_f32, stream)


def ctts_split_bf16(x, n, hi, lo, stream):
    ctts_split_planes(x, n, 2, _PA(hi, lo), stream)


def ctts_layernorm_split(x, gamma, beta, eps, lens, B, T, C, y, y_hi, y_lo, stream):
    ctts_layernorm_planes(x, gamma, beta, eps, lens, B, T, C, y, 2, _PA(y_hi, y_lo), stream)


def ctts_decode_durations(log_d, d_control, n, dur, stream):
    _v(dur, n).copy_(torch.clamp(torch.round(torch.exp(_v(log_d, n)) - 1) * d_control, min=0))


def ctts_length_scan(dur_f, dur_i, src_lens, B, S, cum_lr, cum_m2p, mel_len, stream):
    if dur_f is not None:
        d = _v(dur_f, B, S)
        reps = d.double().trunc().clamp(min=0).long()
        rnd = torch.round(d).long()
    else:
        d = _v(dur_i, B, S)
        reps, rnd = d.clamp(min=0), d.clone()
    rnd = rnd * _keep(src_lens, B, S).long()
    c1, c2 = torch.cumsum(reps, 1), torch.cumsum(rnd, 1)
    _v(cum_lr, B, S).copy_(c1.int())
    _v(cum_m2p, B, S).copy_(c2.int())
    ml = _v(mel_len, 2 * B)
    ml[:B] = c1[:, -1]
    ml[B:] = c2[:, -1]


def ctts_length_expand(src, table, row_index, cum_lr, B, S, C, M, accumulate, out, cum_m2p, mel2ph, M2, stream):
    cum = _v(cum_lr, B, S).long()
    t = torch.arange(M)
    j = torch.searchsorted(cum, t[None].expand(B, M).contiguous(), right=True)       # first j with cum[j] > t
    valid = j < S
    jc = j.clamp(max=S - 1)
    if table is not None:
        rows = _v(row_index, B, S).gather(1, jc)
        vals = _flat(table).view(-1, C)[rows]
    else:
        vals = _v(src, B, S, C).gather(1, jc[:, :, None].expand(B, M, C))
    vals = vals * valid[:, :, None]
    o = _v(out, B, M, C)
    o.copy_(o + vals if accumulate else vals)
    if mel2ph is not None and M2 > 0:
        cm = _v(cum_m2p, B, S).long()
        t2 = torch.arange(M2)
        j2 = torch.searchsorted(cm, t2[None].expand(B, M2).contiguous(), right=True)
        _v(mel2ph, B, M2).copy_(torch.where(j2 < S, j2 + 1, torch.zeros_like(j2)))


def _f0_to_coarse(f0):
    mel_min = 1127 * np.log(1 + 50.0 / 700)
    mel_max = 1127 * np.log(1 + 1100.0 / 700)
    m = 1127 * (1 + f0 / 700).log()
    m = torch.where(m > 0, (m - mel_min) * 254 / (mel_max - mel_min) + 1, m)
    m = torch.where(m <= 1, torch.ones_like(m), m)
    m = torch.where(m > 255, torch.full_like(m, 255), m)
    return (m + 0.5).long()


def ctts_cwt_to_pitch(cwt, cwt_stride, scale_w, mean, std, stat_stride, std_scale, eps, uv_src, use_uv, B, T, f0_norm,
                      f0_denorm, pitch_idx, stream):
    c = _v(cwt, B, T, cwt_stride)
    rec = (c[..., :10] * _v(scale_w, 10)).sum(-1)
    rec = (rec - rec.mean(-1, keepdim=True)) / rec.std(-1, keepdim=True)
    m = _flat(mean)[: (B - 1) * stat_stride + 1: stat_stride]
    s = _flat(std)[: (B - 1) * stat_stride + 1: stat_stride] * std_scale
    fn = torch.log2((rec * s[:, None] + m[:, None]).exp() + eps)
    fd = 2 ** fn
    if use_uv:
        uv = (_v(uv_src, B, T) > 0) if uv_src is not None else (c[..., 10] > 0)
        fd = torch.where(uv, torch.zeros_like(fd), fd)
    if f0_norm is not None:
        _v(f0_norm, B, T).copy_(fn)
    _v(f0_denorm, B, T).copy_(fd)
    _v(pitch_idx, B, T).copy_(_f0_to_coarse(fd))


def ctts_f0_to_pitch(f0n, uv_src, n, f0_denorm, pitch_idx, stream):
    fd = 2 ** _v(f0n, n)
    if uv_src is not None:
        fd = torch.where(_v(uv_src, n) > 0, torch.zeros_like(fd), fd)
    _v(f0_denorm, n).copy_(fd)
    _v(pitch_idx, n).copy_(_f0_to_coarse(fd))


def ctts_frame_pitch(pred, ldp, f0_target, uv_target, mel2ph, use_uv, n, f0_out, f0_denorm, pitch_idx, stream):
    pr = _v(pred, n, ldp) if pred is not None else None
    f0 = _v(f0_target, n).clone() if f0_target is not None else pr[:, 0].clone()
    uv = torch.zeros(n, dtype=torch.bool)
    if use_uv:
        uv = (_v(uv_target, n) > 0) if uv_target is not None else (pr[:, 1] > 0)
    pad = _v(mel2ph, n) == 0
    fd = torch.where(uv | pad, torch.zeros(n), 2 ** f0)
    _v(f0_out, n).copy_(torch.where(pad, torch.zeros(n), f0))
    if f0_target is None:
        pr[:, 0] = torch.where(pad, torch.zeros(n), pr[:, 0])
    _v(f0_denorm, n).copy_(fd)
    _v(pitch_idx, n).copy_(_f0_to_coarse(fd))


def ctts_gather_index(idx_ph, mel2ph, B, S, M, out, stream):
    padded = F.pad(_v(idx_ph, B, S), [1, 0])
    _v(out, B, M).copy_(torch.gather(padded, 1, _v(mel2ph, B, M).clamp(0, S)))


def ctts_phoneme_pitch(f0, mel2ph, src_lens, mel_lens, B, S, M, out, stream):
    from . import ctts_oracle as O
    r = O.phoneme_level_pitch(None, _v(src_lens, B), _v(mel2ph, B, M), _v(mel_lens, B), _v(f0, B, M))
    o = _v(out, B, S)
    o.zero_()
    o[:, : r.shape[1]] = r


def ctts_gather_add(table, idx, rows, C, table_rows, x, stream):
    xv = _v(x, rows, C)
    xv.copy_(xv + _v(table, table_rows, C)[_v(idx, rows).clamp(0, table_rows - 1)])


def ctts_bucketize(v, v_scale, bins, n_bins, n, idx, stream):
    _v(idx, n).copy_(torch.bucketize(_v(v, n) * v_scale, _v(bins, n_bins)))


def ctts_add_row_broadcast(x, row, B, T, C, y, stream):
    _v(y, B, T, C).copy_(_v(x, B, T, C) + _v(row, B, C)[:, None, :])


def ctts_binary(a, b, op, b_rowwise, lens, B, T, C, y, stream):
    av = _v(a, B, T, C)
    bv = _v(b, B, C)[:, None, :] if b_rowwise else _v(b, B, T, C)
    out = av + bv if op == 0 else av * bv
    _v(y, B, T, C).copy_(out * _keep(lens, B, T)[:, :, None])


def ctts_aligner_attention(q, k, prior, src_lens, temperature, B, M, S, C, soft, logprob, stream):
    qv, kv = _v(q, B, M, C), _v(k, B, S, C)
    a = -temperature * ((qv[:, :, None, :] - kv[:, None, :, :]) ** 2).sum(-1)
    lp = F.log_softmax(a, dim=2) + torch.log(_v(prior, B, S, M).transpose(1, 2) + 1e-8)
    _v(logprob, B, M, S).copy_(lp)
    keep = _keep(src_lens, B, S)
    _v(soft, B, M, S).copy_(torch.softmax(lp.masked_fill(~keep[:, None, :], float("-inf")), dim=2))


def ctts_mas(attn, src_lens, mel_lens, B, M, S, prev_ws, hard, dur, stream):
    from . import ctts_oracle as O
    a = _v(attn, B, 1, M, S)
    h = O.binarize_attention(a, _v(src_lens, B), _v(mel_lens, B))
    _v(hard, B, 1, M, S).copy_(h)
    _v(dur, B, S).copy_(h.sum(2)[:, 0, :])


def ctts_phoneme_energy(dur, src_lens, energy, B, S, M, work, out, stream):
    from . import ctts_oracle as O
    e = O.phoneme_level_energy(_v(dur, B, S), _v(src_lens, B), _v(energy, B, M))
    o = _v(out, B, S)
    o.zero_()
    o[:, : e.shape[1]] = e[:, :S]


def ctts_batched_gemm_fp32(x, w, alpha, lens, lens_div, Z, mod, T, K, N, x_so, x_sh, x_ld, w_so, w_sh, w_ld, y_so, y_sh,
                           y_ld, y, stream):
    xf, wf, yf = _flat(x), _flat(w), _flat(y)
    for z in range(Z):
        zo, zh = divmod(z, mod)
        A = xf.as_strided((T, K), (x_ld, 1), xf.storage_offset() + zo * x_so + zh * x_sh)
        Bm = wf.as_strided((N, K), (w_ld, 1), wf.storage_offset() + zo * w_so + zh * w_sh)
        out = alpha * (A @ Bm.t())
        if lens is not None:
            out = out * (torch.arange(T) < int(_v(lens, (Z - 1) // max(lens_div, 1) + 1)[z // max(lens_div, 1)]))[:, None]
        yf.as_strided((T, N), (y_ld, 1), yf.storage_offset() + zo * y_so + zh * y_sh).copy_(out)


def ctts_transpose_heads(x, B, T, ld_in, c0, H, DH, ldt, xt, stream):
    xv = _v(x, B, T, ld_in)[:, :, c0:c0 + H * DH].reshape(B, T, H, DH).permute(0, 2, 3, 1)     # [B,H,DH,T]
    o = _v(xt, B * H, DH, ldt)
    o.zero_()
    o[:, :, :T] = xv.reshape(B * H, DH, T)


def ctts_fastformer_pool(logits, values, lens, B, T, heads, hs, pooled, stream):
    lg = _v(logits, B, T, heads) / math.sqrt(hs)
    keep = _keep(lens, B, T)
    s = lg + torch.where(keep, torch.full((), -10000.0), torch.zeros(()))[:, :, None]
    w = torch.softmax(s, dim=1)
    v = _v(values, B, T, heads, hs)
    _v(pooled, B, heads * hs).copy_((w[..., None] * v).sum(1).reshape(B, heads * hs))


def ctts_glu(h, rows, C, g, stream):
    hv = _v(h, rows, 2 * C)
    _v(g, rows, C).copy_(hv[:, :C] * torch.sigmoid(hv[:, C:]))


def ctts_relshift_softmax(content, pos, Z, T, ldp, sqrt_dim, P, stream):
    c, p = _v(content, Z, T, T), _v(pos, Z, T, T)
    padded = torch.cat([p.new_zeros(Z, T, 1), p], dim=-1).view(Z, T + 1, T)[:, 1:].reshape(Z, T, T)
    out = _v(P, Z, T, ldp)
    out.zero_()
    out[:, :, :T] = torch.softmax((c + padded) / sqrt_dim, -1)


def ctts_relshift_softmax_planes(content, pos, Z, T, ld, ldp, sqrt_dim, n, planes, stream):
    c, p = _v(content, Z, T, ld)[:, :, :T], _v(pos, Z, T, ld)[:, :, :T]
    padded = torch.cat([p.new_zeros(Z, T, 1), p], dim=-1).reshape(Z, T + 1, T)[:, 1:].reshape(Z, T, T)
    out = torch.zeros(Z, T, ldp)
    out[:, :, :T] = torch.softmax((c + padded) / sqrt_dim, -1)
    _split_into(out, _planes(planes, n), Z, T, ldp)


def ctts_pad_heads_planes(x, bias, rows, ld_in, c0, H, DH, DHp, n, planes, stream):
    xv = _v(x, rows, ld_in)[:, c0:c0 + H * DH]
    if bias is not None:
        xv = xv + _v(bias, H * DH)
    out = torch.zeros(rows, H, DHp)
    out[:, :, :DH] = xv.reshape(rows, H, DH)
    _split_into(out, _planes(planes, n), rows, H, DHp)


def ctts_gru_bidir(gi_f, gi_b, whh_f, bhh_f, whh_b, bhh_b, B, T, H, out, h_final, stream):
    o = _v(out, B, T, 2 * H)
    hf = _v(h_final, B, 2 * H)
    for d, (gi, whh, bhh) in enumerate(((gi_f, whh_f, bhh_f), (gi_b, whh_b, bhh_b))):
        g, w, b = _v(gi, B, T, 3 * H), _v(whh, 3 * H, H), _v(bhh, 3 * H)
        h = torch.zeros(B, H)
        for t in (range(T - 1, -1, -1) if d else range(T)):
            gh = h @ w.t() + b
            r = torch.sigmoid(g[:, t, :H] + gh[:, :H])
            z = torch.sigmoid(g[:, t, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(g[:, t, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
            o[:, t, d * H:(d + 1) * H] = h
        hf[:, d * H:(d + 1) * H] = h


def ctts_linear_smallk(x, w, bias, residual, rows, K, N, y, stream):
    out = _v(x, rows, K) @ _v(w, N, K).t()
    if bias is not None:
        out = out + _v(bias, N)
    if residual is not None:
        out = out + _v(residual, rows, N)
    _v(y, rows, N).copy_(out)


def ctts_dwconv_bn_swish(g, w, K, scale, shift, B, T, C, y, stream):
    conv = F.conv1d(_v(g, B, T, C).transpose(1, 2), _v(w, C, 1, K), None, padding=K // 2, groups=C).transpose(1, 2)
    v = conv * _v(scale, C) + _v(shift, C)
    _v(y, B, T, C).copy_(v * torch.sigmoid(v))


# ---- training step ------------------------------------------------------------------------------------------------
def ctts_gemm_generic(a, b, y, Z, zmod, M, N, K, a_str, b_str, y_str, Kin, shift0, shift_z, alpha, accumulate, stream):
    a_str, b_str, y_str = list(a_str), list(b_str), list(y_str)
    if Kin <= 0:
        Kin = K
    assert K % Kin == 0, "emulator: K must be a whole number of Kin blocks"
    KB = K // Kin
    af, bf, yf = _flat(a), _flat(b), _flat(y)
    for z in range(Z):
        zo, zi = divmod(z, zmod)
        A = af.as_strided((M, KB, Kin), (a_str[2], a_str[4], a_str[3]), af.storage_offset() + zo * a_str[0] + zi * a_str[1])
        shift = shift0 + z * shift_z
        Bsrc = bf.as_strided((N, KB, Kin), (b_str[2], b_str[4], b_str[3]), bf.storage_offset() + zo * b_str[0] + zi * b_str[1])
        Bm = torch.zeros(N, KB, Kin)
        lo, hi = max(0, -shift), min(Kin, Kin - shift)
        if hi > lo:
            Bm[:, :, lo:hi] = Bsrc[:, :, lo + shift:hi + shift]
        out = alpha * (A.reshape(M, K) @ Bm.reshape(N, K).t())
        Y = yf.as_strided((M, N), (y_str[2], y_str[3]), yf.storage_offset() + zo * y_str[0] + zi * y_str[1])
        Y.copy_(Y + out if accumulate else out)


def ctts_act_bwd(dy, ref, act, alpha, lens, Z, T, rows, N, dz, dbias, stream):
    g = _v(dy, Z * rows, N).clone()
    if lens is not None:
        Bn = Z * rows // T
        g = g * _keep(lens, Bn, T).reshape(-1, 1)
    g = g * alpha
    if act != ACT_NONE:
        g = g * _act_grad(_v(ref, Z * rows, N), act)
    if dz is not None:
        _v(dz, Z * rows, N).copy_(g)
    if dbias is not None:
        db = _v(dbias, Z, N)
        db.copy_(db + g.view(Z, rows, N).sum(1))


def ctts_act_bwd_planes(dy, ref, act, alpha, lens, B, T, N, Tp, dz, n, dzp, dztp, dbias, stream):
    g = _v(dy, B, T, N).clone() * _keep(lens, B, T)[:, :, None] * alpha
    if act != ACT_NONE:
        g = g * _act_grad(_v(ref, B, T, N), act)
    if dz is not None:
        _v(dz, B, T, N).copy_(g)
    _split_into(g, _planes(dzp, n), B, T, N)
    if dztp is not None:
        gt = torch.zeros(B, N, Tp)
        gt[:, :, :T] = g.transpose(1, 2)
        _split_into(gt, _planes(dztp, n), B, N, Tp)
    if dbias is not None:
        _v(dbias, N).add_(g.reshape(-1, N).sum(0))


def ctts_dropout_add(x, res, lens, B, T, C, p, seed, offset, offset_dev, y, stream):
    n = B * T * C
    if offset_dev is not None:
        offset = int(offset) + int(_v(offset_dev, 1)[0])
    keep = dropout_mask(n, p, seed, offset)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    out = (_v(res, n) + _v(x, n) * keep * float(inv)).view(B, T, C) * _keep(lens, B, T)[:, :, None]
    _v(y, B, T, C).copy_(out)


def ctts_layernorm_bwd(x, gamma, dy, eps, lens, B, T, C, dx, accumulate, dgamma, dbeta, stream):
    xv = _v(x, B, T, C).clone().requires_grad_(True)
    gm = _v(gamma, C).clone().requires_grad_(True)
    bt = torch.zeros(C, requires_grad=True)
    y = F.layer_norm(xv, (C,), gm, bt, eps) * _keep(lens, B, T)[:, :, None]
    y.backward(_v(dy, B, T, C))
    d = _v(dx, B, T, C)
    d.copy_(d + xv.grad if accumulate else xv.grad)
    if dgamma is not None:
        g = _v(dgamma, C)
        g.copy_(g + gm.grad)
    if dbeta is not None:
        g = _v(dbeta, C)
        g.copy_(g + bt.grad)


def ctts_mask_rows(x, lens, B, T, C, stream):
    xv = _v(x, B, T, C)
    xv.copy_(xv * _keep(lens, B, T)[:, :, None])


def ctts_axpy(x, a, n, accumulate, y, stream):
    yv = _v(y, n)
    yv.copy_((yv if accumulate else 0) + a * _v(x, n))


def ctts_rowscale_axpy(x, s, a, rows, C, accumulate, y, stream):
    yv = _v(y, rows, C)
    yv.copy_((yv if accumulate else 0) + a * _v(x, rows, C) * _v(s, rows)[:, None])


def ctts_scatter_add_rows(dy, idx, lens, T, rows, C, table_rows, skip_idx, scale, dtable, stream):
    g = _v(dy, rows, C) * scale
    ids = _v(idx, rows).clamp(0, table_rows - 1)
    m = ids != skip_idx
    if lens is not None:
        m = m & _keep(lens, rows // T, T).reshape(-1)
    _v(dtable, table_rows, C).index_add_(0, ids[m], g[m])


def ctts_length_expand_bwd(dy, cum_lr, B, S, C, M, accumulate, dsrc, stream):
    cum = _v(cum_lr, B, S).long()
    g = _v(dy, B, M, C)
    out = torch.zeros(B, S, C)
    j = torch.searchsorted(cum, torch.arange(M)[None].expand(B, M).contiguous(), right=True)
    for b in range(B):
        ok = j[b] < S
        out[b].index_add_(0, j[b][ok], g[b][ok])
    d = _v(dsrc, B, S, C)
    d.copy_(d + out if accumulate else out)


def ctts_add_positions_bwd(dy, x, pe, pe_rows, lens, B, T, C, pos_mode, dalpha, stream):
    xv = _v(x, B, T, C)
    pos = _positions(xv[..., 0] != 0) if pos_mode == 0 else torch.arange(T)[None].expand(B, T)
    g = _v(dy, B, T, C) * _keep(lens, B, T)[:, :, None]
    _v(dalpha, 1).add_((g * _v(pe, pe_rows, C)[pos]).sum())


def ctts_masked_softmax(S, lens, H, Z, T, Tk, ld, mask_rows, P, stream):
    s = _v(S, Z, T, ld)[:, :, :Tk]
    out = _v(P, Z, T, ld)
    out.zero_()
    if lens is None:
        out[:, :, :Tk] = torch.softmax(s, -1)
        return
    ln = _v(lens, (Z - 1) // H + 1)
    for z in range(Z):
        L = min(int(ln[z // H]), Tk)
        p = torch.softmax(s[z, :, :L], -1)
        if mask_rows:
            p = p * (torch.arange(T) < int(ln[z // H]))[:, None]
        out[z, :, :L] = p


def ctts_softmax_bwd(P, dP, Z, T, Tk, ld, scale, dS, stream):
    p = _v(P, Z, T, ld)[:, :, :Tk]
    g = _v(dP, Z, T, ld)[:, :, :Tk]
    res = p * (g - (p * g).sum(-1, keepdim=True)) * scale
    out = _v(dS, Z, T, ld)
    out.zero_()
    out[:, :, :Tk] = res


def ctts_bn_stats(x, rows, C, mean, var, stream):
    xv = _v(x, rows, C)
    _v(mean, C).copy_(xv.mean(0))
    _v(var, C).copy_(xv.var(0, unbiased=False))


def ctts_bn_act_fwd(x, mean, var, gamma, beta, eps, act, rows, C, y, n, planes, stream):
    v = (_v(x, rows, C) - _v(mean, C)) * torch.rsqrt(_v(var, C) + eps) * _v(gamma, C) + _v(beta, C)
    v = _act(v, act)
    if y is not None:
        _v(y, rows, C).copy_(v)
    if n:
        _split_into(v, _planes(planes, n), rows, C)


def ctts_bn_update_running(mean, var, rows, momentum, C, rmean, rvar, count, stream):
    unb = rows / (rows - 1) if rows > 1 else 1.0
    rm, rv = _v(rmean, C), _v(rvar, C)
    rm.copy_((1 - momentum) * rm + momentum * _v(mean, C))
    rv.copy_((1 - momentum) * rv + momentum * _v(var, C) * unb)
    if count is not None:
        _v(count, 1).add_(1)


def ctts_bn_bwd(dy, x, mean, var, gamma, beta, eps, act, rows, C, dx, dgamma, dbeta, ws, stream):
    xv = _v(x, rows, C).clone().requires_grad_(True)
    gm = _v(gamma, C).clone().requires_grad_(True)
    bt = _v(beta, C).clone().requires_grad_(True)
    y = _act(F.batch_norm(xv, None, None, gm, bt, True, 0.1, eps), act)
    y.backward(_v(dy, rows, C))
    _v(dx, rows, C).copy_(xv.grad)
    if dgamma is not None:
        _v(dgamma, C).add_(gm.grad)
    if dbeta is not None:
        _v(dbeta, C).add_(bt.grad)


def _philox(seed, ctr, offset):
    """Philox4x32-10, vectorised over `ctr` (uint64 array); returns [len, 4] uint32."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c = [ctr & np.uint64(0xFFFFFFFF), ctr >> np.uint64(32), np.full_like(ctr, offset & 0xFFFFFFFF),
         np.full_like(ctr, offset >> 32)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack(c, 1).astype(np.uint32)


def dropout_mask(n, p, seed, offset):
    q = np.arange((n + 3) // 4, dtype=np.uint64)
    r = _philox(int(seed), q, int(offset)).reshape(-1)[:n]
    thresh = np.uint32(min(np.float32(p) * np.float32(4294967296.0), np.float32(4294967295.0)))
    return torch.from_numpy((r >= thresh).astype(np.float32))


def ctts_dropout(x, n, p, seed, offset, offset_dev, y, stream):
    if offset_dev is not None:
        offset = int(offset) + int(_v(offset_dev, 1)[0])
    keep = dropout_mask(n, p, seed, offset)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    _v(y, n).copy_(_v(x, n) * keep * float(inv))


def ctts_pack_conv_weight_dgrad(w, N, Cin, taps, wd, stream):
    wv = _v(w, N, Cin, taps)
    _v(wd, Cin, taps, N).copy_(wv.flip(2).permute(1, 2, 0))


def ctts_unpack_conv_wgrad(dwp, N, Cin, taps, accumulate, dw, stream):
    v = _v(dwp, N, taps, Cin).permute(0, 2, 1)
    d = _v(dw, N, Cin, taps)
    d.copy_(d + v if accumulate else v)


def ctts_split_transpose(x, Z, R, C, ld_in, c0, Rp, taps, n, planes, stream):
    xv = _v(x, Z, R, ld_in)[:, :, c0:c0 + C].transpose(1, 2)          # [Z, C, R]
    out = torch.zeros(Z, taps, C, Rp)
    pad = taps // 2
    for j in range(taps):
        sh = j - pad
        lo, hi = max(0, -sh), min(R, R - sh)
        if hi > lo:
            out[:, j, :, lo:hi] = xv[:, :, lo + sh:hi + sh]
    _split_into(out, _planes(planes, n), Z, taps, C, Rp)


def ctts_gemm_wgrad(n, dzT, xT, B, T, Tp, Cin, N, taps, alpha, accumulate, dwp, stream):
    dz = _val(dzT, n, B, N, Tp)[:, :, :T]
    xs = _val(xT, n, B, taps, Cin, Tp)[:, :, :, :T]
    out = torch.einsum("bnt,bjct->njc", dz, xs)
    d = _v(dwp, N, taps, Cin)
    d.copy_((d if accumulate else 0) + alpha * out)


def ctts_gemm_wgrad_rowmajor(n, dzp, xp, B, T, Cin, N, taps, alpha, accumulate, dwp, stream):
    dz, x = _val(dzp, n, B, T, N), _val(xp, n, B, T, Cin)
    pad = taps // 2
    xs = torch.stack([F.pad(x, (0, 0, pad, pad))[:, j:j + T] for j in range(taps)], 1)       # [B, taps, T, Cin]
    out = torch.einsum("btn,bjtc->njc", dz, xs)
    d = _v(dwp, N, taps, Cin)
    d.copy_((d if accumulate else 0) + alpha * out)


def ctts_aligner_attention_bwd(soft, logprob, prior, dsoft, dlogprob, src_lens, B, M, S, da, stream):
    so, lp = _v(soft, B, M, S), _v(logprob, B, M, S)
    keep = _keep(src_lens, B, S)[:, None, :]
    dlp = _v(dlogprob, B, M, S).clone() if dlogprob is not None else torch.zeros(B, M, S)
    if dsoft is not None:
        g = _v(dsoft, B, M, S)
        dlp = dlp + keep * so * (g - (so * g * keep).sum(-1, keepdim=True))
    L = lp - torch.log(_v(prior, B, S, M).transpose(1, 2) + 1e-8)
    _v(da, B, M, S).copy_(dlp - torch.exp(L) * dlp.sum(-1, keepdim=True))


def ctts_glu_bwd(h, dg, rows, C, dh, stream):
    hv, g = _v(h, rows, 2 * C), _v(dg, rows, C)
    s = torch.sigmoid(hv[:, C:])
    out = _v(dh, rows, 2 * C)
    out[:, :C] = g * s
    out[:, C:] = g * hv[:, :C] * s * (1 - s)


def ctts_dwconv(x, w, K, B, T, C, y, stream):
    _v(y, B, T, C).copy_(F.conv1d(_v(x, B, T, C).transpose(1, 2), _v(w, C, 1, K), None, padding=K // 2,
                                  groups=C).transpose(1, 2))


def ctts_dwconv_bwd(dy, x, w, K, B, T, C, dx, dw, stream):
    xv = _v(x, B, T, C).clone().requires_grad_(True)
    wv = _v(w, C, 1, K).clone().requires_grad_(True)
    y = F.conv1d(xv.transpose(1, 2), wv, None, padding=K // 2, groups=C).transpose(1, 2)
    y.backward(_v(dy, B, T, C))
    if dx is not None:
        _v(dx, B, T, C).copy_(xv.grad)
    if dw is not None:
        _v(dw, C, 1, K).add_(wv.grad)


def ctts_relshift_bwd(dscore, Z, T, ld, ld_out, sqrt_dim, dcontent, dpos, stream):
    ds = _v(dscore, Z, T, ld)[:, :, :T]
    p = torch.zeros(Z, T, T, requires_grad=True)
    padded = torch.cat([p.new_zeros(Z, T, 1), p], dim=-1).view(Z, T + 1, T)[:, 1:].reshape(Z, T, T)
    (padded / sqrt_dim).backward(ds)
    dc, dp = _v(dcontent, Z, T, ld_out), _v(dpos, Z, T, ld_out)
    dc.zero_()
    dp.zero_()
    dc[:, :, :T] = ds / sqrt_dim
    dp[:, :, :T] = p.grad


def ctts_fastformer_pool_bwd(logits, values, lens, dpooled, B, T, heads, hs, dlogits, dvalues, stream):
    lg = _v(logits, B, T, heads).clone().requires_grad_(True)
    v = _v(values, B, T, heads, hs).clone().requires_grad_(True)
    keep = _keep(lens, B, T)
    s = lg / math.sqrt(hs) + torch.where(keep, torch.full((), -10000.0), torch.zeros(()))[:, :, None]
    pooled = (torch.softmax(s, dim=1)[..., None] * v).sum(1).reshape(B, heads * hs)
    pooled.backward(_v(dpooled, B, heads * hs))
    _v(dlogits, B, T, heads).copy_(lg.grad)
    _v(dvalues, B, T, heads, hs).copy_(v.grad)


def ctts_mul_bwd(dy, a, b, b_rowwise, lens, B, T, C, da, db, stream):
    g = _v(dy, B, T, C) * _keep(lens, B, T)[:, :, None]
    av = _v(a, B, T, C)
    bv = _v(b, B, C)[:, None, :] if b_rowwise else _v(b, B, T, C)
    if da is not None:
        _v(da, B, T, C).copy_(g * bv)
    if db is not None:
        if b_rowwise:
            _v(db, B, C).add_((g * av).sum(1))
        else:
            _v(db, B, T, C).copy_(g * av)


def ctts_gru_bwd(gi, whh, bhh, out, out_ld, out_off, dout, dh_final, dhf_ld, B, T, H, reverse, dgi, dgh, stream):
    g, w, b = _v(gi, B, T, 3 * H), _v(whh, 3 * H, H), _v(bhh, 3 * H)
    of = _flat(out)
    o = of.as_strided((B, T, H), (T * out_ld, out_ld, 1), of.storage_offset() + out_off)
    do = None
    if dout is not None:
        df = _flat(dout)
        do = df.as_strided((B, T, H), (T * out_ld, out_ld, 1), df.storage_offset() + out_off)
    dh = torch.zeros(B, H)
    if dh_final is not None:
        hf = _flat(dh_final)
        dh = hf.as_strided((B, H), (dhf_ld, 1), hf.storage_offset()).clone()
    dgi_v, dgh_v = _v(dgi, B, T, 3 * H), _v(dgh, B, T, 3 * H)
    for step in range(T - 1, -1, -1):
        t = (T - 1 - step) if reverse else step
        tp = t + 1 if reverse else t - 1
        hp = o[:, tp] if step > 0 else torch.zeros(B, H)
        gh = hp @ w.t() + b
        r = torch.sigmoid(g[:, t, :H] + gh[:, :H])
        z = torch.sigmoid(g[:, t, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(g[:, t, 2 * H:] + r * gh[:, 2 * H:])
        dht = dh + (do[:, t] if do is not None else 0)
        d_n = dht * (1 - z) * (1 - n * n)
        d_z = dht * (hp - n) * z * (1 - z)
        d_r = d_n * gh[:, 2 * H:] * r * (1 - r)
        dgi_v[:, t] = torch.cat([d_r, d_z, d_n], 1)
        dg = torch.cat([d_r, d_z, d_n * r], 1)
        dgh_v[:, t] = dg
        dh = dht * z + dg @ w


def ctts_gemm_batched_planes(n, ap, a_view, wp, w_view, addr, y_outer, y_inner, alpha, residual, lens, Z, T, K, N, y, yp,
                             stream):
    a_view, w_view, addr = list(a_view), list(w_view), list(addr)
    mod, a_div, a_c0, a_step, w_div, w_c0, w_step, lens_div, ldy = addr
    A = sum(_flat(p).float() for p in _planes(ap, n))
    W = sum(_flat(p).float() for p in _planes(wp, n))
    for z in range(Z):
        zh = z % mod
        za, zw = z // a_div, z // w_div
        # out-of-range rows / columns of the views are TMA zero fill on the device
        ka = max(0, min(K, a_view[0] - (a_c0 + zh * a_step)))
        kw = max(0, min(K, w_view[0] - (w_c0 + zh * w_step)))
        kk = min(ka, kw)
        ra, rw = min(T, a_view[1]), min(N, w_view[1])
        Az = A.as_strided((ra, kk), (a_view[3], 1), za * a_view[4] + a_c0 + zh * a_step)
        Wz = W.as_strided((rw, kk), (w_view[3], 1), zw * w_view[4] + w_c0 + zh * w_step)
        out = torch.zeros(T, N)
        out[:ra, :rw] = alpha * (Az @ Wz.t())
        off = (z // mod) * y_outer + zh * y_inner
        if residual is not None:
            rf = _flat(residual)
            out = out + rf.as_strided((T, N), (ldy, 1), rf.storage_offset() + off)
        if lens is not None:
            out = out * (torch.arange(T) < int(_flat(lens)[z // lens_div]))[:, None]
        if y is not None:
            yf = _flat(y)
            yf.as_strided((T, N), (ldy, 1), yf.storage_offset() + off).copy_(out)
        if yp is not None:
            rem = out.clone()
            for p in _planes(yp, n):
                pf = _flat(p)
                h = rem.to(torch.bfloat16)
                pf.as_strided((T, N), (ldy, 1), pf.storage_offset() + off).copy_(h)
                rem = rem - h.float()


def ctts_act_fwd(x, n, act, y, n_planes, planes, stream):
    v = _act(_v(x, n), act)
    if y is not None:
        _v(y, n).copy_(v)
    if n_planes:
        _split_into(v, _planes(planes, n_planes), n)


def ctts_add_coords(x, N, H, W, y, stream):
    xv = _v(x, N, H, W)
    xx = (torch.arange(H).float() / (H - 1) * 2 - 1)[None, :, None].expand(N, H, W)
    yy = (torch.arange(W).float() / (W - 1) * 2 - 1)[None, None, :].expand(N, H, W)
    rr = torch.sqrt((xx - 0.5) ** 2 + (yy - 0.5) ** 2)
    _v(y, N, H, W, 4).copy_(torch.stack([xv, xx, yy, rr], -1))


def _unfold(xv, N, H, W, C):
    """[N,H,W,C] -> [N*H*Wo, 9*C] with column order (kh*3 + kw)*C + c (3x3, stride (1,2), pad (1,1))."""
    cols = F.unfold(xv.permute(0, 3, 1, 2), (3, 3), padding=(1, 1), stride=(1, 2))       # [N, C*9, H*Wo], rows c*9 + k
    Wo = (W + 2 - 3) // 2 + 1
    return cols.view(N, C, 9, H * Wo).permute(0, 3, 2, 1).reshape(N * H * Wo, 9 * C), Wo


def ctts_im2col_3x3_s12(x, N, H, W, C, col, stream):
    c, Wo = _unfold(_v(x, N, H, W, C), N, H, W, C)
    _v(col, N * H * Wo, 9 * C).copy_(c)


def ctts_col2im_3x3_s12(dcol, N, H, W, C, dx, stream):
    xv = torch.zeros(N, H, W, C, requires_grad=True)
    c, Wo = _unfold(xv, N, H, W, C)
    c.backward(_v(dcol, N * H * Wo, 9 * C))
    _v(dx, N, H, W, C).copy_(xv.grad)


def ctts_permute_last2(x, rows, A, Bd, y, stream):
    _v(y, rows, Bd, A).copy_(_v(x, rows, A, Bd).transpose(1, 2))


def ctts_merge_planes(n_planes, planes, n, y, stream):
    _v(y, n).copy_(_val(planes, n_planes, n))


def ctts_copy_rows(src, src_stride, rows, C, dst, dst_stride, accumulate, stream):
    sf, df = _flat(src), _flat(dst)
    s = sf.as_strided((rows, C), (src_stride, 1), sf.storage_offset())
    d = df.as_strided((rows, C), (dst_stride, 1), df.storage_offset())
    d.copy_(d + s if accumulate else s)


ENTRIES = {k: v for k, v in globals().items() if k.startswith("ctts_")}
CALLS = []      # names of the entry points called since the last reset (tests assert on it)


def call(name, *args):
    if name not in ENTRIES:
        raise NotImplementedError("capi emulator: %s is not restated" % name)
    CALLS.append(name)
    args = [a.detach() if torch.is_tensor(a) else a for a in args]     # same storage, no autograd history
    with torch.enable_grad():      # some restatements differentiate a torch expression internally
        ENTRIES[name](*args)


def install(monkeypatch):
    """Route the product's capi.call to this emulator and let CPU tensors through (tests only)."""
    from ctts_b200 import capi, engine
    monkeypatch.setattr(capi, "call", call)
    monkeypatch.setattr(capi, "require_device", lambda: None)
    monkeypatch.setattr(capi, "require_cuda_tensor", lambda t: None)
    monkeypatch.setattr(engine, "_stream", lambda: 0)
    del CALLS[:]
