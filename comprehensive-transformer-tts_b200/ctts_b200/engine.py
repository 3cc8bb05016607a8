"""Host-side orchestration of the CompTransTTS forward path over libctts_b200's C ABI.

PyTorch is used for device memory (torch.empty on the caching allocator), the current stream and
trivial shape bookkeeping only; every arithmetic step of the path is a kernel of libctts_b200
reached through `capi.call` with raw device pointers.  Function names mirror the reference's
modules; file:line citations point at the reference code each function stands in for.
"""
import math
import os

import torch

from . import capi
from .capi import ACT_GELU, ACT_NONE, ACT_RELU, ACT_TANH

_ACTS = {"gelu": ACT_GELU, "relu": ACT_RELU, "none": ACT_NONE, "tanh": ACT_TANH}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _i64(t):
    return t if (t.dtype == torch.int64 and t.is_contiguous()) else t.long().contiguous()


# ---------------------------------------------------------------------------------------------
# thin op wrappers (shape checking lives in the C layer; these only allocate outputs)
# ---------------------------------------------------------------------------------------------
def conv_gemm(x, w, bias=None, alpha=1.0, bn=None, act=ACT_NONE, residual=None, lens=None, taps=1, out=None):
    """y = epilogue(conv1d_k(x) or linear(x)); x [B,T,Cin] fp32, w packed [N, taps*Cin]."""
    B, T, Cin = x.shape
    N = w.shape[0]
    assert w.shape[1] == taps * Cin, (tuple(w.shape), taps, Cin)
    y = out if out is not None else torch.empty(B, T, N, device=x.device, dtype=torch.float32)
    capi.call("ctts_conv1d_gemm", x, w, bias, float(alpha), bn[0] if bn else None, bn[1] if bn else None, int(act),
              residual, lens, B, T, Cin, N, taps, y, _stream())
    return y


class Planes:
    """A [B, T, C] activation (or [N, K] weight) stored as 2 or 3 bf16 planes: value ~= sum(p).
    2 planes = 16 mantissa bits ("bf16x3": 3 MMAs), 3 planes = 24 bits ("bf16x6": 6 MMAs, FP32-equivalent)."""
    __slots__ = ("p",)

    def __init__(self, planes):
        self.p = list(planes)

    @property
    def shape(self):
        return self.p[0].shape

    @property
    def n(self):
        return len(self.p)

    @property
    def hi(self):
        return self.p[0]

    @property
    def lo(self):
        assert len(self.p) == 2
        return self.p[1]

    def value(self):
        return sum(t.float() for t in self.p)

    @staticmethod
    def empty(shape, device, n=2):
        return Planes([torch.empty(shape, device=device, dtype=torch.bfloat16) for _ in range(n)])


def split_planes(x, n=2):
    """fp32 tensor -> Planes (ctts_split_planes)."""
    x = _f32(x)
    p = Planes.empty(x.shape, x.device, n)
    capi.call("ctts_split_planes", x, x.numel(), n, capi.ptr_array(p.p), _stream())
    return p


def gemm_tc(xp, wp, bias=None, alpha=1.0, bn=None, act=ACT_NONE, residual=None, lens=None, taps=1, out=None,
            want_fp32=True, want_planes=False):
    """tcgen05 implicit-GEMM conv / linear on bf16 operand planes (2 planes: bf16x3, 3 planes: bf16x6).
    xp: Planes [B,T,Cin]; wp: Planes [N, taps*Cin].  Returns (y fp32 or None, Planes or None)."""
    B, T, Cin = xp.shape
    N = wp.shape[0]
    assert wp.shape[1] == taps * Cin, (tuple(wp.shape), taps, Cin)
    assert xp.n == wp.n, "activation and weight must be split into the same number of planes"
    dev = xp.p[0].device
    y = out if out is not None else (torch.empty(B, T, N, device=dev, dtype=torch.float32) if want_fp32 else None)
    yp = Planes.empty((B, T, N), dev, xp.n) if want_planes else None
    capi.call("ctts_gemm_split", xp.n, capi.ptr_array(xp.p), capi.ptr_array(wp.p), bias, float(alpha),
              bn[0] if bn else None, bn[1] if bn else None, int(act), residual, lens, B, T, Cin, N, taps, y,
              capi.ptr_array(yp.p) if yp else None, _stream())
    return y, yp


FUSED_LN = os.environ.get("CTTS_FUSED_LN", "1") != "0"


def gemm_tc_ln(xp, wp, bias, residual, lens, out, gamma, beta, eps, ln_lens=None, want_fp32=False, taps=1):
    """Projection + residual + LayerNorm in one launch (ctts_gemm_split_ln; 2 planes, N = 256): `out` receives
    conv(x) + bias + residual (rows beyond lens zeroed), returns (LayerNorm(out) fp32 or None, LayerNorm(out) planes)."""
    B, T, Cin = xp.shape
    N = wp.shape[0]
    dev = xp.p[0].device
    ln_y = torch.empty(B, T, N, device=dev, dtype=torch.float32) if want_fp32 else None
    lp = Planes.empty((B, T, N), dev, 2)
    assert ln_lens is None or ln_lens is lens
    capi.call("ctts_gemm_split_ln", capi.ptr_array(xp.p), capi.ptr_array(wp.p), bias, 1.0, residual, lens, B, T, Cin, N, taps,
              out, gamma, beta, float(eps), 1 if ln_lens is not None else 0, ln_y, capi.ptr_array(lp.p), _stream())
    return ln_y, lp


def layernorm_planes(x, gamma, beta, eps, lens=None, want_fp32=False, n=2):
    """LayerNorm whose result is written as bf16 planes (and optionally fp32)."""
    B, T, C = x.shape
    y = torch.empty_like(x) if want_fp32 else None
    yp = Planes.empty((B, T, C), x.device, n)
    capi.call("ctts_layernorm_planes", x, gamma, beta, float(eps), lens, B, T, C, y, n, capi.ptr_array(yp.p), _stream())
    return y, yp


def layernorm(x, gamma, beta, eps, lens=None):
    B, T, C = x.shape
    y = torch.empty_like(x)
    capi.call("ctts_layernorm", x, gamma, beta, float(eps), lens, B, T, C, y, _stream())
    return y


def attention(qkv, lens, n_head):
    B, T, C3 = qkv.shape
    C = C3 // 3
    out = torch.empty(B, T, C, device=qkv.device, dtype=torch.float32)
    capi.call("ctts_attention", qkv, lens, B, T, C, n_head, 1.0 / math.sqrt(C // n_head), out, _stream())
    return out


FLASH = os.environ.get("CTTS_FLASH_ATTENTION", "1") != "0"
PARALLEL_BRANCHES = os.environ.get("CTTS_PARALLEL_BRANCHES", "1") != "0"
SMALL_ATTENTION = os.environ.get("CTTS_SMALL_ATTENTION", "1") != "0"
_SIDE_STREAMS = {}


def _side_stream(device):
    """One extra stream per device for the independent predictor branches of a captured forward."""
    k = str(device)
    if k not in _SIDE_STREAMS:
        _SIDE_STREAMS[k] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[k]


def attention_flash(qkv_planes, lens, n_head):
    """Fused tensor-core attention (ctts_flash_attention_bf16x3): head_dim 128, 2 planes; scores never reach HBM."""
    B, T, C3 = qkv_planes.shape
    C = C3 // 3
    dev = qkv_planes.p[0].device
    out = Planes.empty((B, T, C), dev, 2)
    # (the two V^T arguments are vestigial: the kernel reads V from the qkv planes as an MN-major operand)
    capi.call("ctts_flash_attention_bf16x3", qkv_planes.p[0], qkv_planes.p[1], None, None, lens, B, T, C, n_head,
              1.0 / math.sqrt(C // n_head), out.p[0], out.p[1], _stream())
    return out


def attention_tc(qkv_planes, lens, n_head):
    """Tensor-core masked self-attention on bf16 planes (ctts_attention_split: 2 planes = bf16x3, 3 planes = bf16x6);
    returns the output planes."""
    B, T, C3 = qkv_planes.shape
    C = C3 // 3
    n = qkv_planes.n
    if FLASH and n == 2 and C // n_head == 128:
        return attention_flash(qkv_planes, lens, n_head)
    dev = qkv_planes.p[0].device
    Tp = (T + 7) // 8 * 8
    Z = B * n_head
    if SMALL_ATTENTION and n == 3 and C // n_head == 128 and T <= 128:
        # short sequences (the encoder at phoneme lengths): one fused kernel, one CTA per (batch, head)
        out = Planes.empty((B, T, C), dev, 3)
        capi.call("ctts_attention_small", capi.ptr_array(qkv_planes.p), lens, B, T, C, n_head, 1.0 / math.sqrt(C // n_head),
                  capi.ptr_array(out.p), _stream())
        return out
    scores = torch.empty(Z * T * Tp, device=dev, dtype=torch.float32)
    pp = [torch.empty(Z * T * Tp, device=dev, dtype=torch.bfloat16) for _ in range(n)]
    vt = [torch.empty(B * C * Tp, device=dev, dtype=torch.bfloat16) for _ in range(n)]
    out = Planes.empty((B, T, C), dev, n)
    capi.call("ctts_attention_split", n, capi.ptr_array(qkv_planes.p), lens, B, T, C, n_head, 1.0 / math.sqrt(C // n_head),
              scores, capi.ptr_array(pp), capi.ptr_array(vt), capi.ptr_array(out.p), None, _stream())
    return out


def gemm_batched_planes(a, a_view, w, w_view, addr, y_outer, y_inner, alpha, Z, T, K, N, y=None, y_planes=None,
                        residual=None, lens=None):
    """ctts_gemm_batched_planes: y[z][t, n] = alpha * sum_k A[z][t, k] W[z][n, k] on operand views of bf16 planes
    (a_view / w_view = (d0, d1, d2, s1, s2) in elements; addr = (mod, a_div, a_c0, a_step, w_div, w_c0, w_step, lens_div, ldy))."""
    import ctypes
    ll = lambda v: (ctypes.c_longlong * len(v))(*[int(x) for x in v])
    ii = (ctypes.c_int * len(addr))(*[int(x) for x in addr])
    capi.call("ctts_gemm_batched_planes", a.n, capi.ptr_array(a.p), ll(a_view), capi.ptr_array(w.p), ll(w_view), ii,
              int(y_outer), int(y_inner), float(alpha), residual, lens, Z, T, K, N, y,
              capi.ptr_array(y_planes.p) if y_planes is not None else None, _stream())


def pad_mask(lens, max_len):
    """utils/tools.py:188-196 (bool bookkeeping tensor handed back to the caller; True = padding)."""
    return torch.arange(int(max_len), device=lens.device)[None, :] >= lens[:, None]


# ---------------------------------------------------------------------------------------------
# prepared (device-resident, kernel-layout) weights
# ---------------------------------------------------------------------------------------------
class Prepared:
    """Kernel-layout copies of the parameters: packed conv weights, folded BatchNorm, sinusoid tables.

    Rebuilt whenever a parameter's storage or version changes (load_state_dict, .to(device), an
    optimizer step), so the nn.Module stays the single source of truth.
    """

    def __init__(self, module):
        self.module = module
        self.sig = None
        self.w = {}
        self.pe = {}
        self.pe_retired = []
        self._names = self._tensors = None
        self._epoch = -1
        self._host_maxes = None

    def host_maxes(self, device):
        """Pinned landing buffer of the step's single device -> host read (allocated outside any graph capture)."""
        if self._host_maxes is None:
            self._host_maxes = torch.zeros(2, dtype=torch.int64).pin_memory() if device.type == "cuda" \
                else torch.zeros(2, dtype=torch.int64)
        return self._host_maxes

    def _signature(self, tensors):
        return [(v.data_ptr(), v._version) for v in tensors]

    def params(self):
        from .module import _Tracked
        if self._names is None or self._epoch != _Tracked.structure_epoch:
            # walk the module tree only when its structure changed (a tensor or sub-module attached / replaced)
            P = {k: v for k, v in self.module.named_parameters(remove_duplicate=False)}  # tied names included
            P.update({k: v for k, v in self.module.named_buffers(remove_duplicate=False)})
            self._names, self._tensors = P, list(P.values())
            self._epoch = _Tracked.structure_epoch
            self.sig = None
        P = self._names
        sig = self._signature(self._tensors)
        if self.sig is None or sig != self.sig:
            self._build(P)
            self.sig = sig
        return P

    def _build(self, P):
        self.w = {}
        st = _stream()
        with torch.no_grad():
            for name, t in P.items():
                if t.dim() == 3 and t.dtype == torch.float32 and not name.endswith("position_enc"):
                    n, cin, taps = t.shape
                    packed = torch.empty(n, taps * cin, device=t.device, dtype=torch.float32)
                    capi.call("ctts_pack_conv_weight", _f32(t), n, cin, taps, packed, st)
                    self.w[name] = packed
            for i in range(5):  # eval-mode BatchNorm1d folded to a per-channel affine (modules.py:140-148)
                pre = "postnet.convolutions.%d.1." % i
                if pre + "weight" in P:
                    scale = P[pre + "weight"] / torch.sqrt(P[pre + "running_var"] + 1e-5)
                    shift = P[pre + "bias"] - P[pre + "running_mean"] * scale
                    self.w[pre + "fold"] = (scale.float().contiguous(), shift.float().contiguous())
            if self.module.decoder_math == "bf16x3":
                # bf16 hi/lo planes of every weight on the tensor-core part of the path (decoder, mel head)
                for name, t in P.items():
                    if not (name.startswith("decoder.") or name.startswith("postnet.") or name.startswith("mel_linear.")):
                        continue
                    if name.endswith("weight") and t.dim() in (2, 3) and t.shape[1] != 1:
                        src = self.w[name] if t.dim() == 3 else _f32(t)
                        self.w[name + "#planes"] = split_planes(src)
            if self.module.encoder_math == "bf16x6":
                # 3-plane (24-bit) copies of the weights upstream of the quantisers: encoder + variance predictors
                for name, t in P.items():
                    if not (name.startswith("encoder.") or name.startswith("variance_adaptor.")):
                        continue
                    if name.endswith("weight") and t.dim() in (2, 3) and t.shape[1] != 1 and t.shape[0] % 4 == 0 \
                            and "embed" not in name and "emb" not in name.split(".")[-2]:
                        src = self.w[name] if t.dim() == 3 else _f32(t)
                        if src.shape[1] % 8 == 0:
                            self.w[name + "#planes3"] = split_planes(src, 3)
            block = self.module.model_config["block_type"]
            if block != "transformer_fs2":
                from . import engine_blocks
                engine_blocks.PREPARE[block](self, P, self.module.decoder_math == "bf16x3")
            self.w["cwt_scale_w"] = ((torch.arange(0, 10).float() + 1 + 2.5) ** (-2.5)).to(
                next(iter(P.values())).device)  # utils/pitch_tools.py:260

    def table_fs2(self, dim, rows, device):
        """fairseq-style [sin | cos] table (blocks.py:66-83); row 0 zero; grows on demand (:88-95)."""
        key = ("fs2", dim, str(device))
        tab = self.pe.get(key)
        if tab is None or tab.shape[0] < rows:
            if tab is not None:
                # captured CUDA graphs hold the old table's address: keep it alive (its rows are a prefix of the new one)
                self.pe_retired.append(tab)
                rows = max(rows, 2 * tab.shape[0])     # grow geometrically: regrowth stays rare
            rows = max(rows, 2048)
            half = dim // 2
            step = math.log(10000) / (half - 1)
            freq = torch.exp(torch.arange(half, dtype=torch.float) * -step)
            ang = torch.arange(rows, dtype=torch.float).unsqueeze(1) * freq.unsqueeze(0)
            tab = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
            tab[0, :] = 0
            tab = tab.to(device).contiguous()
            self.pe[key] = tab
        return tab


# ---------------------------------------------------------------------------------------------
# transformer_fs2 blocks
# ---------------------------------------------------------------------------------------------
def _fft_layers_fs2(prep, P, pre, x, lens, n_layers, n_head, kernel, act):
    """EncSALayer x n + final LayerNorm (transformer_fs2.py:60-66,176-200).  x is updated in place."""
    W = prep.w
    for i in range(n_layers):
        lp = "%slayers.%d.op." % (pre, i)
        h = layernorm(x, P[lp + "layer_norm1.weight"], P[lp + "layer_norm1.bias"], 1e-12)
        qkv = conv_gemm(h, P[lp + "self_attn.in_proj_weight"])
        a = attention(qkv, lens, n_head)
        conv_gemm(a, P[lp + "self_attn.out_proj.weight"], residual=x, lens=lens, out=x)
        h = layernorm(x, P[lp + "layer_norm2.weight"], P[lp + "layer_norm2.bias"], 1e-12)
        f = conv_gemm(h, W[lp + "ffn.ffn_1.weight"], P[lp + "ffn.ffn_1.bias"], alpha=kernel ** -0.5, act=act,
                      taps=kernel)
        conv_gemm(f, P[lp + "ffn.ffn_2.weight"], P[lp + "ffn.ffn_2.bias"], residual=x, lens=lens, out=x)
    return layernorm(x, P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"], 1e-5, lens)


def encoder_fs2(prep, P, cfg, tokens, src_lens, math_mode="fp32"):
    """TextEncoder.forward, transformer_fs2.py:100-119."""
    c = cfg["transformer_fs2"]
    B, S = tokens.shape
    C = c["encoder_hidden"]
    table = P["encoder.embed_tokens.weight"]
    pe = prep.table_fs2(C, S + 1, tokens.device)
    x = torch.empty(B, S, C, device=tokens.device, dtype=torch.float32)
    word = torch.empty_like(x)
    capi.call("ctts_embed_tokens", tokens, table, pe, pe.shape[0], math.sqrt(C), B, S, C, table.shape[0], x, word,
              src_lens, 0, _stream())
    act = _ACTS[cfg["variance_predictor"]["ffn_act"]]
    if math_mode == "bf16x6":
        x, _ = _fft_layers_fs2_tc(prep, P, "encoder.", x, src_lens, c["encoder_layer"], c["encoder_head"],
                                  c["ffn_kernel_size"], act, n=3)
    else:
        x = _fft_layers_fs2(prep, P, "encoder.", x, src_lens, c["encoder_layer"], c["encoder_head"],
                            c["ffn_kernel_size"], act)
    return x, word


def _fft_layers_fs2_tc(prep, P, pre, x, lens, n_layers, n_head, kernel, act, n=2):
    """Same block as _fft_layers_fs2 with the dense contractions and the attention on tcgen05.  n = 2: bf16x3 operands
    (decoder); n = 3: bf16x6 operands (FP32-equivalent, encoder).
    LayerNorm, softmax and the residual stream stay FP32.  Returns (final LN fp32, final LN planes)."""
    W = prep.w
    tag = "#planes" if n == 2 else "#planes3"
    if n == 2 and FUSED_LN and x.shape[-1] == 256:
        # every LayerNorm but the first rides in the epilogue of the projection in front of it: 5 launches per block
        _, hp = layernorm_planes(x, P[pre + "layers.0.op.layer_norm1.weight"], P[pre + "layers.0.op.layer_norm1.bias"], 1e-12, n=2)
        for i in range(n_layers):
            lp = "%slayers.%d.op." % (pre, i)
            _, qkvp = gemm_tc(hp, W[lp + "self_attn.in_proj_weight" + tag], want_fp32=False, want_planes=True)
            ap = attention_tc(qkvp, lens, n_head)
            _, hp = gemm_tc_ln(ap, W[lp + "self_attn.out_proj.weight" + tag], None, x, lens, x, P[lp + "layer_norm2.weight"],
                               P[lp + "layer_norm2.bias"], 1e-12)
            _, fp = gemm_tc(hp, W[lp + "ffn.ffn_1.weight" + tag], P[lp + "ffn.ffn_1.bias"], alpha=kernel ** -0.5, act=act,
                            taps=kernel, want_fp32=False, want_planes=True)
            if i + 1 < n_layers:
                nx = "%slayers.%d.op." % (pre, i + 1)
                _, hp = gemm_tc_ln(fp, W[lp + "ffn.ffn_2.weight" + tag], P[lp + "ffn.ffn_2.bias"], x, lens, x,
                                   P[nx + "layer_norm1.weight"], P[nx + "layer_norm1.bias"], 1e-12)
            else:
                return gemm_tc_ln(fp, W[lp + "ffn.ffn_2.weight" + tag], P[lp + "ffn.ffn_2.bias"], x, lens, x,
                                  P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"], 1e-5, ln_lens=lens, want_fp32=True)
    for i in range(n_layers):
        lp = "%slayers.%d.op." % (pre, i)
        _, hp = layernorm_planes(x, P[lp + "layer_norm1.weight"], P[lp + "layer_norm1.bias"], 1e-12, n=n)
        _, qkvp = gemm_tc(hp, W[lp + "self_attn.in_proj_weight" + tag], want_fp32=False, want_planes=True)
        ap = attention_tc(qkvp, lens, n_head)
        gemm_tc(ap, W[lp + "self_attn.out_proj.weight" + tag], residual=x, lens=lens, out=x)
        _, hp = layernorm_planes(x, P[lp + "layer_norm2.weight"], P[lp + "layer_norm2.bias"], 1e-12, n=n)
        _, fp = gemm_tc(hp, W[lp + "ffn.ffn_1.weight" + tag], P[lp + "ffn.ffn_1.bias"], alpha=kernel ** -0.5, act=act,
                        taps=kernel, want_fp32=False, want_planes=True)
        gemm_tc(fp, W[lp + "ffn.ffn_2.weight" + tag], P[lp + "ffn.ffn_2.bias"], residual=x, lens=lens, out=x)
    return layernorm_planes(x, P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"], 1e-5, lens, want_fp32=True, n=n)


def decoder_fs2(prep, P, cfg, x, mel_lens, math_mode="fp32"):
    """Decoder = FFTBlocks with learnable-scale sinusoid positions, transformer_fs2.py:47-72,122-134.
    `x` must be a private buffer: it is overwritten.  Returns (dec fp32, dec planes or None)."""
    c = cfg["transformer_fs2"]
    B, T, C = x.shape
    pe = prep.table_fs2(C, T + 1, x.device)
    xin, x = x, torch.empty_like(x)
    capi.call("ctts_add_positions", xin, pe, pe.shape[0], P["decoder.pos_embed_alpha"], mel_lens, B, T, C, 0, x, _stream())
    act = _ACTS[cfg["variance_predictor"]["ffn_act"]]
    if math_mode == "bf16x3":
        return _fft_layers_fs2_tc(prep, P, "decoder.", x, mel_lens, c["decoder_layer"], c["decoder_head"],
                                  c["ffn_kernel_size"], act)
    return _fft_layers_fs2(prep, P, "decoder.", x, mel_lens, c["decoder_layer"], c["decoder_head"],
                           c["ffn_kernel_size"], act), None


# ---------------------------------------------------------------------------------------------
# variance adaptor
# ---------------------------------------------------------------------------------------------
def _predictor_stack(prep, P, pre, x, n_layers, kernel, lens):
    """[pad, Conv1d, ReLU, LayerNorm(channels), Dropout] x n (modules.py:1277-1288,1330-1338).
    On tcgen05 with 3-plane (FP32-equivalent) operands when the module's encoder_math is "bf16x6"."""
    if ("%sconv.0.1.weight#planes3" % pre) in prep.w:
        xp = split_planes(x, 3)
        for l in range(n_layers):
            h, _ = gemm_tc(xp, prep.w["%sconv.%d.1.weight#planes3" % (pre, l)], P["%sconv.%d.1.bias" % (pre, l)],
                           act=ACT_RELU, taps=kernel)
            last = l == n_layers - 1
            x, xp = layernorm_planes(h, P["%sconv.%d.3.weight" % (pre, l)], P["%sconv.%d.3.bias" % (pre, l)], 1e-12, lens,
                                     want_fp32=last, n=3)
        return x
    for l in range(n_layers):
        h = conv_gemm(x, prep.w["%sconv.%d.1.weight" % (pre, l)], P["%sconv.%d.1.bias" % (pre, l)], act=ACT_RELU,
                      taps=kernel)
        x = layernorm(h, P["%sconv.%d.3.weight" % (pre, l)], P["%sconv.%d.3.bias" % (pre, l)], 1e-12, lens)
    return x


def duration_predictor(prep, P, cfg, x, src_lens):
    """DurationPredictor.forward, modules.py:1299-1310 -> log-durations [B, S]."""
    vp = cfg["variance_predictor"]
    pre = "variance_adaptor.duration_predictor."
    h = _predictor_stack(prep, P, pre, x, vp["dur_predictor_layers"], vp["dur_predictor_kernel"], src_lens)
    out = conv_gemm(h, P[pre + "linear.weight"], P[pre + "linear.bias"], lens=src_lens)
    return out.squeeze(-1)


def pitch_style_predictor(prep, P, cfg, pre, xs, alpha=1.0):
    """PitchPredictor / EnergyPredictor.forward, modules.py:1343-1356."""
    vp = cfg["variance_predictor"]
    B, T, C = xs.shape
    pe = prep.table_fs2(C, T + 1, xs.device)
    xp = torch.empty_like(xs)
    capi.call("ctts_add_positions", xs, pe, pe.shape[0], P[pre + "pos_embed_alpha"], None, B, T, C, 0, xp, _stream())
    h = _predictor_stack(prep, P, pre, xp, vp["predictor_layers"], vp["predictor_kernel"], None)
    return conv_gemm(h, P[pre + "linear.weight"], P[pre + "linear.bias"], alpha=alpha)


def prosody_predictor(prep, P, cfg, pre, x, phoneme_level):
    """ParallelProsodyPredictor.forward, modules.py:630-648: (conv k3 + ReLU + LayerNorm 1e-5) x 2, bi-GRU over all S
    steps, bottleneck Linear.  FP32 (it feeds the duration / pitch / energy quantisers)."""
    k = cfg["prosody_modeling"]["liu2021"]["predictor_kernel_size"]
    B, S, E = x.shape
    st = _stream()
    h = x
    for i in (1, 2):
        if i == 2 and k != 3:
            raise NotImplementedError("conv1d_2 uses padding=1: only predictor_kernel_size 3 is 'same' (modules.py:610)")
        c = conv_gemm(h, prep.w[pre + "conv_layer.conv1d_%d.conv.weight" % i], P[pre + "conv_layer.conv1d_%d.conv.bias" % i],
                      act=ACT_RELU, taps=k)
        h = layernorm(c, P[pre + "conv_layer.layer_norm_%d.weight" % i], P[pre + "conv_layer.layer_norm_%d.bias" % i], 1e-5)
    gi_f = conv_gemm(h, P[pre + "gru.weight_ih_l0"], P[pre + "gru.bias_ih_l0"])
    gi_b = conv_gemm(h, P[pre + "gru.weight_ih_l0_reverse"], P[pre + "gru.bias_ih_l0_reverse"])
    Hd = P[pre + "gru.weight_hh_l0"].shape[1]
    mem = torch.empty(B, S, 2 * Hd, device=x.device, dtype=torch.float32)
    h_fin = torch.empty(B, 2 * Hd, device=x.device, dtype=torch.float32)
    capi.call("ctts_gru_bidir", gi_f, gi_b, P[pre + "gru.weight_hh_l0"], P[pre + "gru.bias_hh_l0"],
              P[pre + "gru.weight_hh_l0_reverse"], P[pre + "gru.bias_hh_l0_reverse"], B, S, Hd, mem, h_fin, st)
    if phoneme_level:
        return conv_gemm(mem, P[pre + "predictor_bottleneck.weight"], P[pre + "predictor_bottleneck.bias"])
    return conv_gemm(h_fin.view(1, B, 2 * Hd), P[pre + "predictor_bottleneck.weight"],
                     P[pre + "predictor_bottleneck.bias"]).view(B, 1, -1)


def alignment_encoder(prep, P, cfg, mel, text_embedding, src_lens, attn_prior, spk):
    """AlignmentEncoder.forward, modules.py:1176-1213.  mel [B,M,80], text_embedding [B,S,C] (token-major), attn_prior
    [B,S,M].  Returns (attn_soft, attn_logprob) as [B,1,M,S]."""
    pre = "variance_adaptor.aligner."
    B, M, _ = mel.shape
    S = text_embedding.shape[1]
    st = _stream()
    keys, queries = text_embedding, mel
    if spk is not None:
        ks = conv_gemm(_f32(spk).view(1, B, -1), P[pre + "key_spk_proj.linear.weight"]).view(B, -1)
        qs = conv_gemm(_f32(spk).view(1, B, -1), P[pre + "query_spk_proj.linear.weight"]).view(B, -1)
        k2 = torch.empty_like(keys)
        capi.call("ctts_add_row_broadcast", keys, ks, B, S, keys.shape[2], k2, st)
        q2 = torch.empty_like(queries)
        capi.call("ctts_add_row_broadcast", queries, qs, B, M, queries.shape[2], q2, st)
        keys, queries = k2, q2
    k = conv_gemm(keys, prep.w[pre + "key_proj.0.conv.weight"], P[pre + "key_proj.0.conv.bias"], act=ACT_RELU, taps=3)
    k = conv_gemm(k, prep.w[pre + "key_proj.2.conv.weight"], P[pre + "key_proj.2.conv.bias"])
    q = conv_gemm(queries, prep.w[pre + "query_proj.0.conv.weight"], P[pre + "query_proj.0.conv.bias"], act=ACT_RELU, taps=3)
    q = conv_gemm(q, prep.w[pre + "query_proj.2.conv.weight"], P[pre + "query_proj.2.conv.bias"], act=ACT_RELU)
    q = conv_gemm(q, prep.w[pre + "query_proj.4.conv.weight"], P[pre + "query_proj.4.conv.bias"])
    soft = torch.empty(B, 1, M, S, device=mel.device, dtype=torch.float32)
    logprob = torch.empty(B, 1, M, S, device=mel.device, dtype=torch.float32)
    capi.call("ctts_aligner_attention", q, k, attn_prior, src_lens, float(cfg["duration_modeling"]["aligner_temperature"]),
              B, M, S, q.shape[2], soft, logprob, st)
    return soft, logprob


def length_regulate(x, dur, src_lens, max_len, need_mel2ph, expand=True):
    """LengthRegulator (modules.py:1222-1249) + dur_to_mel2ph (utils/tools.py:598-628).

    Returns (expanded [B,M,C], mel_len [B] i64, mel2ph [B,M2] i64 or None, cum_lr).  When max_len is
    None the two maxima are read back from the device: the only host sync of the forward path."""
    B, S, C = x.shape
    dev = x.device
    cum_lr = torch.empty(B, S, device=dev, dtype=torch.int32)
    cum_m2p = torch.empty(B, S, device=dev, dtype=torch.int32)
    lens2 = torch.empty(2 * B, device=dev, dtype=torch.int64)
    if dur.is_floating_point():
        capi.call("ctts_length_scan", _f32(dur), None, src_lens, B, S, cum_lr, cum_m2p, lens2, _stream())
    else:
        capi.call("ctts_length_scan", None, _i64(dur), src_lens, B, S, cum_lr, cum_m2p, lens2, _stream())
    mel_len = lens2[:B]
    if max_len is None or need_mel2ph:
        maxes = lens2.view(2, B).max(dim=1).values.tolist()  # host sync (D2H of two integers)
        M = int(max_len) if max_len is not None else int(maxes[0])
        M2 = int(maxes[1])
    else:
        M, M2 = int(max_len), 0
    if M <= 0:
        raise capi.CttsError("length_regulate: every duration is zero (empty mel)")
    mel2ph = torch.empty(B, M2, device=dev, dtype=torch.int64) if (need_mel2ph and M2 > 0) else None
    if expand:
        out = torch.empty(B, M, C, device=dev, dtype=torch.float32)
        capi.call("ctts_length_expand", x, None, None, cum_lr, B, S, C, M, 0, out, cum_m2p, mel2ph,
                  M2 if mel2ph is not None else 0, _stream())
    else:  # only the frame -> phoneme map is wanted (soft-upsampling branch)
        out = torch.empty(B, 1, C, device=dev, dtype=torch.float32)
        capi.call("ctts_length_expand", x, None, None, cum_lr, B, S, C, 1, 0, out, cum_m2p, mel2ph,
                  M2 if mel2ph is not None else 0, _stream())
    return out, mel_len, mel2ph, cum_lr


def length_scan(dur, src_lens, B, S, device):
    """ctts_length_scan: (cum_lr, cum_m2p, lens2) with lens2[:B] = LR lengths, lens2[B:] = mel2ph lengths."""
    cum_lr = torch.empty(B, S, device=device, dtype=torch.int32)
    cum_m2p = torch.empty(B, S, device=device, dtype=torch.int32)
    lens2 = torch.empty(2 * B, device=device, dtype=torch.int64)
    if dur.is_floating_point():
        capi.call("ctts_length_scan", _f32(dur), None, src_lens, B, S, cum_lr, cum_m2p, lens2, _stream())
    else:
        capi.call("ctts_length_scan", None, _i64(dur), src_lens, B, S, cum_lr, cum_m2p, lens2, _stream())
    return cum_lr, cum_m2p, lens2


def variance_stage_a(prep, P, pcfg, cfg, tcfg, spk, text, text_embedding, src_lens, mel, mel_lens, duration_target,
                     attn_prior, d_control):
    """VarianceAdaptor.forward up to the point where the regulated length must be known on the host
    (modules.py:962-1063): speaker add, liu2021 predictors, duration predictor, aligner + MAS, duration decoding and the
    LengthRegulator scan.  No host synchronisation: capturable in a CUDA graph."""
    B, S, C = text.shape
    st = _stream()
    if spk is not None:
        x = torch.empty_like(text)
        capi.call("ctts_add_row_broadcast", text, _f32(spk), B, S, C, x, st)
    else:
        x = text
    prosody_info = None
    model_type = cfg["prosody_modeling"]["model_type"]
    if model_type == "liu2021":
        # eval mode: the parallel predictors stand in for the reference encoders (modules.py:1002-1023)
        pre = "variance_adaptor."
        u_vec = prosody_predictor(prep, P, cfg, pre + "utterance_prosody_predictor.", x, False)       # [B, 1, 256]
        u_add = conv_gemm(u_vec.view(1, B, -1), P[pre + "utterance_prosody_prj.weight"],
                          P[pre + "utterance_prosody_prj.bias"]).view(B, -1)
        x2 = torch.empty_like(x)
        capi.call("ctts_add_row_broadcast", x, u_add, B, S, C, x2, st)
        p_vec = prosody_predictor(prep, P, cfg, pre + "phoneme_prosody_predictor.", x2, True)        # [B, S, 4]
        x = torch.empty_like(x2)
        capi.call("ctts_linear_smallk", p_vec, P[pre + "phoneme_prosody_prj.weight"], P[pre + "phoneme_prosody_prj.bias"],
                  x2, B * S, p_vec.shape[-1], C, x, st)
        prosody_info = (None, None, u_vec, p_vec, None)
    elif model_type != "none":
        raise NotImplementedError("prosody model %r" % model_type)
    log_d = duration_predictor(prep, P, cfg, x, src_lens)

    attn_out = (None, None, None, None)
    if attn_prior is not None:
        # unsupervised duration modelling: AlignmentEncoder + monotonic alignment search (modules.py:1031-1053)
        assert cfg["duration_modeling"]["learn_alignment"] and duration_target is None and mel is not None
        attn_soft, attn_logprob = alignment_encoder(prep, P, cfg, _f32(mel), text_embedding, src_lens, _f32(attn_prior), spk)
        M_in = mel.shape[1]
        prev_ws = torch.empty(B * M_in * S, device=x.device, dtype=torch.uint8)
        attn_hard = torch.empty(B, 1, M_in, S, device=x.device, dtype=torch.float32)
        attn_hard_dur = torch.empty(B, S, device=x.device, dtype=torch.float32)
        capi.call("ctts_mas", attn_soft, src_lens, mel_lens, B, M_in, S, prev_ws, attn_hard, attn_hard_dur, st)
        attn_out = (attn_soft, attn_hard, attn_hard_dur, attn_logprob)
        duration_rounded = attn_hard_dur
    elif duration_target is not None:
        assert not cfg["duration_modeling"]["learn_alignment"] and attn_prior is None
        duration_rounded = duration_target
    else:
        duration_rounded = torch.empty_like(log_d)
        capi.call("ctts_decode_durations", log_d, float(d_control), B * S, duration_rounded, st)
    cum_lr, cum_m2p, lens2 = length_scan(duration_rounded, src_lens, B, S, x.device)
    return dict(x=x, log_d=log_d, duration_rounded=duration_rounded, cum_lr=cum_lr, cum_m2p=cum_m2p, lens2=lens2,
                attn_out=attn_out, prosody_info=prosody_info)


def variance_stage_b(prep, P, pcfg, cfg, tcfg, a, src_lens, mel_lens, mel_mask, max_len, M, M2, pitch_target,
                     energy_target, attn_prior, p_control, e_control, step):
    """The rest of VarianceAdaptor.forward once the regulated length M (and the mel2ph length M2) are known
    (modules.py:1044-1114): upsampling, pitch / energy embeddings."""
    pitch_cfg = pcfg["preprocessing"]["pitch"]
    x = a["x"]
    x_org = x
    B, S, C = x.shape
    st = _stream()
    dev = x.device
    cum_lr, cum_m2p = a["cum_lr"], a["cum_m2p"]
    need_m2p = attn_prior is not None or a.get("free_running", False)
    mel2ph = torch.empty(B, M2, device=dev, dtype=torch.int64) if (need_m2p and M2 > 0) else None
    soft = attn_prior is not None and step < tcfg["duration"]["binarization_start_steps"]
    if soft:
        # soft upsampling x = bmm(A_soft, x) (modules.py:1047-1049); mel_len stays the caller's
        attn_soft = a["attn_out"][0]
        M_in = attn_soft.shape[2]
        Sp = (S + 15) // 16 * 16
        a_pad = torch.zeros(B, M_in, Sp, device=dev, dtype=torch.float32)
        capi.call("ctts_copy_rows", attn_soft, S, B * M_in, S, a_pad, Sp, 0, st)      # re-stride [.., S] -> [.., Sp]
        xt = torch.empty(B, C, Sp, device=dev, dtype=torch.float32)
        capi.call("ctts_transpose_heads", x, B, S, C, 0, 1, C, Sp, xt, st)
        xe = torch.empty(B, M_in, C, device=dev, dtype=torch.float32)
        capi.call("ctts_batched_gemm_fp32", a_pad, xt, 1.0, None, 1, B, 1, M_in, Sp, C, M_in * Sp, 0, Sp, C * Sp, 0, Sp,
                  M_in * C, 0, C, xe, st)
        dummy = torch.empty(B, 1, C, device=dev, dtype=torch.float32)   # only the frame -> phoneme map is wanted
        capi.call("ctts_length_expand", x, None, None, cum_lr, B, S, C, 1, 0, dummy, cum_m2p, mel2ph,
                  M2 if mel2ph is not None else 0, st)
        mel_len = mel_lens
    else:
        xe = torch.empty(B, M, C, device=dev, dtype=torch.float32)
        capi.call("ctts_length_expand", x, None, None, cum_lr, B, S, C, M, 0, xe, cum_m2p, mel2ph,
                  M2 if mel2ph is not None else 0, st)
        mel_len = a["lens2"][:B]
    if attn_prior is not None:
        m2p = mel2ph if mel2ph is not None else torch.zeros(B, 0, device=dev, dtype=torch.int64)
        pitch_target["mel2ph"] = m2p[:, :max_len]
    if a.get("free_running", False):
        mel_mask = pad_mask(mel_len, xe.shape[1])
    M = xe.shape[1]

    x_sum = xe.clone()
    pitch_pred = energy_pred = None
    use_pitch = cfg["variance_embedding"]["use_pitch_embed"]
    use_energy = cfg["variance_embedding"]["use_energy_embed"]
    level = pcfg["preprocessing"]["energy"]["feature"]
    pre = "variance_adaptor."

    # The CWT statistics MLP (3 tiny launches on x_org[:, 0]) and the phoneme-level energy predictor (~10 launches on
    # [B, S] rows) are independent of the frame-level CWT predictor chain: while the forward is being captured into a CUDA
    # graph they are issued on a side stream (fork / join events become parallel graph branches), so their latency-bound
    # kernels overlap the longer chain instead of queueing behind it.
    side = {}

    pitch_type = pitch_cfg["pitch_type"]

    def side_work():
        if use_pitch and pitch_type == "cwt":
            first = torch.empty(1, B, C, device=dev, dtype=torch.float32)
            capi.call("ctts_copy_rows", x_org, S * C, B, C, first, C, 0, _stream())
            s = conv_gemm(first, P[pre + "cwt_stats_layers.0.weight"], P[pre + "cwt_stats_layers.0.bias"], act=ACT_RELU)
            s = conv_gemm(s, P[pre + "cwt_stats_layers.2.weight"], P[pre + "cwt_stats_layers.2.bias"], act=ACT_RELU)
            side["stats"] = conv_gemm(s, P[pre + "cwt_stats_layers.4.weight"], P[pre + "cwt_stats_layers.4.bias"]).view(B, 2)
        if use_energy and level != "frame_level":
            et = energy_target
            if attn_prior is not None:  # frame-level target -> phoneme level by the hard durations (modules.py:1096-1097)
                etf = _f32(energy_target)
                M_e = etf.shape[1]
                work = torch.empty(B * M_e, device=dev, dtype=torch.float32)
                et = torch.empty(B, S, device=dev, dtype=torch.float32)
                capi.call("ctts_phoneme_energy", a["attn_out"][2], src_lens, etf, B, S, M_e, work, et, _stream())
            pred = pitch_style_predictor(prep, P, cfg, pre + "energy_predictor.", x_org,
                                         alpha=1.0 if et is not None else e_control).squeeze(-1)
            src_vals = _f32(et) if et is not None else pred
            eidx = torch.empty(B, S, device=dev, dtype=torch.int64)
            bins = P[pre + "energy_bins"]
            capi.call("ctts_bucketize", src_vals, 1.0, bins, bins.shape[0], B * S, eidx, _stream())
            side["energy"] = (et, pred, eidx)

    forked = None
    if PARALLEL_BRANCHES and dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
        main = torch.cuda.current_stream()
        branch = _side_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main)
        with torch.cuda.stream(branch):
            branch.wait_event(fork)
            side_work()
            forked = torch.cuda.Event()
            forked.record(branch)
    else:
        side_work()

    if use_pitch and pitch_type != "cwt":
        if forked is not None:
            torch.cuda.current_stream().wait_event(forked)
            forked = None
        pitch_pred = _pitch_frame_or_ph(prep, P, cfg, pitch_cfg, pitch_type, xe, x_org, x_sum, pitch_target, mel2ph, src_lens,
                                        mel_len, p_control, B, S, M, C)
    if use_pitch and pitch_type == "cwt":
        if (pre + "cwt_predictor.0.weight#planes3") in prep.w:
            h, _ = gemm_tc(split_planes(xe, 3), prep.w[pre + "cwt_predictor.0.weight#planes3"],
                           P[pre + "cwt_predictor.0.bias"])
        else:
            h = conv_gemm(xe, P[pre + "cwt_predictor.0.weight"], P[pre + "cwt_predictor.0.bias"])
        cwt = pitch_style_predictor(prep, P, cfg, pre + "cwt_predictor.1.", h, alpha=p_control)
    if forked is not None:
        torch.cuda.current_stream().wait_event(forked)      # join
    if use_pitch and pitch_type == "cwt":
        stats = side["stats"]
        f0_denorm = torch.empty(B, M, device=dev, dtype=torch.float32)
        idx = torch.empty(B, M, device=dev, dtype=torch.int64)
        use_uv = 1 if pitch_cfg["use_uv"] else 0
        assert pitch_cfg["pitch_norm"] == "log", "only pitch_norm 'log' (the shipped configs) is built"
        if pitch_target is not None:
            m2p = pitch_target["mel2ph"]
            assert m2p.shape[1] == M, "mel2ph length %d != regulated length %d" % (m2p.shape[1], M)
            f0n = torch.empty(B, M, device=dev, dtype=torch.float32)
            spec = _f32(pitch_target["cwt_spec"])
            capi.call("ctts_cwt_to_pitch", spec, spec.shape[-1], prep.w["cwt_scale_w"], _f32(pitch_target["f0_mean"]),
                      _f32(pitch_target["f0_std"]), 1, 1.0, float(pitch_cfg["pitch_norm_eps"]),
                      _f32(pitch_target["uv"]), use_uv, B, M, f0n, f0_denorm, idx, st)
            pitch_target["f0"] = f0n
            pitch_target["f0_cwt"] = f0n
        else:
            assert mel2ph is not None and mel2ph.shape[1] == M, "mel2ph / regulated length mismatch"
            capi.call("ctts_cwt_to_pitch", cwt, cwt.shape[-1], prep.w["cwt_scale_w"], stats, stats[:, 1:], 2,
                      float(cfg["variance_predictor"]["cwt_std_scale"]), float(pitch_cfg["pitch_norm_eps"]), None,
                      use_uv, B, M, None, f0_denorm, idx, st)
        emb = P[pre + "pitch_embed.weight"]
        capi.call("ctts_gather_add", emb, idx, B * M, C, emb.shape[0], x_sum, st)
        pitch_pred = {"pitch_pred": None, "f0_denorm": f0_denorm, "cwt": cwt, "f0_mean": stats[:, 0],
                      "f0_std": stats[:, 1]}
    if use_energy:
        bins = P[pre + "energy_bins"]
        emb = P[pre + "energy_embedding.weight"]
        if level == "frame_level":
            pred = pitch_style_predictor(prep, P, cfg, pre + "energy_predictor.", xe,
                                         alpha=1.0 if energy_target is not None else e_control).squeeze(-1)
            src_vals = _f32(energy_target) if energy_target is not None else pred
            eidx = torch.empty(B, M, device=dev, dtype=torch.int64)
            capi.call("ctts_bucketize", src_vals, 1.0, bins, bins.shape[0], B * M, eidx, st)
            capi.call("ctts_gather_add", emb, eidx, B * M, C, emb.shape[0], x_sum, st)
        else:
            energy_target, pred, eidx = side["energy"]
            capi.call("ctts_length_expand", None, emb, eidx, cum_lr, B, S, C, M, 1, x_sum, None, None, 0, st)
        energy_pred = pred
    return (x_sum, pitch_target, pitch_pred, energy_target, energy_pred, mel_len, mel_mask)


def _pitch_frame_or_ph(prep, P, cfg, pitch_cfg, pitch_type, xe, x_org, x_sum, pitch_target, mel2ph, src_lens, mel_len, p_control,
                       B, S, M, C):
    """get_pitch_embedding for pitch_type 'frame' / 'ph' (modules.py:890-906,927-938): one PitchPredictor on the frame- or
    phoneme-level input; adds the pitch embedding to x_sum and returns the prediction dict."""
    pre = "variance_adaptor."
    st = _stream()
    dev = xe.device
    assert pitch_cfg["pitch_norm"] == "log", "only pitch_norm 'log' (the shipped configs) is built"
    m2p = pitch_target["mel2ph"] if pitch_target is not None else mel2ph
    assert m2p is not None and m2p.shape[1] == M, "mel2ph / regulated length mismatch"
    m2p = _i64(m2p)
    emb = P[pre + "pitch_embed.weight"]
    idx = torch.empty(B, M, device=dev, dtype=torch.int64)
    if pitch_type == "frame":
        pred = pitch_style_predictor(prep, P, cfg, pre + "pitch_predictor.", xe, alpha=p_control)         # [B, M, 2]
        f0_denorm = torch.empty(B, M, device=dev, dtype=torch.float32)
        if pitch_target is not None:
            f0t = _f32(pitch_target["f0"])
            capi.call("ctts_frame_pitch", None, 0, f0t, _f32(pitch_target["uv"]), m2p, 1 if pitch_cfg["use_uv"] else 0, B * M,
                      f0t, f0_denorm, idx, st)
            pitch_target["f0"] = f0t
        else:
            f0 = torch.empty(B, M, device=dev, dtype=torch.float32)
            capi.call("ctts_frame_pitch", pred, pred.shape[-1], None, None, m2p, 1 if pitch_cfg["use_uv"] else 0, B * M, f0,
                      f0_denorm, idx, st)
    else:
        pred = pitch_style_predictor(prep, P, cfg, pre + "pitch_predictor.", x_org, alpha=p_control)      # [B, S, 1]
        if pitch_target is not None:
            f0 = torch.empty(B, S, device=dev, dtype=torch.float32)
            capi.call("ctts_phoneme_pitch", _f32(pitch_target["f0"]), m2p, src_lens, _i64(mel_len), B, S, M, f0, st)
            pitch_target["f0"] = f0
        else:
            f0 = pred.view(B, S)
        f0_denorm = torch.empty(B, S, device=dev, dtype=torch.float32)
        idx_ph = torch.empty(B, S, device=dev, dtype=torch.int64)
        capi.call("ctts_f0_to_pitch", f0, None, B * S, f0_denorm, idx_ph, st)     # pitch_padding is a scalar False (:894)
        capi.call("ctts_gather_index", idx_ph, m2p, B, S, M, idx, st)
    capi.call("ctts_gather_add", emb, idx, B * M, C, emb.shape[0], x_sum, st)
    return {"pitch_pred": pred, "f0_denorm": f0_denorm, "cwt": None, "f0_mean": None, "f0_std": None}


# ---------------------------------------------------------------------------------------------
def mel_head(prep, P, dec, dec_planes=None):
    """mel_linear + PostNet + residual (CompTransTTS.py:133-135, modules.py:140-148; eval-mode BN folded)."""
    if dec_planes is not None:
        mel, hp = gemm_tc(dec_planes, prep.w["mel_linear.weight#planes"], P["mel_linear.bias"], want_planes=True)
        post = None
        for i in range(5):
            pre = "postnet.convolutions.%d." % i
            post, hp = gemm_tc(hp, prep.w[pre + "0.conv.weight#planes"], P[pre + "0.conv.bias"], bn=prep.w[pre + "1.fold"],
                               act=ACT_TANH if i < 4 else ACT_NONE, residual=mel if i == 4 else None, taps=5,
                               want_fp32=(i == 4), want_planes=(i < 4))
        return mel, post
    mel = conv_gemm(dec, P["mel_linear.weight"], P["mel_linear.bias"])
    h = mel
    for i in range(5):
        pre = "postnet.convolutions.%d." % i
        h = conv_gemm(h, prep.w[pre + "0.conv.weight"], P[pre + "0.conv.bias"], bn=prep.w[pre + "1.fold"],
                      act=ACT_TANH if i < 4 else ACT_NONE, residual=mel if i == 4 else None, taps=5)
    return mel, h


# ---------------------------------------------------------------------------------------------
# CUDA-graph plumbing: the forward is two capturable stages around its single host sync
# ---------------------------------------------------------------------------------------------
def _flatten(prefix, v, out):
    if v is None:
        return
    if torch.is_tensor(v):
        out[prefix] = v
    elif isinstance(v, dict):
        for k in v:
            _flatten(prefix + "." + k if prefix else k, v[k], out)
    elif isinstance(v, (tuple, list)):
        for i, x in enumerate(v):
            _flatten("%s.%d" % (prefix, i), x, out)


def _tree_map(fn, v):
    if torch.is_tensor(v):
        return fn(v)
    if isinstance(v, dict):
        return {k: _tree_map(fn, x) for k, x in v.items()}
    if isinstance(v, tuple):
        return tuple(_tree_map(fn, x) for x in v)
    if isinstance(v, list):
        return [_tree_map(fn, x) for x in v]
    return v


class GraphCache:
    """Shape-keyed cache of captured stages.  A stage is run eagerly the first time a key is seen (this also warms the
    lazily built tables and kernel attributes), captured into a CUDA graph the second time, and replayed afterwards:
    at the bench shape the step is otherwise bound by ~120 Python/ctypes launches, not by the GPU."""

    def __init__(self, max_entries=16):
        self.entries = {}
        self.max_entries = max_entries
        self.pool = None

    def clear(self):
        self.entries.clear()

    def run(self, key, fn, tensor_inputs, table=None):
        """fn(inputs_dict) -> pytree.  Returns (outputs, sub_table or None).  Outputs of a replayed graph are static
        buffers (valid until the next replay of the same key)."""
        table = self.entries if table is None else table
        e = table.get(key)
        if e is None:
            if len(table) >= self.max_entries:
                table.pop(next(iter(table)))
            table[key] = False
            return fn(tensor_inputs), None
        if e is False:
            static = {k: v.clone() for k, v in tensor_inputs.items()}
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            if self.pool is None:
                self.pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(g, pool=self.pool):
                out = fn(static)
            e = table[key] = (g, static, out, {})
        g, static, out, sub = e
        table[key] = table.pop(key)          # LRU: a hit moves the entry to the young end, eviction takes the oldest
        for k, v in tensor_inputs.items():
            static[k].copy_(v, non_blocking=True)
        g.replay()
        return out, sub


def _encode(module, prep, P, cfg, texts, src_lens):
    block = cfg["block_type"]
    if block == "transformer_fs2":
        return encoder_fs2(prep, P, cfg, texts, src_lens, module.encoder_math)
    from . import engine_blocks
    if block not in engine_blocks.ENCODERS:
        raise NotImplementedError("block_type %r: kernels not built yet" % block)
    return engine_blocks.ENCODERS[block](prep, P, cfg, texts, src_lens)


def _decode(module, prep, P, cfg, x, mel_lens):
    block = cfg["block_type"]
    if block == "transformer_fs2":
        dec, dec_planes = decoder_fs2(prep, P, cfg, x, mel_lens, module.decoder_math)
    else:
        from . import engine_blocks
        dec, dec_planes = engine_blocks.DECODERS[block](prep, P, cfg, x, mel_lens, module.decoder_math)
        if module.decoder_math == "bf16x3" and dec_planes is None:
            dec_planes = split_planes(dec)
    return mel_head(prep, P, dec, dec_planes)


def forward(module, speakers, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None, p_targets=None,
            e_targets=None, d_targets=None, attn_priors=None, spker_embeds=None, p_control=1.0, e_control=1.0,
            d_control=1.0, step=None):
    """CompTransTTS.forward, model/CompTransTTS.py:64-152.  Returns the reference's 14-tuple."""
    capi.require_device()
    capi.require_cuda_tensor(texts)
    pcfg, cfg, tcfg = module.preprocess_config, module.model_config, module.train_config
    prep = module._prepared
    sig_before = prep.sig
    P = prep.params()
    graphs = module._graphs if module.use_cuda_graphs else None
    if prep.sig is not sig_before:
        module._graphs.clear()   # weights were re-laid-out: captured graphs point at stale buffers (also when graphs are
        #                          switched off right now -- they may be switched on again later)
    texts = _i64(texts)
    src_lens = _i64(src_lens)
    B, S = texts.shape
    free_running = attn_priors is None and d_targets is None
    soft = attn_priors is not None and step < tcfg["duration"]["binarization_start_steps"]

    # ---- stage A: encoder, speaker / prosody, duration model, LengthRegulator scan ------------------------------------
    in_a = {}
    _flatten("", dict(speakers=speakers if module.has_speaker_emb and module.embedder_type == "none" else None,
                      texts=texts, src_lens=src_lens, mels=mels if attn_priors is not None else None,
                      mel_lens=_i64(mel_lens) if mel_lens is not None else None, d_targets=d_targets,
                      attn_priors=attn_priors,
                      spker_embeds=spker_embeds if module.has_speaker_emb and module.embedder_type != "none" else None),
             in_a)

    host_maxes = prep.host_maxes(texts.device)

    def stage_a(t):
        enc, word = _encode(module, prep, P, cfg, t["texts"], t["src_lens"])
        spk = None
        if module.has_speaker_emb:
            if module.embedder_type == "none":
                tab = P["speaker_emb.weight"]
                idx = _i64(t["speakers"])
                spk = torch.zeros(idx.shape[0], tab.shape[1], device=idx.device, dtype=torch.float32)
                capi.call("ctts_gather_add", tab, idx, idx.shape[0], tab.shape[1], tab.shape[0], spk, _stream())
            else:
                assert "spker_embeds" in t, "Speaker embedding should not be None"
                Bs = t["spker_embeds"].shape[0]
                spk = conv_gemm(_f32(t["spker_embeds"]).view(1, Bs, -1), P["speaker_emb.weight"],
                                P["speaker_emb.bias"]).view(Bs, -1)
        a = variance_stage_a(prep, P, pcfg, cfg, tcfg, spk, enc, word, t["src_lens"], t.get("mels"), t.get("mel_lens"),
                             t.get("d_targets"), t.get("attn_priors"), d_control)
        a["free_running"] = free_running
        # the two maxima the host needs (longest regulated length with / without d_control) go to pinned host memory from
        # inside the stage: the host then only waits for the stream instead of launching a reduction and a blocking copy
        host_maxes.copy_(a["lens2"].view(2, -1).max(dim=1).values, non_blocking=True)
        return a

    key_a = ("A", float(d_control), tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(in_a.items())))
    if graphs is not None:
        a, sub = graphs.run(key_a, stage_a, in_a)
    else:
        a, sub = stage_a(in_a), None

    # ---- the single host sync: the regulated lengths --------------------------------------------------------------
    need_m2p = attn_priors is not None or free_running
    if max_mel_len is None or need_m2p:
        if texts.is_cuda:
            torch.cuda.current_stream(texts.device).synchronize()
        maxes = host_maxes.tolist()
        M = int(max_mel_len) if max_mel_len is not None else int(maxes[0])
        M2 = int(maxes[1])
    else:
        M, M2 = int(max_mel_len), 0
    if soft:
        M = int(mels.shape[1])
    if M <= 0:
        raise capi.CttsError("length_regulate: every duration is zero (empty mel)")

    # ---- stage B: upsampling, pitch / energy embeddings, decoder, mel head ------------------------------------------
    in_b = {}
    _flatten("", dict(p_targets=p_targets, e_targets=e_targets, mel_lens=_i64(mel_lens) if mel_lens is not None else None),
             in_b)
    a_live = a   # static buffers of graph A when it was replayed, plain tensors otherwise

    def stage_b(t):
        pt = None
        if p_targets is not None:
            pt = {k[len("p_targets."):]: v for k, v in t.items() if k.startswith("p_targets.")}
        xs, pt, p_pred, e_t, e_pred, mel_len, mel_mask = variance_stage_b(
            prep, P, pcfg, cfg, tcfg, a_live, a_src_lens[0], t.get("mel_lens"), None, max_mel_len, M, M2, pt,
            t.get("e_targets"), attn_priors, p_control, e_control, step)
        mel, post = _decode(module, prep, P, cfg, xs, mel_len)
        return dict(mel=mel, post=post, p_pred=p_pred, e_pred=e_pred, mel_len=mel_len, mel_mask=mel_mask, p_targets=pt,
                    e_targets=e_t)

    a_src_lens = [src_lens]
    key_b = ("B", M, M2, float(p_control), float(e_control), bool(soft), max_mel_len,
             tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(in_b.items())))
    if graphs is not None and sub is not None:
        # graph B reads graph A's static outputs in place; its own inputs are the caller's targets
        a_src_lens[0] = graphs.entries[key_a][1]["src_lens"]
        o, _ = graphs.run(key_b, stage_b, in_b, table=sub)
    else:
        o = stage_b(in_b)
    # (bookkeeping masks are launched behind stage B so that their launch latency hides under it)
    src_masks = pad_mask(src_lens, max_src_len)
    mel_masks = pad_mask(_i64(mel_lens), max_mel_len) if mel_lens is not None else None
    if graphs is not None and sub is not None:
        # hand out private copies: the static buffers are overwritten by the next replay
        o = _tree_map(lambda v: v.clone(), o)
        a = _tree_map(lambda v: v.clone(), {k: a[k] for k in ("log_d", "duration_rounded", "attn_out", "prosody_info")})
    mel_masks_out = o["mel_mask"] if o["mel_mask"] is not None else mel_masks
    d_rounded = d_targets if (d_targets is not None and attn_priors is None) else a["duration_rounded"]
    return (o["mel"], o["post"], o["p_pred"], o["e_pred"], a["log_d"], d_rounded, src_masks, mel_masks_out, src_lens,
            o["mel_len"], a["attn_out"], a["prosody_info"], o["p_targets"], o["e_targets"])
