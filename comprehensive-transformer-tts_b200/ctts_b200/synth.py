"""Deterministic synthetic weights and batches (there are no checkpoints or datasets offline).

`synthetic_state_dict` fills every tensor of `spec.parameter_spec` from a generator seeded by the
tensor's NAME, so the same weights can be rebuilt on any machine without the reference (the
golden fixtures in tests/golden were produced by loading exactly these tensors into the
reference model).  The values are deliberately *not* a neutral init: LayerNorm/BatchNorm affine
terms, biases and running statistics are all non-trivial so that padding hazards (SURVEY.md
section 7, H4: conv halos see LN(0) = beta, PostNet runs over padded frames) show up in parity tests.

`ljspeech_batch` / `sweep_batch` follow BASELINE.md section 2 "Inputs".
"""
import hashlib
import math

import numpy as np
import torch


def _gen(name, seed):
    h = hashlib.sha256(("%d:%s" % (seed, name)).encode()).digest()
    g = torch.Generator()
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def synthetic_state_dict(spec_entries, seed=1234, pin_frames_per_phoneme=None):
    """name -> CPU fp32 tensor for every entry of `spec.parameter_spec(...)[0]`.

    pin_frames_per_phoneme=r sets duration_predictor.linear to weight 0 / bias log(1+r) so that the
    free-running inference path emits exactly r frames for every valid phoneme.
    """
    out = {}
    for name, shape, kind, init in spec_entries:
        g = _gen(name, seed)
        leaf = name.rsplit(".", 1)[-1]
        if init.startswith("tie:"):
            t = out[init[4:]]
        elif init == "sinusoid_interleaved":
            _, rows, d = shape
            pos = np.arange(rows, dtype=np.float64)[:, None]
            tab = pos / np.power(10000, 2 * (np.arange(d)[None, :] // 2) / d)
            tab[:, 0::2] = np.sin(tab[:, 0::2])
            tab[:, 1::2] = np.cos(tab[:, 1::2])
            t = torch.FloatTensor(tab).unsqueeze(0)
        elif init == "emb1":
            t = torch.randn(shape, generator=g) * 0.1
            t[0] = 0
        elif leaf in ("u_bias", "v_bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif init == "count":
            t = torch.zeros((), dtype=torch.long)
        elif init.startswith("linspace") or init.startswith("logspace"):
            _, lo, hi = init.split(":")
            if init.startswith("logspace"):
                t = torch.exp(torch.linspace(np.log(float(lo)), np.log(float(hi)), shape[0]))
            else:
                t = torch.linspace(float(lo), float(hi), shape[0])
        elif leaf == "_float_tensor":
            t = torch.zeros(shape)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "running_mean":
            t = 0.2 * torch.randn(shape, generator=g)
        elif leaf in ("pos_embed_alpha",):
            t = torch.full(shape, 0.7) + 0.1 * torch.randn(shape, generator=g)
        elif init == "ones":  # LN / BN scale
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf.startswith("bias") or init == "zeros":
            t = 0.05 * torch.randn(shape, generator=g)
        elif init.startswith("emb"):
            d = int(init.split(":")[1])
            t = torch.randn(shape, generator=g) * d ** -0.5
            t[0] = 0
        elif init == "normal01":
            t = torch.randn(shape, generator=g)
        elif init == "normal05":
            t = 0.5 * torch.randn(shape, generator=g)
        else:  # dense / conv weights: variance-preserving
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        if t is not None:
            out[name] = t.contiguous()
    k = "variance_adaptor.cwt_stats_layers.4.bias"
    if k in out:  # put the free-running f0 near 200 Hz (log-f0 mean 5.3, std 0.4) so the pitch quantiser sees many bins
        out[k] = torch.tensor([5.3, 0.4])
    k = "variance_adaptor.pitch_predictor.linear.bias"
    if k in out:  # pitch_type 'frame' / 'ph': free-running log2-f0 around 7.5 (~180 Hz), voiced / unvoiced logit around 0
        out[k] = torch.tensor([7.5, 0.0][: out[k].numel()])
        out["variance_adaptor.pitch_predictor.linear.weight"] = out["variance_adaptor.pitch_predictor.linear.weight"] * 4.0
    if pin_frames_per_phoneme is not None:
        out["variance_adaptor.duration_predictor.linear.weight"].zero_()
        out["variance_adaptor.duration_predictor.linear.bias"].fill_(math.log(1.0 + pin_frames_per_phoneme))
    return out


# ---------------------------------------------------------------------------------------------
def _beta_binomial_prior(S, M, scaling=1.0):
    """Beta-binomial attention prior [M, S]; same family as preprocessor/preprocessor.py:551-560."""
    from scipy.stats import betabinom
    rows = []
    k = np.arange(S)
    for i in range(1, M + 1):
        rows.append(betabinom(S - 1, scaling * i, scaling * (M + 1 - i)).pmf(k))
    return np.asarray(rows, dtype=np.float32)


def ljspeech_batch(batch=16, s_max=100, s_step=2, frames_per_phoneme=8, mode="infer", seed=0, vocab=361,
                   n_speakers=1, spk_dim=None):
    """Synthetic LJSpeech-shape batch (CPU tensors).

    mode "infer": (speakers, texts, src_lens, max_src_len) only -- free-running path.
    mode "teacher": adds mels / mel_lens / max_mel_len / p_targets / e_targets / d_targets with integer
        durations U[1, 15] (supervised branch, modules.py:1054-1057).
    mode "unsup": like teacher but with attn_priors instead of d_targets and frame-level energy
        (learn_alignment branch, modules.py:1031-1053).
    """
    g = torch.Generator()
    g.manual_seed(seed)
    src_lens = torch.tensor([max(s_max - s_step * b, 1) for b in range(batch)], dtype=torch.long)
    texts = torch.zeros(batch, s_max, dtype=torch.long)
    for b in range(batch):
        texts[b, : src_lens[b]] = torch.randint(1, vocab, (int(src_lens[b]),), generator=g)
    speakers = torch.randint(0, n_speakers, (batch,), generator=g) if n_speakers > 1 else torch.zeros(batch,
                                                                                                       dtype=torch.long)
    out = dict(speakers=speakers, texts=texts, src_lens=src_lens, max_src_len=int(s_max))
    if spk_dim:
        out["spker_embeds"] = torch.randn(batch, spk_dim, generator=g)
    if mode == "infer":
        return out
    dur = torch.zeros(batch, s_max, dtype=torch.long)
    for b in range(batch):
        dur[b, : src_lens[b]] = torch.randint(1, 16, (int(src_lens[b]),), generator=g)
    mel_lens = dur.sum(1)
    m_max = int(mel_lens.max())
    frame_mask = (torch.arange(m_max)[None, :] < mel_lens[:, None])
    mels = (torch.randn(batch, m_max, 80, generator=g) * 1.5 - 5.0) * frame_mask[:, :, None]
    f0 = (torch.randn(batch, m_max, generator=g) * 0.3 + 7.5) * frame_mask
    uv = (torch.rand(batch, m_max, generator=g) < 0.2).float() * frame_mask
    cwt = torch.randn(batch, m_max, 10, generator=g) * frame_mask[:, :, None]
    csum = torch.cumsum(dur, 1)
    mel2ph = (torch.arange(m_max)[None, None, :] >= torch.nn.functional.pad(csum, [1, -1])[:, :, None]) & \
             (torch.arange(m_max)[None, None, :] < csum[:, :, None])
    mel2ph = (mel2ph.long() * torch.arange(1, s_max + 1)[None, :, None]).sum(1)
    p_targets = dict(f0=f0, uv=uv, cwt_spec=cwt, f0_mean=torch.full((batch,), 5.3), f0_std=torch.full((batch,), 0.3),
                     mel2ph=mel2ph)
    out.update(mels=mels, mel_lens=mel_lens, max_mel_len=m_max, p_targets=p_targets)
    if mode == "teacher":
        out["e_targets"] = torch.randn(batch, s_max, generator=g) * (torch.arange(s_max)[None] < src_lens[:, None])
        out["d_targets"] = dur
    elif mode == "unsup":
        out["e_targets"] = torch.randn(batch, m_max, generator=g) * frame_mask
        pri = torch.zeros(batch, s_max, m_max)
        for b in range(batch):
            S, M = int(src_lens[b]), int(mel_lens[b])
            pri[b, :S, :M] = torch.from_numpy(_beta_binomial_prior(S, M)).t()
        out["attn_priors"] = pri
        del p_targets["mel2ph"]  # produced by the model from the hard alignment, modules.py:1053
    else:
        raise ValueError(mode)
    return out


def valid_frames(src_lens, frames_per_phoneme):
    return int(src_lens.sum().item()) * int(frames_per_phoneme)
