"""GPU: the five BASELINE.json configurations at (or near) their quoted shapes, forward pass against the CPU oracle.

configs[0] / [1] (fs2 batch 2 / batch 16 inference) are covered by test_gpu_e2e.py; this file adds
  configs[2]  conformer + unsupervised alignment, LJSpeech shape, batch 16            (eval-mode forward of the training batch)
  configs[3]  fastformer, VCTK multi-speaker shape, batch 32                           (eval-mode forward of the training batch)
  configs[4]  transformer_fs2 + liu2021 prosody, mel-length sweep 64 / 256 / 1024, batch 16
Backward / optimizer steps of configs[2], [3] are not built yet.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
import ctts_b200  # noqa: E402
from ctts_b200 import spec, synth  # noqa: E402
from oracle import ctts_oracle as O  # noqa: E402

DEV = "cuda:0"


def to_dev(v):
    if torch.is_tensor(v):
        return v.to(DEV)
    if isinstance(v, dict):
        return {k: to_dev(x) for k, x in v.items()}
    return v


def run_both(cfgs, sd, batch):
    p, m, t = cfgs
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd, strict=True)
    net.to(DEV)
    args, kw = cases.call_kwargs(batch)
    out = net(*[to_dev(a) for a in args], **{k: to_dev(v) for k, v in kw.items()})
    args, kw = cases.call_kwargs(batch)
    with torch.no_grad():
        ref = O.comp_trans_tts_forward(sd, p, m, t, *args, **kw)
    return out, ref


def assert_mels(out, ref, flip_budget=0.0):
    for i in (0, 1):
        got, want = out[i].cpu().numpy(), ref[i].numpy()
        assert got.shape == want.shape
        bad = np.abs(got - want) > (1e-3 + 1e-2 * np.abs(want))
        assert bad.mean() <= flip_budget, "%.3f%% of mel[%d] outside 1e-3 abs + 1e-2 rel" % (100 * bad.mean(), i)


def _log(record):
    import json
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/inference_parity.jsonl", "a") as f:
        f.write(json.dumps(record, sort_keys=True) + "\n")


@pytest.mark.parametrize("s_max", [48, 100])
def test_config2_conformer_unsupervised_batch16(s_max):
    """BASELINE configs[2] at its quoted shape (S ~ 100, B 16) and at a shorter one."""
    p, m, t = ctts_b200.builtin_configs("LJSpeech", block_type="conformer", learn_alignment=True)
    sd = synth.synthetic_state_dict(spec.parameter_spec(p, m)[0])
    batch = synth.ljspeech_batch(batch=16, s_max=s_max, s_step=1 if s_max < 100 else 2, mode="unsup", seed=21)
    out, ref = run_both((p, m, t), sd, batch)
    assert torch.equal(out[10][1].cpu(), ref[10][1]), "MAS path (attn_hard) must be bit-exact"
    assert torch.equal(out[5].cpu(), ref[5]) and torch.equal(out[9].cpu(), ref[9])
    np.testing.assert_allclose(out[10][0].cpu().numpy(), ref[10][0].numpy(), atol=1e-5, rtol=1e-4)
    assert_mels(out, ref)
    _log({"test": "config2_conformer_unsup_b16", "s_max": s_max,
          "postnet_max_err": float((out[1].cpu() - ref[1]).abs().max()), "mas_path_mismatches": 0})


def test_config3_fastformer_vctk_batch32():
    p, m, t = ctts_b200.builtin_configs("VCTK", block_type="fastformer", learn_alignment=True)
    sd = synth.synthetic_state_dict(spec.parameter_spec(p, m)[0])
    batch = synth.ljspeech_batch(batch=32, s_max=50, s_step=1, mode="unsup", seed=22, spk_dim=512)
    out, ref = run_both((p, m, t), sd, batch)
    assert torch.equal(out[10][1].cpu(), ref[10][1]) and torch.equal(out[5].cpu(), ref[5])
    # fastformer's inverted -10000 mask makes it ill-conditioned: flip accounting
    assert_mels(out, ref, flip_budget=0.01)
    bad = (out[1].cpu() - ref[1]).abs() > (1e-3 + 1e-2 * ref[1].abs())
    _log({"test": "config3_fastformer_vctk_b32", "fraction_outside_tolerance": float(bad.float().mean()),
          "postnet_max_err": float((out[1].cpu() - ref[1]).abs().max())})


@pytest.mark.parametrize("M", [64, 256, 1024])
def test_config4_fs2_liu2021_length_sweep(M):
    """S = M / 8 phonemes, every utterance full length, 8 frames per phoneme; M = 1024 exceeds max_seq_len = 1000
    (dynamic sinusoid table, blocks.py:88-95)."""
    p, m, t = ctts_b200.builtin_configs("LJSpeech", block_type="transformer_fs2", learn_alignment=False, prosody="liu2021")
    sd = synth.synthetic_state_dict(spec.parameter_spec(p, m)[0], pin_frames_per_phoneme=8)
    batch = synth.ljspeech_batch(batch=16, s_max=M // 8, s_step=0, mode="infer", seed=23)
    out, ref = run_both((p, m, t), sd, batch)
    assert out[0].shape == (16, M, 80)
    assert torch.equal(out[9].cpu(), ref[9]) and torch.equal(out[5].cpu(), ref[5])
    np.testing.assert_allclose(out[11][2].cpu().numpy(), ref[11][2].numpy(), atol=1e-4, rtol=1e-3)   # utterance prosody
    np.testing.assert_allclose(out[11][3].cpu().numpy(), ref[11][3].numpy(), atol=1e-4, rtol=1e-3)   # phoneme prosody
    pidx, pref = O.f0_to_coarse(out[2]["f0_denorm"].cpu()), O.f0_to_coarse(ref[2]["f0_denorm"])
    bins = sd["variance_adaptor.energy_bins"]
    eidx, eref = torch.bucketize(out[3].cpu(), bins), torch.bucketize(ref[3], bins)
    p_flips, e_flips = int((pidx != pref).sum()), int((eidx != eref).sum())
    print("M=%d: prosody err u %.2e p %.2e | e_pred err %.2e | pitch flips %d energy flips %d of %d phonemes" % (
        M, (out[11][2].cpu() - ref[11][2]).abs().max(), (out[11][3].cpu() - ref[11][3]).abs().max(),
        (out[3].cpu() - ref[3]).abs().max(), p_flips, e_flips, out[3].numel()))
    # Flip accounting.  The quantiser inputs agree to FP32 noise (asserted), but a value that sits within
    # that noise of a bucket edge can land on the other side; the flipped embedding row then moves its whole utterance
    # (through self-attention) by far more than the tolerance.  That is a property of quantising, not an arithmetic error:
    # utterances WITHOUT a flip must meet the tolerance everywhere, and flips must be rare (<= 1 per 500 phonemes).
    _log({"test": "config4_fs2_liu2021", "M": M, "pitch_flips": p_flips, "energy_flips": e_flips, "phonemes": out[3].numel(),
          "e_pred_max_err": float((out[3].cpu() - ref[3]).abs().max())})
    assert (out[3].cpu() - ref[3]).abs().max() < 3e-5 and (out[2]["cwt"].cpu() - ref[2]["cwt"]).abs().max() < 1e-4
    assert p_flips + e_flips <= max(2, out[3].numel() // 500), (p_flips, e_flips)
    clean = ~((pidx != pref).any(1) | (eidx != eref).any(1))
    assert int(clean.sum()) >= 14
    for i in (0, 1):
        got, want = out[i].cpu().numpy()[clean.numpy()], ref[i].numpy()[clean.numpy()]
        np.testing.assert_allclose(got, want, atol=1e-3, rtol=1e-2)
