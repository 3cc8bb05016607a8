"""GPU: the whole forward path (model.forward -> C ABI -> sm_100a kernels) against the reference's golden
fixtures and against the CPU oracle.

north_star tolerance: mel outputs within 1e-3 abs + 1e-2 rel (fp32); LengthRegulator indices bit-exact.
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
from oracle import ctts_oracle as O  # noqa: E402

DEV = "cuda:0"
MEL_ATOL, MEL_RTOL = 1e-3, 1e-2
EXACT = ("d_rounded", "mel_lens", "src_lens", "src_masks", "mel_masks", "p_targets.mel2ph",
         "attn_outs.1", "attn_outs.2")  # attn_hard (MAS path) and attn_hard_dur are integer-valued: bit-exact


def to_dev(v):
    if torch.is_tensor(v):
        return v.to(DEV)
    if isinstance(v, dict):
        return {k: to_dev(x) for k, x in v.items()}
    return v


def run_case(name):
    (p, m, t), sd, batch = cases.build_case(name)
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd, strict=True)
    net.to(DEV)
    args, kw = cases.call_kwargs(batch)
    out = net(*[to_dev(a) for a in args], **{k: to_dev(v) for k, v in kw.items()})
    torch.cuda.synchronize()
    return out, (p, m, t), sd, batch


def check_against(flat, gold, prefix):
    n = 0
    for key in gold:
        if not key.startswith(prefix):
            continue
        k = key[len(prefix):]
        assert k in flat, "output lacks %s" % k
        a, b = np.asarray(gold[key]), flat[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if k in EXACT or a.dtype.kind in "biu":
            assert np.array_equal(a, b), "%s must be bit-exact" % k
        elif k in ("mel", "postnet_mel"):
            np.testing.assert_allclose(b, a, atol=MEL_ATOL, rtol=MEL_RTOL, err_msg=k)
        else:
            np.testing.assert_allclose(b, a, atol=1e-3, rtol=1e-3, err_msg=k)
        n += 1
    return n


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_forward_matches_reference_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    out, _, _, _ = run_case(name)
    flat = cases.flatten_outputs(out)
    assert check_against(flat, {k: gold[k] for k in gold.files}, "ref.") >= 10
    # the fp32 path should in fact be far inside the tolerance
    scale = max(1.0, float(np.abs(gold["ref.postnet_mel"]).max()) / 4.0)
    assert np.abs(flat["postnet_mel"] - gold["ref.postnet_mel"]).max() < 2e-4 * scale


def test_forward_matches_oracle_batch16():
    """BASELINE configs[1] shape (B 16, S 100..70) at 4 frames / phoneme against the CPU oracle."""
    p, m, t = ctts_b200.builtin_configs("LJSpeech", learn_alignment=False)
    from ctts_b200 import spec, synth
    sd = synth.synthetic_state_dict(spec.parameter_spec(p, m)[0], pin_frames_per_phoneme=4)
    batch = synth.ljspeech_batch(batch=16, s_max=100, s_step=2, mode="infer")
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd)
    net.to(DEV)
    args, _ = cases.call_kwargs(batch)
    out = net(*[to_dev(a) for a in args])
    with torch.no_grad():
        ref = O.comp_trans_tts_forward(sd, p, m, t, *args)
    assert torch.equal(out[9].cpu(), ref[9]) and torch.equal(out[5].cpu(), ref[5])
    assert out[0].shape == (16, 400, 80)
    pidx, pref = O.f0_to_coarse(out[2]["f0_denorm"].cpu()), O.f0_to_coarse(ref[2]["f0_denorm"])
    assert int((pidx != pref).sum()) == 0, "pitch bucket flips"
    for i in (0, 1):
        np.testing.assert_allclose(out[i].cpu().numpy(), ref[i].numpy(), atol=MEL_ATOL, rtol=MEL_RTOL)


def test_longest_utterance_is_independent_and_controls_work():
    """Utterance 0 (the longest: no padding anywhere) must not depend on what else is in the batch, bit for bit;
    d_control scales the regulated length.  (Shorter utterances DO depend on the padded length in the reference --
    CWT normalisation and PostNet run over padded frames, SURVEY.md H4 -- and the golden tests cover that.)"""
    (p, m, t), sd, batch = cases.build_case("fs2_infer_c1")
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd)
    net.to(DEV)
    both = net(batch["speakers"].to(DEV), batch["texts"].to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"])
    other = batch["texts"].clone()
    other[1, :50] = other[1, :50].flip(0)
    swapped = net(batch["speakers"].to(DEV), other.to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"])
    assert torch.equal(both[1][0], swapped[1][0])
    assert not torch.equal(both[1][1], swapped[1][1])
    slow = net(batch["speakers"].to(DEV), batch["texts"].to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"],
               d_control=2.0)
    assert slow[9].tolist() == [2 * v for v in both[9].tolist()]


def test_training_mode_is_refused_not_faked():
    (p, m, t), sd, batch = cases.build_case("fs2_infer_c1")
    net = ctts_b200.CompTransTTS(p, m, t).train().to(DEV)
    with pytest.raises(NotImplementedError):
        net(batch["speakers"].to(DEV), batch["texts"].to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"])
