"""CPU: the oracle (oracle/ctts_oracle.py) against the golden fixtures produced by the unmodified reference."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
from oracle import ctts_oracle as O  # noqa: E402

EXACT = ("d_rounded", "mel_lens", "src_lens", "src_masks", "mel_masks", "p_targets.mel2ph",
         "attn_outs.1", "attn_outs.2")  # attn_hard (MAS path) and attn_hard_dur are integer-valued: bit-exact


def run_oracle(name):
    (p, m, t), sd, batch = cases.build_case(name)
    args, kw = cases.call_kwargs(batch)
    taps = {}
    with torch.no_grad():
        out = O.comp_trans_tts_forward(sd, p, m, t, *args, taps=taps, **kw)
    return cases.flatten_outputs(out), taps, batch


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_reference_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    flat, taps, batch = run_oracle(name)
    assert np.array_equal(gold["in.texts"], batch["texts"].numpy()), "synthetic inputs drifted from the fixtures"
    checked = 0
    for key in gold.files:
        if not key.startswith("ref."):
            continue
        k = key[4:]
        assert k in flat, "oracle output lacks %s" % k
        a, b = gold[key], flat[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if k in EXACT or a.dtype.kind in "biu":
            assert np.array_equal(a, b), "%s must be bit-exact" % k
        else:
            np.testing.assert_allclose(b, a, atol=2e-5, rtol=1e-4, err_msg=k)
        checked += 1
    assert checked >= 10
    for k in ("encoder_out", "decoder_in", "decoder_out"):
        np.testing.assert_allclose(taps[k].numpy()[:, ::cases.TAP_STRIDE], gold["tap." + k], atol=2e-5, rtol=1e-4,
                                   err_msg=k)


def test_golden_has_no_trivial_quantisers(golden_dir):
    """The fixtures must exercise the quantisers: several distinct durations / pitch values / padded rows."""
    g = np.load(os.path.join(golden_dir, "fs2_infer_c1.npz"))
    assert set(np.unique(g["ref.d_rounded"])) == {0.0, 3.0}
    assert g["ref.mel_lens"].tolist() == [300, 273]
    f0 = g["ref.p_predictions.f0_denorm"]
    assert (f0 == 0).any() and (f0 > 100).any()
    t = np.load(os.path.join(golden_dir, "fs2_teacher.npz"))
    assert len(np.unique(t["ref.d_rounded"])) >= 10
